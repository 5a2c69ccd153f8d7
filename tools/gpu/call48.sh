#!/bin/bash
mkdir -p gpurun_out/c48
O=gpurun_out/c48
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_overlap_check.py > $O/ddp8.log 2>&1; echo "exit $?" >> $O/ddp8.log
grep -v "^W\|^\[W\|Warning\|warn" $O/ddp8.log | tail -n 8
