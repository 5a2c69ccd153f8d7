#!/bin/bash
mkdir -p gpurun_out/c49
O=gpurun_out/c49
RESR_NCCL_HIGH_PRIO=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_overlap_check.py > $O/ddp8_hp.log 2>&1; echo "exit $?" >> $O/ddp8_hp.log
grep -v "^W\|^\[W\|Warning\|warn" $O/ddp8_hp.log | tail -n 7
