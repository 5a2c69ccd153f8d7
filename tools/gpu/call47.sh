#!/bin/bash
mkdir -p gpurun_out/c47
O=gpurun_out/c47
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_overlap_check.py > $O/ddp.log 2>&1; echo "exit $?" >> $O/ddp.log
grep -v "^W\|^\[W\|Warning\|warn" $O/ddp.log | tail -n 8
RESR_NO_GRAPH=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ddp_overlap_check.py > $O/ddp_eager.log 2>&1; echo "exit $?" >> $O/ddp_eager.log
grep -v "^W\|^\[W\|Warning\|warn" $O/ddp_eager.log | tail -n 6
