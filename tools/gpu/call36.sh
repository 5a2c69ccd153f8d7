#!/bin/bash
mkdir -p gpurun_out/c36
O=gpurun_out/c36
# no loads (4) + no bias (16) = 20 base; +32 one dx only; +256 unshifted; +64 B K-major; +128 A K-major; +192 both
for d in 20 52 276 84 148 212 22; do RESR_WGRAD_MN_DEBUG=$d timeout 120 python tools/time_wgrad_mn.py 2>&1 | grep "per-kernel\|debug=" >> $O/probe.log; done
cat $O/probe.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv
