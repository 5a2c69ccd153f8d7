#!/bin/bash
mkdir -p gpurun_out/c41
O=gpurun_out/c41
for d in 22 1046 20 1044 0 1024; do N=16 RESR_WGRAD_MN_DEBUG=$d timeout 120 python tools/time_wgrad_mn.py 2>&1 | grep "per-kernel" >> $O/probe.log; done
cat $O/probe.log
