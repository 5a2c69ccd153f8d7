#!/bin/bash
mkdir -p gpurun_out/c38
O=gpurun_out/c38
for n in 4 16 64; do for d in 22 20 4 0; do N=$n RESR_WGRAD_MN_DEBUG=$d timeout 120 python tools/time_wgrad_mn.py 2>&1 | grep "per-kernel\|debug=" >> $O/probe.log; done; done
grep -v "debug=" $O/probe.log
