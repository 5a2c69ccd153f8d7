#!/bin/bash
mkdir -p gpurun_out/c35
O=gpurun_out/c35
for d in 7 15 23 31 16 18; do RESR_WGRAD_MN_DEBUG=$d timeout 120 python tools/time_wgrad_mn.py 2>&1 | grep "per-kernel\|debug=" >> $O/probe.log; done
cat $O/probe.log
