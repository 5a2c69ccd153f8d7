#!/bin/bash
mkdir -p gpurun_out/c39
O=gpurun_out/c39
for d in 84 212 148; do N=16 RESR_WGRAD_MN_DEBUG=$d timeout 120 python tools/time_wgrad_mn.py 2>&1 | tail -n 4 >> $O/probe.log; done
cat $O/probe.log
