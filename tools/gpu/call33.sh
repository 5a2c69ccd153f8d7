#!/bin/bash
mkdir -p gpurun_out/c33
O=gpurun_out/c33
for d in 0 1 2 3 4 6 7; do RESR_WGRAD_MN_DEBUG=$d timeout 120 python tools/time_wgrad_mn.py >> $O/probe.log 2>&1; done
cat $O/probe.log
