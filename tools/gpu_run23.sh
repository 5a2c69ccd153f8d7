#!/bin/bash
timeout 200 python tools/conv_phases.py 2>&1 | grep "cin=192" 
for f in 0 4 2 1 6 7; do RESR_CONV_DBGFLAGS=$f timeout 120 python tools/power_probe.py 4 2>&1 | tail -1; done
RESR_CONV_NEPI=1 timeout 120 python tools/power_probe.py 4 2>&1 | tail -1
nvidia-smi -q -d POWER | head -40
