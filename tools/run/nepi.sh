#!/bin/bash
for rep in 1 2; do
echo -n "default(3): "; timeout 120 python tools/power_probe.py 3 2>&1 | tail -1
echo -n "nepi=2:     "; RESR_CONV_NEPI=2 timeout 120 python tools/power_probe.py 3 2>&1 | tail -1
done
