#!/bin/bash
for rep in 1 2; do
for v in v3 v5; do echo -n "$v: "; RESR_LIB_PATH=$PWD/build/variants/libresr_$v.so timeout 120 python tools/power_probe.py 3 2>&1 | tail -1; done
for ne in 2 3; do echo -n "v3 nepi=$ne: "; RESR_CONV_NEPI=$ne RESR_LIB_PATH=$PWD/build/variants/libresr_v3.so timeout 120 python tools/power_probe.py 3 2>&1 | tail -1; done
done
