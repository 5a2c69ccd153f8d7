#!/bin/bash
for rep in 1 2; do
for f in 0 64 128 192; do RESR_CONV_DBGFLAGS=$f RESR_LIB_PATH=$PWD/build/variants/libresr_vE.so timeout 120 python tools/power_probe.py 3 2>&1 | tail -1; done
done
