#!/bin/bash
# 1 no stores, 2 L2-hot loads, 4 quarter MMAs, 8 epilogue drains only, 16 no MMAs, 32 no loads
for f in 63 31 47 55 39 9 24 40 16 8 32; do RESR_CONV_DBGFLAGS=$f timeout 120 python tools/power_probe.py 2 2>&1 | tail -1; done
