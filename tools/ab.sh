#!/bin/bash
# A/B of library builds on ONE box (the power-capped clocks differ from box to box): tools/ab.sh v0 v1 ...
for rep in 1 2; do
for v in "$@"; do
  echo -n "$v: "; RESR_LIB_PATH=$PWD/build/variants/libresr_$v.so timeout 120 python tools/power_probe.py 3 2>&1 | tail -1
done; done
