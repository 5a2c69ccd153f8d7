#!/bin/bash
mkdir -p gpurun_out/c64
O=gpurun_out/c64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:usm_fused51 --launch-skip 4 -c 2 -o $O/usm_fused python tools/ncu_targets.py degrade > $O/ncu.log 2>&1
ls -la $O
