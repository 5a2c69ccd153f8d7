#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 357 -c 5 -o gpurun_out/prof_rdb_v3 -f python tools/time_generator.py 64 128 128 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
