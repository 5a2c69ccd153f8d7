#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/library_bar.py 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train2.csv python tools/time_train.py 16 64 64 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 352 --csv --log-file gpurun_out/launches_gen4.csv python tools/time_generator.py 64 128 128 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 357 -c 5 -o gpurun_out/prof_rdb_v4 -f python tools/time_generator.py 64 128 128 > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu4.log
