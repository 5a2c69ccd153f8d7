#!/bin/bash
mkdir -p gpurun_out/c66
O=gpurun_out/c66
M2=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed.sum
timeout 600 ncu --metrics $M2 --clock-control none --csv --log-file $O/deg_launches.csv python tools/ncu_targets.py degrade > $O/ncu_deg.log 2>&1
python tools/ncu_table.py $O/deg_launches.csv 15 | tail -n 16
