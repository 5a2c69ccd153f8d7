#!/bin/bash
mkdir -p gpurun_out/c55
O=gpurun_out/c55
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --launch-skip 353 -c 352 --csv --log-file $O/gen_launches.csv python tools/ncu_targets.py gen > $O/ncu_gen.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair --launch-skip 357 -c 5 -o $O/rdb_full python tools/ncu_targets.py rdb > $O/ncu_rdb.log 2>&1
M2=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed.sum
timeout 600 ncu --metrics $M2 --clock-control none --csv --log-file $O/deg_launches.csv python tools/ncu_targets.py degrade > $O/ncu_deg.log 2>&1
# the bench command itself (launch list of the timed headline forwards)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1060 -c 704 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-degrade --no-train --no-tiled --no-other-precision > $O/ncu_bench.log 2>&1
python tools/ncu_table.py $O/gen_launches.csv 352 | tail -n 3
python tools/ncu_table.py $O/deg_launches.csv 17 | tail -n 18
ls -la $O
