#!/bin/bash
mkdir -p gpurun_out/c7
O=gpurun_out/c7
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
timeout 900 python bench.py > $O/bench1.json 2> $O/bench1.err; echo "exit $?" >> $O/bench1.err
timeout 300 python bench.py --precision bf16 --no-train --no-degrade --no-tiled --no-cpu > $O/bench_bf16.json 2> $O/bench_bf16.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum
timeout 600 ncu --metrics $M --clock-control none --launch-skip 353 -c 352 --csv --log-file $O/gen_launches.csv python tools/ncu_targets.py gen > $O/ncu_gen.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/deg_launches.csv python tools/ncu_targets.py degrade > $O/ncu_deg.log 2>&1
tail -n 3 $O/tests.log; cat $O/bench1.err | tail -n 5; head -c 600 $O/bench1.json; echo; head -c 300 $O/bench_bf16.json; echo; tail -n 2 $O/ncu_gen.log $O/ncu_deg.log
