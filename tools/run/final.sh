#!/bin/bash
# end-of-round evidence run (one GPU): tests, smoke, bench, ncu launch list + full capture of one RDB
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 352 --csv --log-file gpurun_out/launches_gen5.csv python tools/time_generator.py 64 128 128 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 357 -c 5 -o gpurun_out/prof_rdb_v5 -f python tools/time_generator.py 64 128 128 > gpurun_out/ncu5.log 2>&1
tail -1 gpurun_out/ncu5.log
