#!/bin/bash
mkdir -p gpurun_out/c8
O=gpurun_out/c8
timeout 900 python -m pytest tests/test_degrade_gpu.py tests/test_parity_at_size_gpu.py -x -q -m gpu -s -k "not cfg3 and not cfg4 and not cfg5" > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
timeout 300 python bench.py --no-train --no-tiled --no-cpu --steps 5 > $O/bench.json 2> $O/bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/deg_launches.csv python tools/ncu_targets.py degrade > $O/ncu_deg.log 2>&1
grep -n "passed\|failed\|exit" $O/tests.log | tail -n 4; python - <<'PY'
import json
d=json.load(open('gpurun_out/c8/bench.json'))['degradation']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'fma', d['fma']['frac'], 'big', d['large_batch']['value'])
PY
