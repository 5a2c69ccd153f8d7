#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_generator_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-train > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_q.json') if x.startswith('{')]
d=json.loads(l[-1])
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e'])
PY
