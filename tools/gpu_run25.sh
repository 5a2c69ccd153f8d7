#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -3 gpurun_out/bench4.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench4.json') if x.startswith('{')]
d=json.loads(l[-1])
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'], d['roofline'])
print("degradation", d.get('degradation'))
print("training", {k:d['training'].get(k) for k in ('value','ms_per_step','roofline','cuda_graph','error')})
PY
