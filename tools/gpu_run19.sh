#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench3.json'))
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], "frac", d['roofline']['frac'], d['clocks'])
print("degradation", d.get('degradation',{}).get('value'), d.get('degradation',{}).get('ms_per_step'), d.get('degradation',{}).get('error'))
print("training", d.get('training'))
print("cpu", d.get('cpu_baseline'))
PY
