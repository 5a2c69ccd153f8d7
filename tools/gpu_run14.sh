#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/all.log 2>&1; echo "rc=$?" >> gpurun_out/all.log
grep -E "^E  .*(assert|Error)|passed|failed|rc=" gpurun_out/all.log | head -20
timeout 300 python tools/time_degrade.py 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; python -c "
import json; d=json.load(open('gpurun_out/bench2.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d.get('degradation',{}).get('value'), d.get('degradation',{}).get('ms_per_step'), d['clocks'])"; tail -3 gpurun_out/bench2.err
