#!/bin/bash
mkdir -p gpurun_out/c21
O=gpurun_out/c21
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "exit $?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench1.json 2> $O/bench1.err; echo "exit $?" >> $O/bench1.err
timeout 300 python bench.py --precision bf16 --no-train --no-degrade --no-tiled --no-cpu > $O/bench_bf16.json 2> $O/bench_bf16.err
timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
tail -n 2 $O/smoke.log; tail -n 3 $O/bench1.err; tail -n 3 $O/tests.log; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c21/bench1.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
t=d['tiled']; print('tiled', t.get('value'), t.get('ms_per_step'), 'u8', t.get('u8_image',{}).get('value'), t.get('u8_image',{}).get('ms_per_step'), t.get('error'))
g=d['degradation']; print('deg', g.get('value'), g.get('ms_per_step'), 'e2e', g.get('e2e',{}).get('value'), 'cpu', g.get('cpu_baseline'), g.get('error'))
print('train', d['training'].get('value'), d['training'].get('ms_per_step'), d['training'].get('error'))
print('cpu', d.get('cpu_baseline'))
b=json.loads([l for l in open('gpurun_out/c21/bench_bf16.json').read().splitlines() if l.startswith('{')][-1])
print('bf16', b['value'], b['ms_per_step'], b['roofline']['frac'])
PY
