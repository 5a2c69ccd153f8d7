#!/bin/bash
mkdir -p gpurun_out/c51
O=gpurun_out/c51
timeout 2400 python -m pytest tests -x -q -m gpu -s > $O/t_all.log 2>&1; echo "exit $?" >> $O/t_all.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
grep "cfg4\|cfg3\|cfg5\|passed\|failed\|exit" $O/t_all.log | tail -n 12; tail -n 2 $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c51/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'clocks', d.get('clocks'))
t=d['training']; print('train', t.get('ms_per_step'), t.get('value'), (t.get('other_precision') or {}).get('ms_per_step'), t.get('error'))
g=d['degradation']; print('deg', g['value'], g['ms_per_step'], g['e2e']['value'])
print('tiled', d.get('tiled', {}).get('value'), d.get('tiled', {}).get('ms_per_step'))
PY
