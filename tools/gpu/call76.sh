#!/bin/bash
mkdir -p gpurun_out/c76
O=gpurun_out/c76
timeout 300 python bench.py --no-degrade --no-tiled --no-cpu --steps 10 > $O/bench.json 2> $O/bench.err
tail -n 3 $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c76/bench.json').read().strip().splitlines()[-1])
t=d['training']; print('train', t.get('ms_per_step'), t.get('value'), 'e2e', t.get('e2e'), t.get('loss'), (t.get('other_precision') or {}).get('ms_per_step'), t.get('error'))
PY
