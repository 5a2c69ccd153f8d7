#!/bin/bash
mkdir -p gpurun_out/c74
O=gpurun_out/c74
timeout 2400 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1; echo "exit $?" >> $O/t_all.log
tail -n 3 $O/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
tail -n 1 $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c74/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'clocks', d['clocks'])
t=d['training']; print('train', t['ms_per_step'], t['value'], t['other_precision']['ms_per_step'], t['optimizer']['ms_per_step'])
g=d['degradation']; print('deg', g['value'], g['ms_per_step'], g['e2e']['value'], g['e2e_u8_images']['value'], g['large_batch']['value'], g['roofline']['frac'], g['fma']['frac'])
t=d['tiled']; print('tiled', t['value'], t['ms_per_step'], t['u8_image']['ms_per_step'])
PY
