#!/bin/bash
mkdir -p gpurun_out/c54
O=gpurun_out/c54
timeout 300 python bench.py --no-train --no-tiled --no-cpu --steps 5 --no-other-precision > $O/bench.json 2> $O/bench.err
tail -n 3 $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c54/bench.json').read().strip().splitlines()[-1])['degradation']
print(d.get('error'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8_images']['value'], d['e2e_u8_images']['ms_per_step'])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
