#!/bin/bash
mkdir -p gpurun_out/c60
O=gpurun_out/c60
timeout 600 python -m pytest tests/test_degrade_gpu.py -x -q -m gpu -k "u8_images or augment" > $O/t.log 2>&1; tail -n 3 $O/t.log
timeout 300 python bench.py --no-train --no-tiled --no-cpu --steps 5 --no-other-precision > $O/bench.json 2> $O/bench.err
tail -n 2 $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c60/bench.json').read().strip().splitlines()[-1])['degradation']
print(d.get('error'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8_images']['value'], d['e2e_u8_images']['ms_per_step'])
PY
