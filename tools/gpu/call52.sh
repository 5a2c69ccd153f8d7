#!/bin/bash
mkdir -p gpurun_out/c52
O=gpurun_out/c52
timeout 1200 python -m pytest tests/test_degrade_gpu.py tests/test_gan_losses_gpu.py tests/test_parity_at_size_gpu.py -x -q -m gpu -k "not cfg3 and not cfg4 and not cfg5 and not pixel_loss" > $O/t_deg.log 2>&1; echo "exit $?" >> $O/t_deg.log
tail -n 3 $O/t_deg.log
timeout 300 python tools/time_degrade.py 2>&1 | tail -n 2
timeout 300 python bench.py --no-train --no-tiled --no-cpu --steps 5 --no-other-precision > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c52/bench.json').read().strip().splitlines()[-1])['degradation']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'big', d['large_batch']['value'])
PY
