#!/bin/bash
mkdir -p gpurun_out/c59
O=gpurun_out/c59
S0=$(date +%s); timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
tail -n 2 $O/bench.err; echo "bench wall $(( $(date +%s) - S0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c59/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ['metric','value','unit','n_gpus','steps','warmup','ms_per_step','scaling','vs_baseline','dtype','gpu_launches']})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['traffic'], 'cpu', d['cpu_baseline']['value'])
print('other', d['config'].get('other_precision'))
t=d['training']; print('train', t['ms_per_step'], t['value'], t['other_precision']['ms_per_step'], t['optimizer'])
g=d['degradation']; print('deg', g['value'], g['ms_per_step'], g['e2e']['value'], g['e2e_u8_images']['value'], g['large_batch']['value'], g['cpu_baseline'].get('value'))
t=d['tiled']; print('tiled', t['value'], t['ms_per_step'], t['nccl_gather']['ms_per_step'], t['shared_pinned_host']['fp32'].get('ms_per_step'), t['u8_image']['ms_per_step'])
PY
