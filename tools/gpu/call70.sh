#!/bin/bash
mkdir -p gpurun_out/c70
O=gpurun_out/c70
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench8.json 2> $O/bench8.err; echo "exit $?" >> $O/bench8.err
tail -n 2 $O/bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c70/bench8.json').read().strip().splitlines()[-1])
print('N=8 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'weak', (d.get('weak_scaling') or {}).get('value'))
t=d['training']; print('train', t.get('ms_per_step'), t.get('value'), (t.get('other_precision') or {}).get('value'), t.get('error'))
g=d['degradation']; print('deg', g.get('value'), g.get('ms_per_step'), g.get('error'))
t=d['tiled']; print('tiled', t.get('value'), t.get('ms_per_step'), 'nccl', t['nccl_gather']['ms_per_step'], 'u8', t['u8_image']['ms_per_step'], t['shared_pinned_host']['u8_image'].get('ms_per_step'))
print('clocks', d.get('clocks'))
PY
