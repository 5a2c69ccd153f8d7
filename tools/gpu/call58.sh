#!/bin/bash
mkdir -p gpurun_out/c58
O=gpurun_out/c58
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-degrade --no-train --no-cpu --no-weak --no-other-precision > $O/bench8.json 2> $O/bench8.err; echo "exit $?" >> $O/bench8.err
tail -n 2 $O/bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c58/bench8.json').read().strip().splitlines()[-1])
print('cfg3 strong', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
t=d['tiled']; print('tiled nccl', t['value'], t['ms_per_step'], 'u8', t['u8_image']['value'], t['u8_image']['ms_per_step'])
sh=t['shared_pinned_host']; print('shared fp32', sh['fp32'].get('value'), sh['fp32'].get('ms_per_step'), sh['fp32'].get('matches_nccl_gather'), 'u8', sh['u8_image'].get('value'), sh['u8_image'].get('ms_per_step'), sh['u8_image'].get('matches_nccl_gather'), sh['fp32'].get('error'))
PY
