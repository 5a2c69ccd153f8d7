#!/bin/bash
mkdir -p gpurun_out/c57
O=gpurun_out/c57
df -h /dev/shm | tail -n 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 --no-degrade --no-train --no-cpu --no-weak --no-other-precision > $O/bench2.json 2> $O/bench2.err; echo "exit $?" >> $O/bench2.err
tail -n 3 $O/bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c57/bench2.json').read().strip().splitlines()[-1])
t=d['tiled']; print('tiled nccl', t['value'], t['ms_per_step'], 'u8', t['u8_image']['ms_per_step'])
print('shared', json.dumps(t['shared_pinned_host'])[:900])
PY
