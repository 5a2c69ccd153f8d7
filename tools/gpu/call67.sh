#!/bin/bash
mkdir -p gpurun_out/c67
O=gpurun_out/c67
timeout 2400 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1; echo "exit $?" >> $O/t_all.log
tail -n 3 $O/t_all.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench2.json 2> $O/bench2.err; echo "exit $?" >> $O/bench2.err
tail -n 2 $O/bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c67/bench2.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'weak', (d.get('weak_scaling') or {}).get('value'))
t=d['training']; print('train', t.get('ms_per_step'), t.get('value'), t.get('error'))
g=d['degradation']; print('deg', g.get('value'), g.get('ms_per_step'), g.get('error'))
t=d['tiled']; print('tiled', t.get('value'), t.get('ms_per_step'), t.get('error'))
PY
