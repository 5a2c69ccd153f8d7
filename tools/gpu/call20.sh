#!/bin/bash
mkdir -p gpurun_out/c20
O=gpurun_out/c20
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "exit $?" >> $O/bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $O/ref_n8.json 2> $O/ref_n8.err
tail -n 3 $O/bench_n8.err $O/bench_n4.err; python - <<'PY'
import json
for n in (8, 4):
    try:
        line=[l for l in open(f'gpurun_out/c20/bench_n{n}.json').read().splitlines() if l.startswith('{')][-1]
        d=json.loads(line)
        print(n, 'strong', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'weak', (d.get('weak') or {}).get('value'), 'tiled', (d.get('tiled') or {}).get('value'), (d.get('tiled') or {}).get('ms_per_step'), (d.get('tiled') or {}).get('error'), 'deg', (d.get('degradation') or {}).get('value'), 'train', (d.get('training') or {}).get('value'), (d.get('training') or {}).get('ms_per_step'))
    except Exception as e:
        print(n, 'parse error', e)
PY
