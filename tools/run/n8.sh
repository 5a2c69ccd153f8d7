#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -2 gpurun_out/bench_n8.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_n8.json') if x.startswith('{')]
d=json.loads(l[-1])
print("N", d['n_gpus'], "value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'])
print("degradation", d['degradation'].get('value'), d['degradation'].get('n_gpus'), d['degradation'].get('large_batch'))
print("training", {k:d['training'].get(k) for k in ('value','ms_per_step','n_gpus','cuda_graph','error')})
PY
