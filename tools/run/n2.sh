#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; tail -2 gpurun_out/bench_ref_n2.err; tail -c 400 gpurun_out/bench_ref_n2.json
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_n2.json') if x.startswith('{')]
d=json.loads(l[-1])
print("N", d['n_gpus'], "value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'])
print("degradation", d['degradation'].get('value'), d['degradation'].get('large_batch'))
print("training", {k:d['training'].get(k) for k in ('value','ms_per_step','n_gpus','cuda_graph','error')})
PY
