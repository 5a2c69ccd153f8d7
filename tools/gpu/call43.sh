#!/bin/bash
mkdir -p gpurun_out/c43
O=gpurun_out/c43
timeout 600 python -m pytest tests/test_gan_losses_gpu.py -x -q -m gpu -s > $O/t_gan.log 2>&1; echo "exit $?" >> $O/t_gan.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1; echo "exit $?" >> $O/t_all.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
grep -v "^$" $O/t_gan.log | tail -n 12; tail -n 4 $O/t_all.log; tail -n 2 $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c43/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
t=d['training']; print('train', t.get('ms_per_step'), t.get('value'), t.get('other_precision'), t.get('error'))
g=d['degradation']; print('deg', g['value'], g['ms_per_step'])
PY
