#!/bin/bash
mkdir -p gpurun_out/c53
O=gpurun_out/c53
for i in 1 2; do
for v in "" "RESR_NO_PDL=1"; do
  echo "variant: ${v:-PDL}" >> $O/ab.log
  env $v timeout 300 python tools/time_degrade.py 2>&1 | tail -n 1 >> $O/ab.log
  env $v timeout 300 python bench.py --no-train --no-tiled --no-cpu --steps 5 --no-other-precision 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])['degradation']
print('bench degradation', round(d['value']), 'pairs/s', round(d['ms_per_step']*1e3,1), 'us; e2e', round(d['e2e']['value']), '; batch 256', round(d['large_batch']['value']))
" >> $O/ab.log
done; done
cat $O/ab.log
