#!/bin/bash
mkdir -p gpurun_out/c65
O=gpurun_out/c65
timeout 1200 python -m pytest tests/test_degrade_gpu.py tests/test_parity_at_size_gpu.py -x -q -m gpu -k "not cfg3 and not cfg4 and not cfg5" > $O/t_deg.log 2>&1; echo "exit $?" >> $O/t_deg.log
tail -n 3 $O/t_deg.log
timeout 300 python tools/time_degrade.py 2>&1 | tail -n 1
timeout 300 python tools/time_degrade_ops.py 2>&1 | tail -n 13
timeout 300 python bench.py --no-train --no-tiled --no-cpu --steps 5 --no-other-precision 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])['degradation']
print('bench degradation', round(d['value']), 'pairs/s', round(d['ms_per_step']*1e3,1), 'us; e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8_images']['value']), '; batch 256', round(d['large_batch']['value']))
"
