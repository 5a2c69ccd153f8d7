#!/bin/bash
mkdir -p gpurun_out/c28
O=gpurun_out/c28
timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
for v in 1 2 3; do
  timeout 300 python bench.py --no-train --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])
" >> $O/ab.log 2>&1
done
tail -n 3 $O/tests.log; cat $O/ab.log
