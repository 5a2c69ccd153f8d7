#!/bin/bash
# pair-kernel bring-up: conv / generator parity, then same-box A/B of the cfg3 forward
mkdir -p gpurun_out/c3
O=gpurun_out/c3
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > $O/conv.log 2>&1; echo "exit $?" >> $O/conv.log
timeout 900 python -m pytest tests/test_generator_gpu.py tests/test_parity_at_size_gpu.py -x -q -m gpu -s > $O/gen.log 2>&1; echo "exit $?" >> $O/gen.log
for v in "0 1" "1 0" "1 1" "0 1" "1 1"; do set -- $v
  echo "PAIR=$1 N64=$2" >> $O/ab.log
  RESR_CONV_PAIR=$1 RESR_CONV_PAIR_N64=$2 timeout 300 python bench.py --no-train --no-degrade --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['e2e']['value'])
except Exception as e:
    print('parse error', e)
" >> $O/ab.log 2>&1
done
tail -4 $O/conv.log $O/gen.log; cat $O/ab.log
