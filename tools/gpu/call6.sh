#!/bin/bash
mkdir -p gpurun_out/c6
O=gpurun_out/c6
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py tests/test_parity_at_size_gpu.py -x -q -m gpu -s > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
for v in "1 1" "1 2" "0 2" "1 2"; do set -- $v
  echo "PAIR=$1 EPI_BUFS=$2" >> $O/ab.log
  RESR_CONV_PAIR=$1 RESR_CONV_EPI_BUFS=$2 timeout 300 python bench.py --no-degrade --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], 'train', d['training'].get('ms_per_step'), d['training'].get('value'), d['training'].get('error'))
except Exception as e:
    print('parse error', e)
" >> $O/ab.log 2>&1
done
grep -n "passed\|failed\|exit\|cfg4" $O/tests.log | tail -n 8; cat $O/ab.log
