#!/bin/bash
mkdir -p gpurun_out/c5
O=gpurun_out/c5
RESR_LIB_PATH=$PWD/build/variants/libresr_prof.so timeout 300 python tools/wait_profile.py > $O/wait_n64.txt 2>&1
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -x -q -m gpu > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
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
cat $O/wait_n64.txt; tail -n 3 $O/tests.log; cat $O/ab.log
