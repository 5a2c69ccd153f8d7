#!/bin/bash
mkdir -p gpurun_out/c26
O=gpurun_out/c26
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -x -q -m gpu > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
RESR_CONV_DIRECT_STORE=1 timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -x -q -m gpu > $O/tests_direct1.log 2>&1; echo "exit $?" >> $O/tests_direct1.log
for v in 0 2 1 0 2 1; do
  echo "DIRECT=$v" >> $O/ab.log
  RESR_CONV_DIRECT_STORE=$v timeout 300 python bench.py --no-train --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])
" >> $O/ab.log 2>&1
done
RESR_CONV_DIRECT_STORE=1 RESR_LIB_PATH=$PWD/build/variants/libresr_prof.so timeout 300 python tools/wait_profile.py > $O/wait_direct.txt 2>&1
tail -n 3 $O/tests.log $O/tests_direct1.log; cat $O/ab.log; cat $O/wait_direct.txt | cut -c1-330
