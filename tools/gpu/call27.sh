#!/bin/bash
mkdir -p gpurun_out/c27
O=gpurun_out/c27
RESR_LIB_PATH=$PWD/build/variants/libresr_biastmem.so timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py -x -q -m gpu > $O/tests_bias.log 2>&1; echo "exit $?" >> $O/tests_bias.log
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -x -q -m gpu > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
for v in base bias base bias base bias; do
  echo "LIB=$v" >> $O/ab.log
  if [ $v = bias ]; then export RESR_LIB_PATH=$PWD/build/variants/libresr_biastmem.so; else unset RESR_LIB_PATH; fi
  timeout 300 python bench.py --no-train --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])
" >> $O/ab.log 2>&1
done
tail -n 3 $O/tests_bias.log $O/tests.log; cat $O/ab.log
