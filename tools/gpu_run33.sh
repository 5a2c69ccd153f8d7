#!/bin/bash
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
echo -n "v9: "; RESR_LIB_PATH=$PWD/build/variants/libresr_v9.so timeout 120 python tools/power_probe.py 3 2>&1 | tail -1
echo -n "v9 nepi3: "; RESR_CONV_NEPI=3 RESR_LIB_PATH=$PWD/build/variants/libresr_v9.so timeout 120 python tools/power_probe.py 3 2>&1 | tail -1
echo -n "v6: "; RESR_LIB_PATH=$PWD/build/variants/libresr_v6.so timeout 120 python tools/power_probe.py 3 2>&1 | tail -1
done
