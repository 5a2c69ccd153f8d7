#!/bin/bash
mkdir -p gpurun_out/c45
O=gpurun_out/c45
timeout 600 python -m pytest tests/test_degrade_gpu.py -x -q -m gpu -k "augment or usm" > $O/t_aug.log 2>&1; echo "exit $?" >> $O/t_aug.log
timeout 900 python tools/library_bar.py > $O/library_bar.log 2>&1
tail -n 3 $O/t_aug.log; cat $O/library_bar.log | grep "library bar"
