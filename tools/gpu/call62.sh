#!/bin/bash
mkdir -p gpurun_out/c62
O=gpurun_out/c62
timeout 600 python -m pytest tests/test_niqe_gpu.py -x -q -m gpu -s > $O/t.log 2>&1; echo "exit $?" >> $O/t.log
grep -v "^$" $O/t.log | tail -n 20
