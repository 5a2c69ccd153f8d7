#!/bin/bash
mkdir -p gpurun_out/c61
O=gpurun_out/c61
timeout 2400 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1; echo "exit $?" >> $O/t_all.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_degrade_gpu.py -x -q -m gpu -k "golden or oracle or end_to_end or augment or u8_images" > $O/san_memcheck_degrade_pdl.log 2>&1; echo "exit $?" >> $O/san_memcheck_degrade_pdl.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "oracle_autograd and bf16 and 0-shape0 or buckets" > $O/san_memcheck_train_bf16.log 2>&1; echo "exit $?" >> $O/san_memcheck_train_bf16.log
tail -n 3 $O/t_all.log; for f in $O/san_*.log; do echo "== $f"; tail -n 4 $f; done
