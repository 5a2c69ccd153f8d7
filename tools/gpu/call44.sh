#!/bin/bash
mkdir -p gpurun_out/c44
O=gpurun_out/c44
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "wgrad_nhwc or 13" > $O/san_memcheck_wgrad_mn.log 2>&1; echo "exit $?" >> $O/san_memcheck_wgrad_mn.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "wgrad_nhwc and (2-12 or 1-7 or 3-5 or 1-3)" > $O/san_racecheck_wgrad_mn.log 2>&1; echo "exit $?" >> $O/san_racecheck_wgrad_mn.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gan_losses_gpu.py -x -q -m gpu -k "usm_backward and not 256" > $O/san_memcheck_usm_bwd.log 2>&1; echo "exit $?" >> $O/san_memcheck_usm_bwd.log
timeout 600 python -m pytest tests/test_gan_losses_gpu.py tests/test_generator_gpu.py -x -q -m gpu -s -k "usm or bf16_recipe" > $O/t_fix.log 2>&1; echo "exit $?" >> $O/t_fix.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -k regex:wgrad_mn --launch-skip 160 -c 24 --csv --log-file $O/wgrad_mn_launches.csv python tools/ncu_targets.py train > $O/ncu_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_mn_kernel --launch-skip 100 -c 2 -o $O/wgrad_mn_full python tools/ncu_targets.py train > $O/ncu_train_full.log 2>&1
for f in $O/san_*.log $O/t_fix.log; do echo "== $f"; tail -n 4 $f; done
tail -n 8 $O/wgrad_mn_launches.csv
