#!/bin/bash
mkdir -p gpurun_out/c17
O=gpurun_out/c17
timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
# sanitizer: pair conv kernel (forced on small shapes), new degradation kernels, u8 path
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "pairs and (epilogues or 16bit or rgb or two_epilogue)" > $O/san_memcheck_pairconv.log 2>&1; echo "exit $?" >> $O/san_memcheck_pairconv.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_generator_gpu.py -x -q -m gpu -k "awkward and pairs or u8_image or bf16" > $O/san_memcheck_gen.log 2>&1; echo "exit $?" >> $O/san_memcheck_gen.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_degrade_gpu.py -x -q -m gpu -k "golden or oracle or end_to_end or device_side" > $O/san_memcheck_degrade.log 2>&1; echo "exit $?" >> $O/san_memcheck_degrade.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_degrade_gpu.py -x -q -m gpu -k "filter2d_golden or usm_golden or jpeg_golden" > $O/san_racecheck_degrade.log 2>&1; echo "exit $?" >> $O/san_racecheck_degrade.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "pairs and (16bit or epilogues)" > $O/san_racecheck_pairconv.log 2>&1; echo "exit $?" >> $O/san_racecheck_pairconv.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "pairs and 16bit" > $O/san_synccheck_pairconv.log 2>&1; echo "exit $?" >> $O/san_synccheck_pairconv.log
# ncu evidence
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --launch-skip 353 -c 352 --csv --log-file $O/gen_launches.csv python tools/ncu_targets.py gen > $O/ncu_gen.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair --launch-skip 357 -c 5 -o $O/rdb_full python tools/ncu_targets.py rdb > $O/ncu_rdb.log 2>&1
M2=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed.sum
timeout 600 ncu --metrics $M2 --clock-control none --csv --log-file $O/deg_launches.csv python tools/ncu_targets.py degrade > $O/ncu_deg.log 2>&1
for f in $O/tests.log $O/san_*.log; do echo "== $f"; tail -n 4 $f; done
