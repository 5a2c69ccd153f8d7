#!/bin/bash
# Round-2 GPU call 1: microbenchmarks that decide the HP1 plan, FMA peak, new at-size parity tests, sanitizer pass.
mkdir -p gpurun_out/c1
O=gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
for n in 48 96 192; do timeout 60 ./tools/mma_pair check $n; done > $O/pair_check.txt 2>&1
( timeout 60 ./tools/mma_power 96 3; timeout 60 ./tools/mma_power 192 3; timeout 60 ./tools/mma_pair power 96 3; timeout 60 ./tools/mma_pair power 192 3 ) > $O/pair_power.txt 2>&1
timeout 60 ./tools/fma_peak 3 > $O/fma_peak.json 2>&1
timeout 900 python -m pytest tests/test_parity_at_size_gpu.py -x -q -m gpu -s > $O/parity_at_size.log 2>&1
echo "parity exit $?" >> $O/parity_at_size.log
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_parity_at_size_gpu.py > $O/gpu_tests.log 2>&1
echo "tests exit $?" >> $O/gpu_tests.log
# sanitizer: small-shape conv / degrade / train tests (memcheck, then racecheck on the shared-memory heavy ones)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "test_conv_epilogues or test_conv_16bit or test_conv_rgb" > $O/san_memcheck_conv.log 2>&1
echo "exit $?" >> $O/san_memcheck_conv.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_degrade_gpu.py -x -q -m gpu -k "golden or oracle or end_to_end" > $O/san_memcheck_degrade.log 2>&1
echo "exit $?" >> $O/san_memcheck_degrade.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_degrade_gpu.py -x -q -m gpu -k "filter2d_golden or usm_golden or jpeg_golden" > $O/san_racecheck_degrade.log 2>&1
echo "exit $?" >> $O/san_racecheck_degrade.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "test_conv_16bit" > $O/san_racecheck_conv.log 2>&1
echo "exit $?" >> $O/san_racecheck_conv.log
tail -3 $O/*.log $O/*.txt $O/*.json
