#!/bin/bash
mkdir -p gpurun_out/c63
O=gpurun_out/c63
timeout 1200 python -m pytest tests/test_degrade_gpu.py tests/test_gan_losses_gpu.py tests/test_parity_at_size_gpu.py -x -q -m gpu -k "not cfg3 and not cfg4 and not cfg5 and not pixel_loss" > $O/t_deg.log 2>&1; echo "exit $?" >> $O/t_deg.log
tail -n 3 $O/t_deg.log
for i in 1 2; do for v in 1 0; do
  echo "RESR_USM_FUSED=$v" >> $O/ab.log
  RESR_USM_FUSED=$v timeout 300 python tools/time_degrade.py 2>&1 | tail -n 1 >> $O/ab.log
  RESR_USM_FUSED=$v timeout 300 python tools/time_degrade_ops.py 2>&1 | grep -i "usm" >> $O/ab.log
done; done
cat $O/ab.log
