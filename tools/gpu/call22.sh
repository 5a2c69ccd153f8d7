#!/bin/bash
mkdir -p gpurun_out/c22
O=gpurun_out/c22
for v in 2 -1 1 0 2 -1; do
  echo "PROMO=$v" >> $O/ab.log
  if [ $v = -1 ]; then unset RESR_TMAP_PROMO; else export RESR_TMAP_PROMO=$v; fi
  timeout 300 python bench.py --no-train --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])
" >> $O/ab.log 2>&1
done
unset RESR_TMAP_PROMO
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --launch-skip 365 -c 10 --csv --log-file $O/rdb_launches.csv python tools/ncu_targets.py gen > $O/ncu.log 2>&1
cat $O/ab.log; python tools/ncu_table.py $O/rdb_launches.csv 10
