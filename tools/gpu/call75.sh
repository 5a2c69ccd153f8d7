#!/bin/bash
mkdir -p gpurun_out/c75
O=gpurun_out/c75
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
# second eager bf16 training step: skip the first step's launches (~1,060 incl. packs), capture the next 1,100
timeout 1500 ncu --metrics $M --clock-control none --launch-skip 1075 -c 1060 --csv --log-file $O/train_launches.csv python tools/ncu_targets.py train > $O/ncu_train.log 2>&1
tail -n 2 $O/ncu_train.log
python - <<'PY'
import csv, io, collections
lines=open('gpurun_out/c75/train_launches.csv').read().splitlines()
start=next(i for i,l in enumerate(lines) if l.startswith('"ID"'))
by=collections.OrderedDict()
for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    d=by.setdefault(int(r["ID"]),{"name":r["Kernel Name"].split("(")[0]}); d[r["Metric Name"]]=float(r["Metric Value"].replace(",",""))
agg=collections.OrderedDict()
for d in by.values():
    a=agg.setdefault(d['name'][-40:],[0,0.0,0.0,0.0])
    t=d.get('gpu__time_duration.sum',0)
    a[0]+=1; a[1]+=t; a[2]+=d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',0)*t; a[3]+=d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0)
T=sum(a[1] for a in agg.values())
print('launches', len(by), 'sum ms', T/1e6)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:42s} n={a[0]:4d} total {a[1]/1e6:7.3f} ms  mean {a[1]/a[0]/1e3:6.1f} us  tensor pipe {a[2]/max(a[1],1):5.1f} %  dram {a[3]/1e9:6.2f} GB")
PY
