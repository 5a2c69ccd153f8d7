#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_degrade_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python tools/time_degrade.py 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_deg4.csv python tools/time_degrade.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_deg4.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki].split('(')[0][-44:]+" grid "+r[gi]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(',',''))/1e3
nb=max(a[0] for a in agg.values())//2 if agg else 1
tot=0; nk=0
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    if a[0] < 50: continue
    print(f"{a[1]/a[0]:8.1f} us avg  x{a[0]:4d}  {k}");
PY
