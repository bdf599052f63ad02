#!/bin/bash
# every launch of two C2 solves with its device time (cold-cache, serialised: compare SHARES): profiles/r2_launches.md source
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-side-configs > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r)>14 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0][:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[14].replace(",",""))/1e3
tot=sum(a[1] for a in agg.values())
print("launches %d total %.1f us"%(len(rows),tot))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-62s n %5d  %10.1f us  %5.1f %%"%(k,a[0],a[1],100*a[1]/tot))
PY
