#!/bin/bash
# dev helper: per-kernel durations (ncu launch list) of the default 1 GiB bench step
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline ${@} > gpurun_out/launches.log 2>&1
echo rc=$?
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r)>5 and r[0].isdigit()]
best=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0].replace("void ","")[:40]; v=float(r[-1])/1e3
    best[k]=max(best.get(k,0),v)
for k,v in best.items(): print(f"{k:42s} max_us={v:10.1f}")
PY
