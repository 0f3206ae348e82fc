#!/bin/bash
mkdir -p gpurun_out
for w in config3 config5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mid_thread|long_piece|giant|find_long|miss_kernel' -c 50 --csv --log-file gpurun_out/mid_${w}.csv \
    python bench.py --workload $w --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${w}.log 2>&1
  echo "== $w rc=$?"
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/mid_${w}.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-12:]:
    print(r[4][:40], r[-1])
PY
done
