#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/gpu_check.py > gpurun_out/check.log 2>&1; echo "gpu_check rc=$?"; tail -1 gpurun_out/check.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config2.log 2>&1; echo "== config2 rc=$?"; tail -1 gpurun_out/bench_config2.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['value'])"
for w in config3 config5 config4; do
  timeout 300 python bench.py --workload $w --bytes 268435456 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}.log 2>&1
  echo "== $w rc=$?"; tail -1 gpurun_out/bench_${w}.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['value'])"
done
bash tools/gpu_mid2.sh > gpurun_out/mid2.log 2>&1
python - <<PY
import csv
for w in ("config3","config5"):
    rows=[r for r in csv.reader(open(f"gpurun_out/mid_{w}.csv")) if len(r)>5 and r[0].isdigit()]
    for k in range(0,min(len(rows),10),5):
        print(w, [ (r[4][:18].replace("void ",""), round(float(r[-1])/1e6,3)) for r in rows[k:k+5]])
PY
