#!/bin/bash
# hybrid: parallel merge for the long classes, group-of-lanes for the short ones -- parity, then class-split sweep
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras "$@" 2>gpurun_out/r2h_err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('  ', round(d['value'],1), 'GB/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), {k[:-3]: round(v,2) for k,v in s.items()})" || tail -3 gpurun_out/r2h_err.log; }
for mc in 3 2 1 0 off; do
  if [ $mc = off ]; then export B200BPE_PMERGE=0; else export B200BPE_PMERGE=1 B200BPE_PMERGE_MIN_CLS=$mc; fi
  echo "== PMERGE_MIN_CLS=$mc"
  echo -n config3; run --workload config3 --bytes 268435456
  echo -n config5; run --workload config5
done
unset B200BPE_PMERGE B200BPE_PMERGE_MIN_CLS
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2h_launches_config3.csv \
  python bench.py --workload config3 --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2h_launches_config3.csv")) if len(r)>5 and r[0].isdigit()]
best=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0].replace("void ","")[:40]; v=float(r[-1])/1e3
    best[k]=max(best.get(k,0),v)
print("config3 launch list (max us per kernel):")
for k,v in best.items():
    if v > 15: print(f"   {k:42s} {v:10.1f}")
PY
