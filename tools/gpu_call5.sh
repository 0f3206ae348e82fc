#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c5_pytest.log
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k[:-3]: round(v,2) for k,v in s.items()})"; }
echo config2; run
echo config3; run --workload config3
echo config4; run --workload config4
echo config5; run --workload config5
# launch list of config5 (which long-piece kernel takes the time) and config3
for w in config5 config3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2c5_launches_$w.csv \
  python bench.py --workload $w --bytes $([ $w = config3 ] && echo 268435456 || echo 0) --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2c5_launches_$w.csv")) if len(r)>5 and r[0].isdigit()]
best=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0].replace("void ","")[:40]; v=float(r[-1])/1e3
    best[k]=max(best.get(k,0),v)
print("$w launch list (max us per kernel):")
for k,v in best.items(): print(f"   {k:42s} {v:10.1f}")
PY
done
# full captures: mid_group + pretok on config3 (256 MiB), gather/probe/miss on config2 (256 MiB)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'mid_group_kernel|pretok_kernel' -s 4 -c 2 -o gpurun_out/r2c5_prof_c3 -f \
    python bench.py --workload config3 --bytes 268435456 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2c5_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'probe_kernel|miss_kernel|gather_kernel' -s 8 -c 4 -o gpurun_out/r2c5_prof_c2 -f \
    python bench.py --bytes 268435456 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2c5_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
