#!/bin/bash
# A/B of the pre-tokeniser classification (old = HEAD, new = FMA-pipe byte tests), launch lists of configs 3 and 5, PCIe probe, chunk-size sweep of e2e
mkdir -p gpurun_out
bash tools/ab.sh run config2 config3:268435456 config5 2>&1 | tee gpurun_out/r2e_ab.txt
for w in config3 config5; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2e_launches_$w.csv \
  python bench.py --workload $w --bytes $([ $w = config3 ] && echo 268435456 || echo 0) --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2e_launches_$w.csv")) if len(r)>5 and r[0].isdigit()]
best=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0].replace("void ","")[:40]; v=float(r[-1])/1e3
    best[k]=max(best.get(k,0),v)
print("$w launch list (max us per kernel):")
for k,v in best.items():
    if v > 15: print(f"   {k:42s} {v:10.1f}")
PY
done
timeout 300 python tools/pcie_probe.py 2>&1 | grep -v "^$" | tail -12
for mb in 16 32 64 128; do
echo "chunk $mb MiB"; B200BPE_CHUNK_MB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('  e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],2), 'h2d_ms', round(e.get('h2d_ms',0),2), 'dev_ms', round(e.get('device_ms',0),2))"
done
