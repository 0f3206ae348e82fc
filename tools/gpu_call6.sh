#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2c6_pytest.log
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k[:-3]: round(v,2) for k,v in s.items()})"; }
echo config2; run
echo config3; run --workload config3
echo config4; run --workload config4
echo config5; run --workload config5
echo "config2 copy threads"; for t in 2 4 16; do B200BPE_COPY_THREADS=$t timeout 300 python - <<PY
import sys, time, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import vocab_util as vu, tiktoken_b200
from tools import corpus
pat, ranks, sp, _ = vu.load_encoding("cl100k_base")
e = tiktoken_b200.Encoding("x", pat_str=pat, mergeable_ranks=ranks, special_tokens=sp)
text, off = corpus.config2(512<<20, 5)
for _ in range(2): e.encode_ordinary_packed(text, off).close()
ts=[]
for _ in range(3):
    t0=time.perf_counter(); e.encode_ordinary_packed(text, off).close(); ts.append(time.perf_counter()-t0)
print("  threads $t pageable e2e GB/s", round(len(text)/min(ts)/1e9,1))
PY
done
for w in config3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2c6_launches_$w.csv \
  python bench.py --workload $w --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2c6_launches_$w.csv")) if len(r)>5 and r[0].isdigit()]
best=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0].replace("void ","")[:40]; v=float(r[-1])/1e3
    best[k]=max(best.get(k,0),v)
print("$w launch list (max us per kernel):")
for k,v in best.items(): print(f"   {k:42s} {v:10.1f}")
PY
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'mid_group32_kernel' -s 2 -c 1 -o gpurun_out/r2c6_prof_c3 -f \
    python bench.py --workload config3 --bytes 268435456 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2c6_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
