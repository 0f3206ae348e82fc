#!/bin/bash
# round-end evidence of the final code: suite, smoke, the default bench line (all configs, api, strong), the reference arm,
# the launch list of the same command, one full ncu capture of the dominant kernels at the bench size
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2z_pytest.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log | cut -c1-200
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('   dev', round(d['value'],1), 'GB/s', round(d['ms_per_step'],3), 'ms', {k[:-3]: round(v,2) for k,v in s.items() if k not in ('h2d_ms','d2h_ms')})"; }

( time timeout 900 python bench.py ) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r2z_bench.err | cut -c1-200
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2z_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roofline", round(d["roofline"]["frac"],4), "pipeline", round(d["roofline"]["pipeline"]["frac"],4))
print("stage", {k: round(v,2) for k,v in d["stage_ms"].items()})
for k,v in (d.get("configs") or {}).items(): print(k, {x: (round(v[x],2) if isinstance(v.get(x),float) else v.get(x)) for x in ("parity","value","ms_per_step","error")}, "e2e", round((v.get("e2e") or {}).get("value",0),1))
print("api", {k: round(v.get("value",0),2) for k,v in (d.get("api") or {}).items() if isinstance(v,dict)})
print("strong", (d.get("strong") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2>/dev/null; echo "ref rc=$?"; tail -1 gpurun_out/r2z_bench_ref.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2z_launches.log 2>&1; echo "launches rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'pretok_kernel|probe_kernel|^miss_kernel|gather_kernel' -s 10 -c 5 -o gpurun_out/r2z_full_1GiB -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"
