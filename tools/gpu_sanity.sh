#!/bin/bash
# dev helper: last check of a round -- suite, smoke, the default bench line, ncu of the long-piece kernels on config 3
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/san_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/san_pytest.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/san_bench.json 2> gpurun_out/san_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/san_bench.err | cut -c1-200
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/san_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roofline", round(d["roofline"]["frac"],4), "traffic", d["roofline"]["traffic"])
for k,v in (d.get("configs") or {}).items(): print(k, {x: (round(v[x],2) if isinstance(v.get(x),float) else v.get(x)) for x in ("parity","value","ms_per_step","error")})
print("api", {k: (round(v.get("value",0),2) if "value" in v else v) for k,v in (d.get("api") or {}).items() if isinstance(v,dict)})
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'pmerge_kernel|mid_group32_kernel|mid_group16_kernel|pretok_slow_kernel|pmerge_long|pretok_kernel' -s 6 -c 6 -o gpurun_out/san_long -f \
    python bench.py --workload config3 --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > gpurun_out/san_ncu.log 2>&1; echo "ncu rc=$?"
