#!/bin/bash
# round-2 call 2: rewritten engine (TMA-staged probe, dense miss records, piece-parallel gather, cluster kernel, device special scan)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2c2_smoke.log | cut -c1-400
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2c2_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2c2_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c2_bench.json").read().strip().splitlines()[-1])
    print("value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", d.get("e2e",{}).get("value"))
    print("stage", d.get("stage_ms"))
    print("api", d.get("api"))
    print("strong", d.get("strong"))
    for k,v in (d.get("configs") or {}).items(): print(k, {x: v.get(x) for x in ("parity","value","ms_per_step","error")}, v.get("stage_ms"), v.get("e2e"))
    print("cpu", d.get("cpu_baseline"))
except Exception as e: print("bench parse failed", e)
PY
