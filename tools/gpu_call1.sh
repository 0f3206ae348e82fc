#!/bin/bash
# round-2 call 1: the new path tests on the round-1 kernels + all BASELINE configs at their stated sizes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2c1_pytest.log
for w in config2 config3 config4 config5; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2c1_bench_$w.err | tail -1 > gpurun_out/r2c1_bench_$w.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c1_bench_$w.json").read()); s=d['stage_ms']
    print("$w", round(d['value'],1),"GB/s", round(d['ms_per_step'],2),"ms e2e", round(d['e2e']['value'],1), {k: round(v,2) for k,v in s.items()})
except Exception as e: print("$w failed", e)
PY
done
