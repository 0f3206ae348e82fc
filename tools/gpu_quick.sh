#!/bin/bash
# dev helper: parity check + GPU tests + the four workloads (device-resident stage timings)
mkdir -p gpurun_out
timeout 400 python tools/gpu_check.py > gpurun_out/check.log 2>&1; echo "gpu_check rc=$?"; tail -1 gpurun_out/check.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'pretok',round(s['pretok_ms'],2),'long',round(s['long_ms'],2),'probe',round(s['probe_ms'],2),'miss+sort',round(s['encode_ms']-s['probe_ms'],2),'gather',round(s['gather_ms'],2))"; }
echo config2; run
for w in config3 config4 config5; do echo $w; run --workload $w --bytes 268435456; done
