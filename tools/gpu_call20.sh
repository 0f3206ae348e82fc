#!/bin/bash
# 256-bit table loads: parity suite, stage times of every config
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2p_pytest.log | tail -3
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras "$@" 2>gpurun_out/r2p_err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; s=d['stage_ms']
print('   dev', round(d['value'],1), 'GB/s', round(d['ms_per_step'],2), 'ms | e2e', round(e['value'],1), {k[:-3]: round(v,2) for k,v in s.items() if k not in ('h2d_ms','d2h_ms')})" || tail -3 gpurun_out/r2p_err.log; }
echo config2; run
echo config3; run --workload config3
echo config4; run --workload config4
echo config5; run --workload config5
