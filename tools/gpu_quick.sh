#!/bin/bash
# dev helper: parity suite + stage times of configs 3, 5, 2
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/q_pytest.log | tail -3
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('   dev', round(d['value'],1), 'GB/s', round(d['ms_per_step'],3), 'ms', {k[:-3]: round(v,2) for k,v in s.items() if k not in ('h2d_ms','d2h_ms')})"; }
for w in "config3 --bytes 268435456" config3 config5 config2; do echo -n " $w"; run --workload $w; done
