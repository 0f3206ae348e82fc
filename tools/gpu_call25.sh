#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over the merge kernels' adversarial tests + a chunked host run; then the suite and the api lines
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "mid_piece or edge_shapes or golden" > gpurun_out/r2r_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|error" gpurun_out/r2r_memcheck.log | tail -4
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "mid_piece" > gpurun_out/r2r_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" gpurun_out/r2r_racecheck.log | tail -4
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2r_pytest.log | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs 2>gpurun_out/r2r_err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; a=d.get('api') or {}
print('   dev', round(d['value'],1), 'e2e', round(e['value'],1), '| api', {k: round(v.get('value',0),2) for k,v in a.items() if isinstance(v,dict)})" || tail -3 gpurun_out/r2r_err.log
