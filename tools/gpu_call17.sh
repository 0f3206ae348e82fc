#!/bin/bash
# ramped chunk schedule: parity suite, e2e / api numbers
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2j_pytest.log | tail -3
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs "$@" 2>gpurun_out/r2j_err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; a=d.get('api') or {}
print('   e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],2), 'h2d_ms', round(e.get('h2d_ms',0),1), '| pageable', round((a.get('packed_pageable') or {}).get('value',0),1), 'encode_batch', round((a.get('encode_batch_default_policy_pinned') or {}).get('value',0),1), 'list_str', round((a.get('list_str_to_numpy') or {}).get('value',0),2))" || tail -3 gpurun_out/r2j_err.log; }
echo "default"; run
echo "chunk cap 32"; B200BPE_CHUNK_MB=32 run --no-extras
echo "chunk cap 128"; B200BPE_CHUNK_MB=128 run --no-extras
echo "PACK=1"; B200BPE_PACK=1 run --no-extras
echo config4; run --workload config4 --no-extras
