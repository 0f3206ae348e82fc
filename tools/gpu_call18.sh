#!/bin/bash
# upload stream per slot + SWAR whitespace scan: parity suite, e2e / api numbers, config3, stream-overlap experiment
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2k_pytest.log | tail -3
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs "$@" 2>gpurun_out/r2k_err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; a=d.get('api') or {}; s=d['stage_ms']
print('   dev', round(d['value'],1), 'GB/s', round(d['ms_per_step'],2), 'ms | e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],2), 'h2d_ms', round(e.get('h2d_ms',0),1), '| pageable', round((a.get('packed_pageable') or {}).get('value',0),1), 'encode_batch', round((a.get('encode_batch_default_policy_pinned') or {}).get('value',0),1), 'list_str', round((a.get('list_str_to_numpy') or {}).get('value',0),2), {k[:-3]: round(v,2) for k,v in s.items() if k in ('pretok_ms','long_ms','probe_ms','gather_ms')})" || tail -3 gpurun_out/r2k_err.log; }
echo "default"; run
echo "chunk 32"; B200BPE_CHUNK_MB=32 run --no-extras
echo "chunk 16"; B200BPE_CHUNK_MB=16 run --no-extras
echo "PACK=1"; B200BPE_PACK=1 run --no-extras
echo config3; run --workload config3 --no-extras
echo config4; run --workload config4 --no-extras
echo config5; run --workload config5 --no-extras
timeout 600 python tools/overlap_probe.py 2>&1 | tail -6
