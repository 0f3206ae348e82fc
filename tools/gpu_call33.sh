#!/bin/bash
# pipeline slots per device (uploads N-2 chunks ahead): e2e and api lines, chunk sizes 64 and 32
mkdir -p gpurun_out
run() { timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs "$@" 2>gpurun_out/r2w_err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; a=d.get('api') or {}
print('   dev', round(d['value'],1), 'e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],2), '| pageable', round((a.get('packed_pageable') or {}).get('value',0),1), 'encode_batch', round((a.get('encode_batch_default_policy_pinned') or {}).get('value',0),1))" || tail -3 gpurun_out/r2w_err.log; }
for so in tiktoken_b200/csrc/variants/*.so; do echo "$(basename $so)"; B200BPE_LIB=$PWD/$so run; B200BPE_LIB=$PWD/$so B200BPE_CHUNK_MB=32 run --no-extras; done
B200BPE_LIB=$PWD/tiktoken_b200/csrc/variants/libb200bpe_s4.so timeout 600 python -m pytest tests/test_gpu_paths.py -q -x -m gpu -k "chunk or 200mib or packed" 2>&1 | tail -1
