#!/bin/bash
# occupancy knob sweeps after the 256-bit loads: probe blocks per SM (env), miss kernel min blocks (builds), host chunk size
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('   dev', round(d['value'],1), 'GB/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value'],1), {k[:-3]: round(v,2) for k,v in s.items() if k in ('probe_ms','encode_ms','gather_ms','pretok_ms')})"; }
for pb in 8 10 12 14 16; do echo -n "probe blocks $pb"; B200BPE_PROBE_BLOCKS=$pb run; done
for so in tiktoken_b200/csrc/variants/*.so; do echo -n "$(basename $so)"; B200BPE_LIB=$PWD/$so run; done
for mb in 48 96; do echo -n "chunk $mb"; B200BPE_CHUNK_MB=$mb run; done
