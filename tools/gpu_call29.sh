#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh run config3:268435456 config5 2>&1 | tee gpurun_out/r2t_ab.txt
for so in tiktoken_b200/csrc/variants/*.so; do echo "$(basename $so) MIN_CLS=2"; B200BPE_LIB=$PWD/$so B200BPE_PMERGE_MIN_CLS=2 timeout 300 python bench.py --workload config3 --bytes 268435456 --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('   dev', round(d['value'],1), 'GB/s', round(d['ms_per_step'],3), 'ms', {k[:-3]: round(v,2) for k,v in s.items() if k in ('long_ms','pretok_ms')})"; done
