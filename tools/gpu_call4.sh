#!/bin/bash
mkdir -p gpurun_out
# 1. the whole GPU suite on the default build (mid_group kernel, gather v3, special scan, async queue, ...)
timeout 1500 python -m pytest tests -q -m gpu -x --durations=6 > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r2c4_pytest.log
# 2. flag variants: parity subset + timings on config2 (1 GiB), config3 (256 MiB), config5 (64 MiB), config4 (256 MiB)
bash tools/ab.sh run config2 config3:268435456 config5 config4:268435456 2>&1 | tee gpurun_out/r2c4_ab.txt
# 3. old mid kernel for comparison
echo "== mid_thread (old) kernel"; export B200BPE_LIB=$PWD/tiktoken_b200/csrc/variants/libb200bpe_base.so
for spec in config3:268435456 config5:0; do w=${spec%%:*}; nb=${spec#*:}
B200BPE_MID_GROUP=0 timeout 300 python bench.py --workload $w --bytes $nb --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('   $w', round(d['value'],1), round(d['ms_per_step'],2), {k[:-3]: round(v,2) for k,v in s.items()})"; done
