#!/bin/bash
# Round-2 starter (run under gpurun): A/B of the shipped library against one built with the next-round pre-tokeniser
# rules (B2_O200K_FAST_PREFIX, B2_O200K_FAST_APOS, B2_CL100K_FAST_CONTRACTION, B2_R50K_FAST_CONTRACTION, B2_CL100K_FAST_WSNL, B2_O200K_FAST_WSNL -- CPU-verified,
# see tests/test_pretok_rules.py).  Build the variant HERE first (nvcc cross-compiles without a GPU):
#   cd tiktoken_b200/csrc && nvcc -DB2_O200K_FAST_PREFIX=1 -DB2_O200K_FAST_APOS=1 -DB2_CL100K_FAST_CONTRACTION=1 \
#      -DB2_R50K_FAST_CONTRACTION=1 -DB2_CL100K_FAST_WSNL=1 -DB2_O200K_FAST_WSNL=1 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared \
#      -Xcompiler -fPIC -o libvariant_b.so b200bpe.cu
# then: gpurun --timeout 1500 -- 'bash tools/gpu_next_rules.sh'.  Parity first (gpu_check + pytest -m gpu on the variant),
# then the four workloads for both builds.  Adopt the flags in tiktoken_b200/_lib.py:NVCC_FLAGS only if parity is green.
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(round(d['value'],1), round(d['ms_per_step'],2), 'pretok',round(s['pretok_ms'],2),'long',round(s['long_ms'],2),'probe',round(s['probe_ms'],2),'miss+sort',round(s['encode_ms']-s['probe_ms'],2),'gather',round(s['gather_ms'],2))"; }
cp tiktoken_b200/csrc/libb200bpe.so /tmp/a.so
for v in a b; do
  [ $v = b ] && cp tiktoken_b200/csrc/libvariant_b.so tiktoken_b200/csrc/libb200bpe.so
  echo "== variant $v"
  timeout 400 python tools/gpu_check.py > gpurun_out/check_$v.log 2>&1; echo "gpu_check rc=$?"; tail -1 gpurun_out/check_$v.log
  timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$v.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_$v.log
  echo config2; run
  echo config3; run --workload config3 --bytes 268435456
  echo config5; run --workload config5 --bytes 268435456
done
cp /tmp/a.so tiktoken_b200/csrc/libb200bpe.so
