#!/bin/bash
# dev helper: parity check + mid-piece mode comparison on the GPU box; logs under gpurun_out/
mkdir -p gpurun_out
timeout 400 python tools/gpu_check.py > gpurun_out/check.log 2>&1; echo "gpu_check rc=$?"; tail -6 gpurun_out/check.log
for m in thread; do for w in config3 config5; do
  B200BPE_MID=$m timeout 300 python bench.py --workload $w --bytes 268435456 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${m}_${w}.log 2>&1
  echo "== $m $w rc=$?"; tail -3 gpurun_out/bench_${m}_${w}.log | cut -c1-1500
done; done
