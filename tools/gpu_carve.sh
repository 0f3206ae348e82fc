#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/gpu_check.py > gpurun_out/check.log 2>&1; echo "gpu_check rc=$?"; tail -1 gpurun_out/check.log
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(round(d['value'],1), round(d['ms_per_step'],2), 'pretok',round(s['pretok_ms'],2),'long',round(s['long_ms'],2),'probe',round(s['probe_ms'],2),'miss+sort',round(s['encode_ms']-s['probe_ms'],2),'gather',round(s['gather_ms'],2))"; }
echo base; run
for c in 25 50 75; do echo "miss carveout $c"; B200BPE_MISS_CARVEOUT=$c run; done
for c in 25 50; do echo "probe carveout $c"; B200BPE_PROBE_CARVEOUT=$c run; done
for w in config3 config5; do echo $w; run --workload $w --bytes 268435456; done
