#!/bin/bash
# per-class pmerge kernels with class-sized buffers: parity subset, then the class split sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu > gpurun_out/r2u_pytest.log 2>&1; echo "pytest (default) rc=$? $(tail -1 gpurun_out/r2u_pytest.log)"
for mc in 2 1; do B200BPE_PMERGE_MIN_CLS=$mc timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "mid_piece or golden or seeded or edge" > gpurun_out/r2u_pytest$mc.log 2>&1; echo "pytest MIN_CLS=$mc rc=$? $(tail -1 gpurun_out/r2u_pytest$mc.log)"; done
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('   dev', round(d['value'],1), 'GB/s', round(d['ms_per_step'],3), 'ms', {k[:-3]: round(v,2) for k,v in s.items() if k in ('long_ms','pretok_ms','device_total_ms')})"; }
for mc in 3 2 1; do echo "MIN_CLS=$mc"; for w in "config3 --bytes 268435456" config5; do echo -n " $w"; B200BPE_PMERGE_MIN_CLS=$mc run --workload $w; done; done
