#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_paths.py -q -m gpu -x -k "giant" > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest giant rc=$?"; tail -3 gpurun_out/r2c3_pytest.log
# knobs on the default build (config2, 1 GiB)
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k[:-3]: round(v,2) for k,v in s.items()})"; }
echo "default"; run
for c in 0 25 50 100; do echo "probe carveout $c"; B200BPE_PROBE_CARVEOUT=$c run; done
for c in 6 8 13 16; do echo "probe blocks $c"; B200BPE_PROBE_BLOCKS=$c run; done
echo "no L2 persist"; B200BPE_L2_PERSIST=0 run
for c in 25 50; do echo "miss carveout $c"; B200BPE_MISS_CARVEOUT=$c run; done
# full captures of the four big kernels at 256 MiB
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'pretok_kernel|probe_kernel|miss_kernel|gather_kernel' -s 8 -c 4 -o gpurun_out/r2c3_prof -f \
    python bench.py --bytes 268435456 --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2c3_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2c3_prof.ncu-rep
