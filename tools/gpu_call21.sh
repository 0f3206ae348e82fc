#!/bin/bash
# ncu --set full of the main kernels (config 2, 256 MiB) and of the long-piece kernels (config 3, 256 MiB)
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'pretok_kernel|probe_kernel|^miss_kernel|gather_kernel' -s 10 -c 5 -o gpurun_out/r2n_full -f \
    python bench.py --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2n_ncu.log 2>&1
echo "ncu rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'pmerge_kernel|mid_group32_kernel|mid_group16_kernel|pretok_slow_kernel|pmerge_long' -s 5 -c 5 -o gpurun_out/r2n_long -f \
    python bench.py --workload config3 --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2n_ncu2.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
