#!/bin/bash
mkdir -p gpurun_out
w=${1:-config3}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'mid_thread' -s 4 -c 4 -o gpurun_out/mid_full_${w} -f \
    python bench.py --workload $w --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${w}.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/mid_full_${w}.ncu-rep --page details --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats 2>&1 | grep -v "^\s*$" | cut -c1-160 | head -150
