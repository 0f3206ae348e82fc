#!/bin/bash
# dev helper: GPU test suite + smoke + a pretok profile on the mixed-script workload; logs under gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'pretok_kernel' -s 2 -c 1 -o gpurun_out/pretok_o200k_config3 -f \
    python bench.py --workload config3 --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pretok.log 2>&1
echo "ncu rc=$?"
