#!/bin/bash
# round-2 re-entry: full GPU suite, smoke, the default bench line (all configs), the reference arm, launch list, full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2d_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2d_smoke.log | cut -c1-300
( time timeout 900 python bench.py ) > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2d_bench.err; grep '^{' gpurun_out/r2d_bench.json | tail -1 | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2d_bench_ref.json 2>/dev/null; echo "ref rc=$?"; tail -1 gpurun_out/r2d_bench_ref.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2d_launches.log 2>&1
echo "launches rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'pretok_kernel|probe_kernel|^miss_kernel|gather_kernel' -c 25 -o gpurun_out/r2d_full -f \
    python bench.py --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > gpurun_out/r2d_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2d_ncu.log
