#!/bin/bash
# dev helper: the round-end checks on one GPU box (tests, smoke, the bench line, an ncu launch list and one
# full capture of the mid-piece kernel on the mixed-script workload); logs under gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.json | cut -c1-600
for w in config3 config4 config5; do timeout 300 python bench.py --workload $w --bytes 268435456 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_$w.json; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.json | cut -c1-300
bash tools/gpu_launches.sh > gpurun_out/launch_summary.txt 2>&1; tail -20 gpurun_out/launch_summary.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'mid_thread_kernel' -s 2 -c 1 -o gpurun_out/mid_thread_config3 -f \
    python bench.py --workload config3 --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_mid.log 2>&1
echo "ncu rc=$?"
