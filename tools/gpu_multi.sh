#!/bin/bash
# dev helper, N-GPU box (gpurun --gpus N -- 'bash tools/gpu_multi.sh N'): the one-process multi-GPU engine test, then the
# torchrun bench at N ranks (NCCL count exchange per step, strong-scaling line, one-process engine over N devices)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_paths.py -q -m gpu -x -k "multi_gpu" > gpurun_out/multi_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/multi_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --no-configs > gpurun_out/multi_bench$N.json 2> gpurun_out/multi_bench$N.err; echo "bench rc=$?"; tail -3 gpurun_out/multi_bench$N.err | cut -c1-200
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/multi_bench$N.json").read().strip().splitlines()[-1])
    print("value", round(d.get("value"),1), "ms", round(d.get("ms_per_step"),3), "e2e", round(d.get("e2e",{}).get("value"),1), "per rank", [round(x,3) for x in d.get("per_rank_ms_per_step")])
    print("strong", d.get("strong")); print("multi", d.get("one_process_multi_gpu")); print("numa", d["config"].get("numa"))
except Exception as e: print("bench parse failed", e)
PY
