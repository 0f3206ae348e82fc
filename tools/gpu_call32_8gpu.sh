#!/bin/bash
# 4-GPU box: torchrun bench at N=4 (NCCL count exchange per step, strong line, one-process engine over 4 devices)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-configs > gpurun_out/r2v_bench8.json 2> gpurun_out/r2v_bench8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/r2v_bench8.err | cut -c1-200
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2v_bench8.json").read().strip().splitlines()[-1])
    print("value", round(d.get("value"),1), "ms", round(d.get("ms_per_step"),3), "e2e", round(d.get("e2e",{}).get("value"),1), "per rank", [round(x,3) for x in d.get("per_rank_ms_per_step")])
    print("strong", d.get("strong")); print("multi", d.get("one_process_multi_gpu")); print("numa", d["config"].get("numa"))
except Exception as e: print("bench parse failed", e)
PY
