#!/bin/bash
# 2-GPU box: the one-process multi-GPU engine test, then the torchrun bench at N=2 (NCCL count exchange, strong line, configs)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_paths.py -q -m gpu -x -k "multi_gpu or queued" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l_bench2.json 2> gpurun_out/r2l_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2l_bench2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2l_bench2.json").read().strip().splitlines()[-1])
    print("value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", d.get("e2e",{}).get("value"), "per rank", d.get("per_rank_ms_per_step"))
    print("stage", d.get("stage_ms"))
    print("strong", d.get("strong"))
    print("multi", d.get("one_process_multi_gpu"))
    for k,v in (d.get("configs") or {}).items(): print(k, {x: v.get(x) for x in ("parity","value","ms_per_step","error")}, (v.get("e2e") or {}).get("value"))
    print("numa", d["config"].get("numa"))
except Exception as e: print("bench parse failed", e)
PY

