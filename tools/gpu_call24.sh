#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh run config2 config4 2>&1 | tee gpurun_out/r2q_ab.txt
