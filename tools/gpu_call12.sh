#!/bin/bash
# per-pipe dynamic instruction counts of pretok_kernel<1> for the three classification variants
mkdir -p gpurun_out
for so in tiktoken_b200/csrc/variants/*.so; do
  name=$(basename $so .so); name=${name#libb200bpe_}
  export B200BPE_LIB=$PWD/$so
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_adu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_cbu.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:'pretok_kernel' -s 2 -c 1 --csv --log-file gpurun_out/r2f_pipes_$name.csv \
     python bench.py --bytes 268435456 --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-extras > /dev/null 2>&1
  echo "== $name"; python - <<PY
import csv
for r in csv.reader(open("gpurun_out/r2f_pipes_$name.csv")):
    if len(r) > 5 and r[0].isdigit(): print("  ", r[-3].replace("sm__inst_executed_","").replace("smsp__",""), r[-1])
PY
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('   bench pretok_ms', round(d['stage_ms']['pretok_ms'],3), 'step', round(d['ms_per_step'],3))"
done
