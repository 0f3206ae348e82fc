#!/bin/bash
# dev helper: A/B of two builds of the library on one box (new = in-tree, old = libold_dev.so)
mkdir -p gpurun_out
cp tiktoken_b200/csrc/libb200bpe.so /tmp/new.so
for v in new old; do
  [ $v = old ] && cp tiktoken_b200/csrc/libold_dev.so tiktoken_b200/csrc/libb200bpe.so
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'miss_kernel|find_long|mid_thread|probe' -c 24 --csv --log-file gpurun_out/ab_${v}.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ab_${v}.log 2>&1
  echo "== $v rc=$?"
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/ab_${v}.csv")) if len(r)>5 and r[0].isdigit()]
best=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0].replace("void ","")[:40]; x=float(r[-1])/1e3
    best[k]=max(best.get(k,0),x)
print({k:round(x,1) for k,x in best.items()})
PY
  timeout 600 ncu --set full --clock-control none -k regex:'miss_kernel' -s 2 -c 1 -o gpurun_out/miss_full_${v} -f \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/abf_${v}.log 2>&1
  ncu -i gpurun_out/miss_full_${v}.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active 2>/dev/null | tail -2
done
cp /tmp/new.so tiktoken_b200/csrc/libb200bpe.so
