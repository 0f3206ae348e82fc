#!/usr/bin/env python
"""dev helper: one compact line block per captured kernel of an .ncu-rep: duration, pipes, L1TEX, stall reasons."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
KEYS = [("dur", "gpu__time_duration.sum"), ("inst_M", "smsp__inst_executed.sum"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"), ("fmaheavy%", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), ("l1tex%", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        ("l1wave%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"), ("l1hit%", "l1tex__t_sector_hit_rate.pct"),
        ("l2hit%", "lts__t_sector_hit_rate.pct"), ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"), ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum")]
stall = [i for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
seen = set()
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    if name in seen and "--all" not in sys.argv: continue
    seen.add(name)
    out = []
    for k, m in KEYS:
        if m in ix:
            try: v = float(r[ix[m]])
            except ValueError: continue
            if k == "inst_M": v /= 1e6
            if k == "dur": out.append(f"dur={v:.3f}{rows[1][ix[m]]}"); continue
            if k.endswith("_MB"): v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(rows[1][ix[m]], 1.0); out.append(f"{k}={v:.1f}"); continue
            out.append(f"{k}={v:.1f}")
    tot = sum(float(r[i]) for i in stall) or 1
    top = sorted(stall, key=lambda i: -float(r[i]))[:5]
    print("##", name); print("  ", " ".join(out))
    print("   stalls:", ", ".join(f"{hdr[i].replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * float(r[i]) / tot:.0f}%" for i in top))
