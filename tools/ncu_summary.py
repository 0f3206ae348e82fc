#!/usr/bin/env python
"""Summarise an .ncu-rep (brought back from the GPU box in gpurun_out/) into profiles/:
per-kernel duration, DRAM bytes, issue / occupancy figures, and the top source lines by stall
samples.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx"""
import csv, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
        "launch__grid_size", "launch__block_size"]


def unit_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return float(v) * f


def main(rep, out_prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines, traffic = [], {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        lines.append(f"## {name}")
        for w in WANT:
            if w in ix:
                lines.append(f"{w} = {r[ix[w]]} {units[ix[w]]}")
        rd = unit_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = unit_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        key = name.split("<")[0]                      # template instances share a key: keep the launch that moves the most
        if key not in traffic or rd + wr > traffic[key]["dram_bytes_per_launch"]:
            traffic[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                            "duration_under_ncu": float(r[ix["gpu__time_duration.sum"]]), "duration_unit": units[ix["gpu__time_duration.sum"]]}
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                              "--kernel-name", "regex:" + name.split("<")[0]], capture_output=True, text=True).stdout
        data, cur = [], None
        for s in csv.reader(src.splitlines()):
            if len(s) >= 2 and s[0] == "File Path":
                cur = s[1].split("/")[-1]
            elif len(s) > 10 and s[0].isdigit():
                try:
                    data.append((int(s[6]), int(s[7]), s[10], cur, int(s[0]), s[1].strip()[:90]))
                except ValueError:
                    pass
        tot = sum(d[0] for d in data) or 1
        toti = sum(d[1] for d in data) or 1
        lines.append(f"top source lines by warp-stall samples (total {tot}, warp instructions {toti}):")
        for d in sorted(data, reverse=True)[:16]:
            lines.append(f"  {100 * d[0] / tot:5.1f}% samples {100 * d[1] / toti:5.1f}% inst  thr/inst={d[2]:>4}  {d[3]}:{d[4]}  {d[5]}")
        lines.append("")
    open(out_prefix + ".txt", "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(out_prefix + "_traffic.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
