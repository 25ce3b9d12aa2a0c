#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + the hottest SASS lines by stall samples.
Usage: ncu_summary.py report.ncu-rep [n_hot_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__inst_executed.sum",
        "smsp__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct"]
for k in keys:
    if k in d:
        print(f"{k:70s} {d[k][1]:>18s} {d[k][0]}")
print("-- stall reasons (warps per issue-active cycle) --")
st = [(float(v[1] or 0), k) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
for v, k in sorted(st, reverse=True)[:9]:
    print(f"  {k.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
rows = [r for r in rows if len(r) > 10]
if rows:
    h = rows[0]
    def col(name):
        for i, x in enumerate(h):
            if x.strip() == name:
                return i
        return None
    ci, cs, ce = col("Source"), col("# Samples") if col("# Samples") is not None else col("Warp Stall Sampling (All Samples)"), col("Instructions Executed")
    if cs is None:
        print("columns:", h[:40])
    else:
        data = []
        for r in rows[1:]:
            try:
                data.append((int(float(r[cs] or 0)), r[ci], r[ce] if ce is not None else ""))
            except Exception:
                pass
        tot = sum(x[0] for x in data) or 1
        print(f"-- hottest SASS lines of {len(data)} (stall samples, % of {tot}) --")
        for smp, ins, ex in sorted(data, reverse=True)[:nhot]:
            print(f"  {100.0*smp/tot:5.1f}%  {ins[:100]}   [exec {ex}]")
