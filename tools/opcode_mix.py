#!/usr/bin/env python3
"""Dynamic opcode mix of one kernel from an .ncu-rep (source page): warp-level executions per opcode, per band.
Usage: opcode_mix.py report.ncu-rep n_bands [n_traceback_steps]"""
import csv, io, subprocess, sys, collections
rep, bands = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(src)) if len(r) > 10]
h = rows[0]
ci = [i for i, x in enumerate(h) if x.strip() == "Source"][0]
ce = [i for i, x in enumerate(h) if x.strip() == "Instructions Executed"][0]
cs = [i for i, x in enumerate(h) if x.strip() in ("# Samples", "Warp Stall Sampling (All Samples)")][0]
tot = collections.Counter(); smp = collections.Counter(); n = 0; static = 0
hist = collections.Counter()
for r in rows[1:]:
    try:
        ex = int(float(r[ce] or 0)); s = int(float(r[cs] or 0))
    except Exception:
        continue
    ins = r[ci].strip()
    parts = ins.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    tot[op] += ex; smp[op] += s; n += ex; static += 1
    hist[round(ex / 1e6, 1)] += 1
allsmp = sum(smp.values()) or 1
print(f"warp instructions executed: {n} = {n / bands:.1f} per band ({static} static SASS instructions)")
print(f"{'opcode':28s} {'per band':>9s} {'share':>7s} {'stall samples':>14s}")
for op, v in tot.most_common(40):
    print(f"{op:28s} {v / bands:9.2f} {100.0 * v / n:6.1f}% {100.0 * smp[op] / allsmp:13.1f}%")
print("\ninstructions by execution count (millions of warp-level executions -> number of distinct SASS instructions):")
for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:12]:
    print(f"  {k:8.1f} M : {v}")
