#!/usr/bin/env python3
"""profiles/traffic_r02.json from an `ncu --set full` capture of the dominant kernel (read here, no GPU needed): the DRAM
bytes of one launch and what binds the kernel, for bench.py's roofline.traffic / binding_resource (which name this file as
their source). Usage: traffic_json.py report.ncu-rep config n_bands out.json"""
import csv, io, json, subprocess, sys
rep, config, bands, out = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
def val(k, scale=1.0):
    u, v = d[k]
    v = float(v.replace(",", ""))
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v * mult * scale
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = [r for r in csv.reader(io.StringIO(src)) if len(r) > 10]
h = srows[0]
ce = [i for i, x in enumerate(h) if x.strip() == "Instructions Executed"][0]
n_inst = sum(int(float(r[ce] or 0)) for r in srows[1:] if r[ce].replace(".", "").isdigit())
j = {"config": config, "kernel": d["Kernel Name"][1][:60],
     "source": f"profiles/ (ncu --set full --clock-control none of the dominant kernel, one launch, {config}, tools/prof_run.py; summary in profiles/fill_narrow_final_r02.txt)",
     "gpu_time_ms": val("gpu__time_duration.sum"), "registers_per_thread": int(float(d["launch__registers_per_thread"][1])),
     "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
     "sm_cycles_elapsed_max": val("sm__cycles_elapsed.max"), "smsp_cycles_active_avg": val("smsp__cycles_active.avg"),
     "warp_instructions": n_inst, "warp_instructions_per_band": n_inst / bands,
     "issue_active_per_cycle_active": n_inst / (val("smsp__cycles_active.avg") * 148 * 4),
     "pipes_pct_of_peak_active": {k: round(val(f"sm__inst_executed_pipe_{k}.avg.pct_of_peak_sustained_active"), 1) for k in ("alu", "fp64", "fma", "xu", "lsu")}}
json.dump(j, open(out, "w"), indent=1)
print(json.dumps(j, indent=1))
