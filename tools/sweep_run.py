#!/usr/bin/env python3
"""Resident-batch sweeps of the scheduling knobs in ONE process (the batch is generated once): each combination gets
its own context (the knobs are read from the environment by abea_create).
Usage: sweep_run.py <config> "K1=V1 K2=V2;K1=V3;..." [runs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from f5c_b200 import synth, models
from f5c_b200.abea import AbeaContext
cfg = sys.argv[1]
combos = [c.strip() for c in sys.argv[2].split(";")]
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 4
world = int(os.environ.get('PROF_WORLD', '1'))
b = synth.make_config_shard(cfg, 0, world, seed=42) if world > 1 else synth.make_config(cfg, seed=42)
k, m = models.load_model(b.meta["model"])
base = dict(os.environ)
for combo in combos:
    os.environ.clear(); os.environ.update(base)
    for kv in combo.split():
        a, v = kv.split("="); os.environ[a] = v
    ctx = AbeaContext(0); ctx.set_model(m, k)
    ctx.upload(b)
    ts = [ctx.run() for _ in range(runs)]
    best = min(ts, key=lambda t: t["kernel_ms"])
    cyc = ctx.read_cycles(b.n_reads)
    tot = (cyc["fill_cycles"] + cyc["trace_cycles"])
    print("%-5s %-60s kernel ms min %.2f med %.2f | n_wide %d | longest read Mcyc %.1f" % (
        cfg, combo or "(defaults)", best["kernel_ms"], float(np.median([t["kernel_ms"] for t in ts])), best["n_wide"],
        tot.max() / 1e6), flush=True)
    ctx.close()
