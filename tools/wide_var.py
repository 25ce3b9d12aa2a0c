#!/usr/bin/env python3
"""Run-to-run variation of the wide reads: per run the kernel time and the fill cycles/band of the wide reads.
Usage: PROF_WORLD=2 wide_var.py cfg2 [runs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from f5c_b200 import synth, models
from f5c_b200.abea import AbeaContext
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
world = int(os.environ.get('PROF_WORLD', '1'))
b = synth.make_config_shard(cfg, 0, world, seed=42) if world > 1 else synth.make_config(cfg, seed=42)
k, m = models.load_model(b.meta["model"])
ctx = AbeaContext(0); ctx.set_model(m, k)
ctx.upload(b)
nb = b.n_bands
for i in range(runs):
    t = ctx.run()
    cyc = ctx.read_cycles(b.n_reads)
    w = np.flatnonzero(cyc["wide"] == 1)
    w = w[np.argsort(-nb[w])]
    st = ctx.read_starts(b.n_reads).astype(np.int64)
    rel = (st - st[st >= 0].min()) / 1e3
    print("run %d kernel_ms %.2f | wide reads: bands %s cyc/band %s start_ms %s | max narrow end? longest total Mcyc %.1f" % (
        i, t["kernel_ms"], nb[w][:4].tolist(), np.round(cyc["fill_cycles"][w] / nb[w], 0)[:7].tolist(),
        np.round(rel[w], 2)[:4].tolist(), (cyc["fill_cycles"] + cyc["trace_cycles"]).max() / 1e6), flush=True)
