#!/usr/bin/env python3
"""Workload for ncu captures: one config resident on the GPU, run the three kernels a few times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from f5c_b200 import synth, models
from f5c_b200.abea import AbeaContext
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] not in ("", "-") else None
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
world = int(os.environ.get('PROF_WORLD', '1'))
b = synth.make_config_shard(cfg, 0, world, seed=42, reads_per_gpu=n) if world > 1 else synth.make_config(cfg, seed=42, n_reads=n)
k, m = models.load_model(b.meta["model"])
ctx = AbeaContext(0); ctx.set_model(m, k)
ctx.upload(b)
for i in range(runs):
    t = ctx.run()
    print({x: round(t[x], 3) for x in ("kmer_ms", "fill_ms", "trace_ms", "kernel_ms")}, "Mev/s %.1f" % (t["n_events"] / t["kernel_ms"] / 1e3), "n_wide", t["n_wide"])

import numpy as np
cyc = ctx.read_cycles(b.n_reads)
nb = b.n_bands
order = np.argsort(-nb)[:6]
ev = b.n_events
print("longest reads: bands", nb[order].tolist())
print("  fill cycles/band ", np.round(cyc["fill_cycles"][order] / nb[order], 1).tolist(), " wide", cyc["wide"][order].tolist())
print("  trace cycles/step", np.round(cyc["trace_cycles"][order] / ev[order], 1).tolist())
sel = np.argsort(nb)[len(nb)//2 - 3: len(nb)//2 + 3]
print("median reads: fill cycles/band", np.round(cyc["fill_cycles"][sel] / nb[sel], 1).tolist(), " trace cycles/step", np.round(cyc["trace_cycles"][sel] / ev[sel], 1).tolist())
try:
    rs = ctx.read_respec(b.n_reads)
    print("traceback: segments re-checked per read: mean %.2f max %d; reads with any %d of %d; margin %s; mode %s" % (
        rs.mean(), rs.max(), int((rs > 0).sum()), b.n_reads, os.environ.get("ABEA_TB_MARGIN", "default"), os.environ.get("ABEA_TB", "default")))
except Exception as e:
    print("respec n/a", e)
print("scheduler model", ctx.scheduler_model())
print("sum fill Mcycles", cyc["fill_cycles"].sum() / 1e6, "sum trace Mcycles", cyc["trace_cycles"].sum() / 1e6, "max fill+trace Mcycles", (cyc["fill_cycles"] + cyc["trace_cycles"]).max() / 1e6)
