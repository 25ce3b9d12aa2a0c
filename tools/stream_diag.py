#!/usr/bin/env python3
"""Where the time goes in abea_align_batch under each ABEA_STREAM mode: e2e time, per-phase timings, when reads
start relative to the first one (by rank in the longest-first schedule) and what a band costs them.
Usage: stream_diag.py <config> <modes e.g. 0,1,3> [iterations] [extra env K=V ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from f5c_b200 import synth, models
from f5c_b200.abea import AbeaContext
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "3"]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
for kv in sys.argv[4:]:
    k_, v_ = kv.split("=")
    os.environ[k_] = v_
b = synth.make_config(cfg, seed=42)
k, m = models.load_model(b.meta["model"])
order = np.argsort(-b.n_bands, kind="stable")
for mode in modes:
    os.environ["ABEA_STREAM"] = mode
    ctx = AbeaContext(0); ctx.set_model(m, k)
    pb = ctx.pin_batch(b)
    out = ctx.alloc_output(b, pinned=True)
    for _ in range(3):
        r = ctx.align_batch(pb, out)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        r = ctx.align_batch(pb, out)
        ts.append((time.perf_counter() - t0) * 1e3)
    t = r.timing
    print(cfg, "STREAM", mode, "e2e ms min/med %.2f %.2f" % (min(ts), float(np.median(ts))),
          {x: round(t[x], 2) for x in ("pack_ms", "h2d_ms", "load_ms", "kmer_ms", "fill_ms", "kernel_ms", "d2h_ms")},
          "streamed", t["streamed"], "wide", t["n_wide"], "pairs", int(r.n_pairs.sum()), flush=True)
    st = ctx.read_starts(b.n_reads).astype(np.int64)
    cyc = ctx.read_cycles(b.n_reads)
    ok = st >= 0
    rel = (st - st[ok].min()) / 1e3
    fpb = cyc["fill_cycles"] / np.maximum(1, b.n_bands)
    dur = (cyc["fill_cycles"] + cyc["trace_cycles"]) / 1.9e6
    el = order[ok[order]]
    print("  12 longest: start ms", np.round(rel[el[:12]], 2).tolist())
    print("              dur ms  ", np.round(dur[el[:12]], 2).tolist())
    print("              cyc/band", np.round(fpb[el[:12]], 0).tolist())
    for lo, hi in ((0, 148), (148, 592), (592, 1200), (1200, 2000), (2000, 3000), (3000, len(el))):
        seg = el[lo:hi]
        if len(seg):
            print("  ranks %4d-%4d: start ms min/med/max %6.2f %6.2f %6.2f | end max %6.2f | cyc/band med %5.0f | trace cyc/step med %4.0f" %
                  (lo, hi, rel[seg].min(), np.median(rel[seg]), rel[seg].max(), (rel[seg] + dur[seg]).max(),
                   np.median(fpb[seg]), np.median(cyc["trace_cycles"][seg] / np.maximum(1, r.n_pairs[seg]))), flush=True)
    late = [i for i in range(min(200, len(el))) if rel[el[i]] > 1.0]
    if late:
        print("  late starters among the 200 longest (rank, bands, start ms, dur ms, cyc/band, wide):",
              [(i, int(b.n_bands[el[i]]), round(float(rel[el[i]]), 2), round(float(dur[el[i]]), 2), int(fpb[el[i]]), int(cyc["wide"][el[i]])) for i in late][:12])
    ctx.close()
