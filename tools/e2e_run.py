#!/usr/bin/env python3
"""End-to-end timing of abea_align_batch with pinned host buffers (the bench's e2e leg), for sweeps of the streaming
knobs: ABEA_STREAM, ABEA_LOAD_CTAS. Usage: e2e_run.py <config> [n_reads|-] [iterations]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from f5c_b200 import synth, models
from f5c_b200.abea import AbeaContext
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] not in ("", "-") else None
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
b = synth.make_config(cfg, seed=42, n_reads=n)
k, m = models.load_model(b.meta["model"])
ctx = AbeaContext(0); ctx.set_model(m, k)
pb = ctx.pin_batch(b)
out = ctx.alloc_output(b, pinned=True)
means = ctx.pin_array(b.event_means()) if os.environ.get("E2E_MEANS", "1") != "0" else None   # 4 B per event in (the bench's leg)
if os.environ.get("E2E_RESIDENT"):      # the same batch resident: upload once, time abea_run alone
    ctx.upload(pb, means=means)
    for _ in range(3):
        ctx.run()
for _ in range(3):
    r = ctx.align_batch(pb, out, means=means)
ts = []
for _ in range(iters):
    t0 = time.perf_counter()
    r = ctx.align_batch(pb, out, means=means)
    ts.append((time.perf_counter() - t0) * 1e3)
if os.environ.get("E2E_RESIDENT"):
    ctx.upload(pb, means=means)
    for _ in range(3):
        tr = ctx.run()
    print("resident run kernel_ms %.3f" % tr["kernel_ms"])
ev = b.events_aligned()
t = r.timing
print(cfg, "STREAM", os.environ.get("ABEA_STREAM", "3"), "LOAD_CTAS", os.environ.get("ABEA_LOAD_CTAS", "-"),
      "e2e ms min/med %.2f %.2f" % (min(ts), float(np.median(ts))), "Mev/s %.1f" % (ev / np.median(ts) / 1e3),
      {x: round(t[x], 2) for x in ("pack_ms", "h2d_ms", "load_ms", "kernel_ms", "d2h_ms", "unpack_ms")}, "streamed", t["streamed"],
      "pairs", int(r.n_pairs.sum()))

if os.environ.get("E2E_STARTS"):
    st = ctx.read_starts(b.n_reads).astype(np.int64)
    print("timeline of the", "resident run" if os.environ.get("E2E_RESIDENT") else "last streamed call")
    cyc = ctx.read_cycles(b.n_reads)
    ok = st >= 0
    t0_ = st[ok].min()
    rel = (st - t0_) / 1e3
    dur = (cyc["fill_cycles"] + cyc["trace_cycles"]) / 1.9e6      # ms at ~1.9 GHz
    order = np.argsort(-b.n_bands)
    print("start ms of the 12 longest reads:", np.round(rel[order[:12]], 2).tolist())
    print("  their duration ms:", np.round(dur[order[:12]], 2).tolist())
    print("  wide:", cyc["wide"][order[:12]].tolist(), " fill cycles/band:", np.round(cyc["fill_cycles"][order[:12]] / b.n_bands[order[:12]], 0).tolist(),
          " n_wide", (tr if os.environ.get("E2E_RESIDENT") else t)["n_wide"], " model", ctx.scheduler_model())
    o2 = order[12:40]
    print("  LPT ranks 12-40: wide", int(cyc["wide"][o2].sum()), "fill cycles/band min/med/max %.0f %.0f %.0f" % tuple(np.percentile(cyc["fill_cycles"][o2] / b.n_bands[o2], [0, 50, 100])))
    el = order[ok[order]]
    for lo, hi in ((0, 148), (148, 592), (592, 1500), (1500, 3000), (3000, len(el))):
        seg = el[lo:hi]
        if len(seg):
            print(f"  LPT ranks {lo}-{hi}: start ms min/med/max %.2f %.2f %.2f; end max %.2f" %
                  (rel[seg].min(), np.median(rel[seg]), rel[seg].max(), (rel[seg] + dur[seg]).max()))
