#!/usr/bin/env python3
"""Timing of event detection on the device (abea_events_kernel, CUDA events inside the library) on synthetic raw
signals with cfg2's event counts, next to the oracle port on one host thread.
Usage: events_run.py [n_reads] [mean_events] [runs] [cpu]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from f5c_b200 import synth
from f5c_b200.abea import AbeaContext
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mean = float(sys.argv[2]) if len(sys.argv) > 2 else 4000
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sg = synth.make_signals(n, mean, 0.5, seed=42)
ctx = AbeaContext(0)
cal = (sg["offset"], sg["range"], sg["digitisation"])
for i in range(runs):
    t0 = time.perf_counter()
    ev, ptr, nev, t = ctx.getevents(sg["raw"], sg["raw_ptr"], sg["n_samples"], cal)
    wall = (time.perf_counter() - t0) * 1e3
    ns = int(sg["n_samples"].sum())
    print("reads %d samples %.1fM events %.1fM | events_ms %.3f (%.1f Gsamples/s) h2d_ms %.2f wall incl. download %.1f ms | longest signal %d samples"
          % (n, ns / 1e6, nev.sum() / 1e6, t["events_ms"], ns / t["events_ms"] / 1e6, t["h2d_ms"], wall, sg["n_samples"].max()))
if len(sys.argv) > 4 and sys.argv[4] == "cpu":
    import oracle_lib as ol
    idx = np.random.default_rng(1).choice(n, min(n, 256), replace=False)
    t0 = time.time(); tot = 0
    for i in idx:
        pa = sg["pa"][sg["raw_ptr"][i]:sg["raw_ptr"][i] + sg["n_samples"][i]]
        w = ol.port_getevents(pa); tot += len(pa)
        assert ol._events_equal(ev[ptr[i]:ptr[i] + nev[i]], w), i
    dt = time.time() - t0
    print("oracle port, 1 thread: %.1f ns/sample (%d signals, parity OK) -> whole batch %.0f ms" % (dt / tot * 1e9, len(idx), dt / tot * ns * 1e3))
