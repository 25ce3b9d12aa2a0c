#!/usr/bin/env python3
"""Timing of the stages either side of ABEA on a resident batch (device-timed, CUDA events inside the library):
abea_mom_kernel (estimate_scalings_using_mom) and abea_scaling_kernel (postalign + recalibrate_model), next to the
alignment itself. Usage: scaling_run.py cfg2|cfg3|cfg4 [n_reads|-] [runs] [cpu]  ("cpu": also time the oracle port
of both stages on one host thread, per-read ctypes calls included)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from f5c_b200 import synth, models
from f5c_b200.abea import AbeaContext
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] not in ("", "-") else None
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
b = synth.make_config(cfg, seed=42, n_reads=n)
k, m = models.load_model(b.meta["model"])
ctx = AbeaContext(0)
m = ctx.set_model(m, k)
for i in range(runs):
    ctx.upload(b, with_scalings=False)
    est, t1 = ctx.estimate_scalings(b.n_reads)
    t2 = ctx.run()
    t3 = ctx.scaling_stage()
    print(cfg, "reads", b.n_reads, "events %.1fM" % (b.n_events.sum() / 1e6), "| mom_ms %.3f  abea kernel_ms %.3f  scaling_ms %.3f"
          % (t1["mom_ms"], t2["kernel_ms"], t3["scaling_ms"]))
assert est["shift"].tobytes() == b.scalings["shift"].tobytes() and est["scale"].tobytes() == b.scalings["scale"].tobytes()
sc = ctx.scaling_download(b)
print("flags: ok %d, failed calibration %d, failed alignment %d, failed quality %d" % (
    (sc.results["flags"] == 0).sum(), (sc.results["flags"] & 1).sum(), ((sc.results["flags"] & 2) != 0).sum(),
    ((sc.results["flags"] & 4) != 0).sum()))
if len(sys.argv) > 4 and sys.argv[4] == "cpu":
    import oracle_lib as ol
    aln = ctx.download(b)
    t0 = time.time(); e = ol.port_estimate_scalings(b, m); t1 = time.time()
    s = ol.port_scaling(b, m, aln); t2 = time.time()
    print("oracle port, 1 thread: estimate %.1f ms, scaling_single %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    assert e.tobytes() == est.tobytes()
    ol.assert_same_scaling(ol.ScalingResult(b, sc.results, sc.maps), s, cfg)
    print("parity with the oracle on all %d reads: OK" % b.n_reads)
