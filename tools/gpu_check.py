#!/usr/bin/env python3
"""First-contact GPU script: parity on small synthetic batches of each config + timing of a full config."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
from f5c_b200 import synth, models
from f5c_b200.abea import AbeaContext

full = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
ctx = AbeaContext(0)
print("device", ctx.device_info())
for cfg, n in (("cfg2", 96), ("cfg3", 48), ("cfg4", 12)):
    b = synth.make_config(cfg, seed=5, n_reads=n)
    k, m = models.load_model(b.meta["model"])
    m = ctx.set_model(m, k)
    a = ctx.align_batch(b)
    p = ol.port_align(b, m)
    try:
        ol.assert_same_alignment(a, p, cfg)
        st = ctx.read_stats(b.n_reads)
        ok = np.array_equal(st["sum_emission"], p.stats["sum_emission"])
        print(cfg, "PARITY OK", "sum_emission bit-equal:", ok, "events", int(b.n_events.sum()), a.timing)
    except AssertionError as e:
        print(cfg, "PARITY FAIL", str(e)[:500])
t0 = time.time()
b = synth.make_config(full, seed=42)
print("generated", full, "in %.1fs" % (time.time() - t0), "reads", b.n_reads, "events", int(b.n_events.sum()),
      "bands", int(b.n_bands.sum()))
k, m = models.load_model(b.meta["model"]); ctx.set_model(m, k)
print("upload", ctx.upload(b))
for it in range(4):
    t = ctx.run()
    print("run", it, {x: round(t[x], 3) for x in ("kmer_ms", "fill_ms", "trace_ms", "kernel_ms")},
          "Mev/s %.1f" % (t["n_events"] / t["kernel_ms"] / 1e3))
a = ctx.download(b)
print("download", {x: a.timing[x] for x in ("d2h_ms", "unpack_ms", "d2h_bytes")}, "pairs", int(a.n_pairs.sum()),
      "reads aligned", int((a.n_pairs > 0).sum()))
