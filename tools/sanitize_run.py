#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel, both fill forms, edge cases."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from edge_cases import edge_batch
from f5c_b200 import synth, models
from f5c_b200.abea import AbeaContext
import oracle_lib as ol

for wide in ("0", "1"):
    os.environ["ABEA_WIDE"] = wide
    with AbeaContext(0) as ctx:
        for name, b in (("r9", synth.make_batch("r9", n_reads=24, mean_events=900, sigma=0.6, epk=1.8, seed=3)),
                        ("r10", synth.make_batch("r10", n_reads=200, mean_events=400, sigma=1.0, epk=1.9, seed=4)),
                        ("r9", edge_batch())):
            k, m = models.load_model(name)
            m = ctx.set_model(m, k)
            a = ctx.align_batch(b)
            ol.assert_same_alignment(a, ol.port_align(b, m), f"sanitize {name} wide={wide}")
            print("ok", name, "wide", wide, "n_wide", a.timing["n_wide"], "pairs", int(a.n_pairs.sum()))
            # pinned buffers: events streamed in by abea_load_kernel, pair lists written to the caller's buffer
            out = ctx.alloc_output(b, pinned=True)
            a = ctx.align_batch(ctx.pin_batch(b), out)
            assert a.timing["streamed"] == 3
            ol.assert_same_alignment(a, ol.port_align(b, m), f"sanitize {name} wide={wide} streamed")
            print("ok", name, "wide", wide, "streamed", a.timing["streamed"], "pairs", int(a.n_pairs.sum()))
