#!/usr/bin/env python3
"""The device-resident chain raw signal -> events -> scalings -> alignment -> recalibration on a cfg2-sized batch of
synthetic signals generated from their sequences (synth.make_signal_batch): per-stage device times.
Usage: chain_run.py [model] [n_reads] [mean_kmers] [runs]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from f5c_b200 import synth, models
from f5c_b200.abea import AbeaContext
from f5c_b200.batch import EVENT_DTYPE, SCALINGS_DTYPE, ReadBatch
model = sys.argv[1] if len(sys.argv) > 1 else "r9"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mean_k = float(sys.argv[3]) if len(sys.argv) > 3 else 2260
runs = int(sys.argv[4]) if len(sys.argv) > 4 else 3
sg, seq, seq_ptr, read_len, k = synth.make_signal_batch(model, n, mean_k, 0.5, 42)
kk, m = models.load_model(model)
ctx = AbeaContext(0); ctx.set_model(m, kk)
cal = (sg["offset"], sg["range"], sg["digitisation"])
for i in range(runs):
    t0 = time.perf_counter()
    _, _, nev, t1 = ctx.getevents(sg["raw"], sg["raw_ptr"], sg["n_samples"], cal, download=False)
    shell = ReadBatch(seq, seq_ptr, read_len, np.zeros(0, dtype=EVENT_DTYPE), np.zeros(n, dtype=np.int64),
                      nev.astype(np.int32), np.zeros(n, dtype=SCALINGS_DTYPE), np.ones(n, dtype=np.uint8), k)
    ctx.upload(shell, with_scalings=False, device_events=True)
    est, t2 = ctx.estimate_scalings(n)
    t3 = ctx.run()
    t4 = ctx.scaling_stage()
    wall = (time.perf_counter() - t0) * 1e3
    sc = ctx.scaling_download(shell)
    aln = ctx.download(shell)
    print("%s reads %d samples %.1fM events %.1fM | raw h2d %.2f ms (pageable) | getevents %.2f  mom %.2f  abea %.2f  scaling %.2f ms | device total %.2f ms | wall %.1f ms | aligned %.3f calibrated-ok %.3f"
          % (model, n, sg["n_samples"].sum() / 1e6, nev.sum() / 1e6, t1["h2d_ms"], t1["events_ms"], t2["mom_ms"], t3["kernel_ms"],
             t4["scaling_ms"], t1["events_ms"] + t2["mom_ms"] + t3["kernel_ms"] + t4["scaling_ms"], wall,
             (aln.n_pairs > 0).mean(), (sc.results["flags"] == 0).mean()), flush=True)
