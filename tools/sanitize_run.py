#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the library — prepare, both
fill forms (resident and streamed instantiations) with both traceback forms, loader (AoS and means), extraction,
method-of-moments, scaling, the five event-detection kernels, the BLOW5 decoders, the result compaction, the ragged
front door — on batches small enough for the tool's 10-100x slowdown, each checked against the oracle.
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py        (shared-memory hazards)
    compute-sanitizer --tool synccheck python tools/sanitize_run.py"""
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
            assert a.timing["streamed"] == 5
            ol.assert_same_alignment(a, ol.port_align(b, m), f"sanitize {name} wide={wide} streamed")
            print("ok", name, "wide", wide, "streamed", a.timing["streamed"], "pairs", int(a.n_pairs.sum()))

# the stages either side, the chain, BLOW5, compaction, ragged front door
import blow5
from f5c_b200.abea import scaling_db
from f5c_b200.batch import EVENT_DTYPE, SCALINGS_DTYPE, ReadBatch
os.environ["ABEA_WIDE"] = "1"
for tb in ("0", "1", "3"):
    os.environ["ABEA_TB"] = tb
    with AbeaContext(0) as ctx:
        b = synth.make_batch("r9", n_reads=40, mean_events=700, sigma=0.7, epk=1.8, seed=9)
        k, m = models.load_model("r9")
        m = ctx.set_model(m, k)
        want = ol.port_align(b, m)
        ol.assert_same_alignment(ctx.align_batch(b, means=b.event_means()), want, "means tb=" + tb)
        out = ctx.alloc_output(b, pinned=True)
        ol.assert_same_alignment(ctx.align_batch(ctx.pin_batch(b), out, means=ctx.pin_array(b.event_means())), want, "means streamed")
        ol.assert_same_alignment(ctx.align_ragged(b, threads=3), want, "ragged")
        print("ok traceback form", tb)
os.environ.pop("ABEA_TB")
with AbeaContext(0) as ctx:
    k, m = models.load_model("r9")
    m = ctx.set_model(m, k)
    ctx.upload(b, with_scalings=False)
    est, _ = ctx.estimate_scalings(b.n_reads)
    ctx.run()
    a = ctx.download(b)
    sc = scaling_db(ctx, b)
    dense = np.zeros(int(b.pair_capacity().sum()) + 1, dtype=a.pairs.dtype)
    import torch
    dd = torch.zeros((dense.shape[0], 2), dtype=torch.int32, device="cuda")
    total = ctx.compact_results(dd.data_ptr(), dd.shape[0])
    assert total == int(a.n_pairs.sum())
    print("ok scaling stages + compaction", total)
    # the lists as path codes in device memory (abea_pairs_to_codes_kernel) and back (abea_expand_codes_kernel)
    _dp, dn, cap, n = ctx.device_results()
    dc, n_words = ctx.device_codes()
    cap_ptr = np.concatenate([[0], np.cumsum(b.pair_capacity().astype(np.int64))])
    d2 = torch.zeros((int(cap), 2), dtype=torch.int32, device="cuda")
    total2 = ctx.expand_codes(dc, dn, cap_ptr, d2.data_ptr(), int(cap), sync=True)
    assert total2 == total and bool(torch.equal(d2[:total2], dd[:total]))
    print("ok device codes + expansion", total2, n_words)
    f = blow5.Blow5(os.path.join(ROOT, "tests", "golden", "ecoli", "ecoli8_zlib_svbzd.blow5"))
    order = np.argsort([r[1] for r in f.records])[:3]
    chunks = [f.record_bytes(int(i)) for i in order]
    rec_len = np.array([len(c) for c in chunks], dtype=np.int32)
    rec_ptr = np.zeros(len(chunks), dtype=np.int64); np.cumsum(rec_len[:-1].astype(np.int64), out=rec_ptr[1:])
    nev, ns, _ = ctx.getevents_blow5(np.frombuffer(b"".join(chunks), dtype=np.uint8).copy(), rec_ptr, rec_len, f.record_method, f.signal_method)
    raw, raw_ptr = ctx.raw_download(ns)
    for j, i in enumerate(order):
        assert np.array_equal(raw[int(raw_ptr[j]):int(raw_ptr[j]) + int(ns[j])], f.read(int(i))[5].astype(np.float32))
    print("ok blow5 decode + events", int(nev.sum()))
    fx = blow5.Blow5(os.path.join(ROOT, "tests", "golden", "ecoli", "ecoli8_zlib_exzd.blow5"))     # ex-zd signal compression
    chunks = [fx.record_bytes(int(i)) for i in order]
    rec_len = np.array([len(c) for c in chunks], dtype=np.int32)
    rec_ptr = np.zeros(len(chunks), dtype=np.int64); np.cumsum(rec_len[:-1].astype(np.int64), out=rec_ptr[1:])
    nev, ns, _ = ctx.getevents_blow5(np.frombuffer(b"".join(chunks), dtype=np.uint8).copy(), rec_ptr, rec_len, fx.record_method, fx.signal_method)
    raw, raw_ptr = ctx.raw_download(ns)
    for j, i in enumerate(order):
        assert np.array_equal(raw[int(raw_ptr[j]):int(raw_ptr[j]) + int(ns[j])], f.read(int(i))[5].astype(np.float32))
    print("ok blow5 ex-zd decode", int(ns.sum()))
    sg, seq, seq_ptr, read_len, k = synth.make_signal_batch("r9", 6, 300, 0.4, seed=12)
    _, _, nev, _ = ctx.getevents(sg["raw"].astype(np.int16), sg["raw_ptr"], sg["n_samples"], (sg["offset"], sg["range"], sg["digitisation"]), download=False)
    shell = ReadBatch(seq, seq_ptr, read_len, np.zeros(0, dtype=EVENT_DTYPE), np.zeros(6, dtype=np.int64), nev.astype(np.int32),
                      np.zeros(6, dtype=SCALINGS_DTYPE), np.ones(6, dtype=np.uint8), k)
    ctx.upload(shell, with_scalings=False, device_events=True)
    ctx.estimate_scalings(6); ctx.run(); ctx.download(shell); scaling_db(ctx, shell, 100)
    print("ok chain from int16 signals")
