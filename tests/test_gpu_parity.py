"""Parity tests proper (run on the B200 with -m gpu): the CUDA path, called through the C ABI (ctypes ->
libabea_b200.so), against the oracle on the same seeded inputs; the committed golden fixtures; and, at BASELINE.json's
full sizes, size-independent properties of the alignment."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from edge_cases import edge_batch
from f5c_b200 import models, synth
from f5c_b200.abea import AbeaContext, align_db
from f5c_b200.batch import ReadBatch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ctx(built):
    c = AbeaContext(0)
    yield c
    c.close()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_and_check(ctx, batch, model_name, what):
    k, m = models.load_model(model_name)
    m = ctx.set_model(m, k)
    got = align_db(ctx, batch)
    st = ctx.read_stats(batch.n_reads)
    want = ol.port_align(batch, m)
    ol.assert_same_alignment(got, want, what)     # integer outputs: bit-exact
    sched = want.stats["n_bands"] > 0
    # emission log-probabilities: north_star allows 1e-4 relative; we get bit equality
    np.testing.assert_allclose(st["sum_emission"][sched], want.stats["sum_emission"][sched], rtol=1e-4)
    assert np.array_equal(st["sum_emission"][sched], want.stats["sum_emission"][sched])
    assert np.array_equal(st["end_event"][sched], want.stats["end_event"][sched])
    assert got.timing["kernel_launches"] >= 2
    assert got.timing["streamed"] == 4            # pageable numpy buffers: in through the copy engine, out as path codes
    # the same batch from pinned buffers: events streamed in by abea_load_kernel, pair lists out as path codes that host
    # threads expand into the caller's buffer while the kernels run
    pb = ctx.pin_batch(batch)
    out = ctx.alloc_output(batch, pinned=True)
    out[0].view(np.uint8)[...] = 0xA5
    gs = ctx.align_batch(pb, out)
    if batch.n_reads > 0 and int(batch.pair_capacity().sum()) > 0:
        assert gs.timing["streamed"] == 5
    ol.assert_same_alignment(gs, want, what + " (streamed)")
    st2 = ctx.read_stats(batch.n_reads)
    assert np.array_equal(st2["sum_emission"][sched], want.stats["sum_emission"][sched])
    # the event means alone (abea_batch_t.event_means, 4 B per event over PCIe): pageable, then pinned and streamed
    gm = ctx.align_batch(batch, means=batch.event_means())
    ol.assert_same_alignment(gm, want, what + " (means)")
    assert gm.timing["h2d_bytes"] < got.timing["h2d_bytes"] or batch.events.shape[0] == 0
    out[0].view(np.uint8)[...] = 0x5A
    gms = ctx.align_batch(pb, out, means=ctx.pin_array(batch.event_means()))
    if batch.n_reads > 0 and int(batch.pair_capacity().sum()) > 0 and batch.events.shape[0] > 0:
        assert gms.timing["streamed"] == 5
    ol.assert_same_alignment(gms, want, what + " (means, streamed)")
    # the batch as db_t holds it (abea_align_ragged): packed / unpacked by host threads while the kernels run
    gr = ctx.align_ragged(batch, threads=4)
    ol.assert_same_alignment(gr, want, what + " (ragged)")
    return got, want


@pytest.mark.parametrize("cfg,n", [("cfg2", 192), ("cfg3", 96), ("cfg4", 24)])
def test_parity_synthetic_configs(ctx, cfg, n):
    b = synth.make_config(cfg, seed=77, n_reads=n)
    got, _ = run_and_check(ctx, b, b.meta["model"], cfg)
    assert (got.n_pairs > 0).mean() > 0.9


def test_parity_rna_r9_5mer(ctx):
    b = synth.make_batch("rna_r9", n_reads=32, mean_events=2500, sigma=0.5, epk=2.2, seed=31)
    run_and_check(ctx, b, "rna_r9", "rna_r9")


def test_parity_short_reads(ctx):
    """Reads shorter than the band: every band is an edge band (validity window, trim column, end column)."""
    b = synth.make_batch("r9", n_reads=256, mean_events=120, sigma=0.9, epk=1.8, seed=32, min_len=12)
    run_and_check(ctx, b, "r9", "short")


def test_parity_edge_cases(ctx):
    b = edge_batch()
    got, _ = run_and_check(ctx, b, "r9", "edge")
    assert got.n_pairs[1] == 0 and got.n_pairs[3] == 0 and got.n_pairs[4] == 0


def test_golden_fixtures(ctx):
    """Committed vectors generated from the reference (tests/golden/make_golden.py): single_read/adaptive.exp and
    a spread of test/ecoli_2kb_region reads."""
    gold = json.load(open(os.path.join(HERE, "golden", "abea_golden.json")))
    npz = np.load(os.path.join(HERE, "golden", "abea_golden.npz"))
    k, m = models.load_model("r9")
    ctx.set_model(m, k)
    b1 = ReadBatch.from_reads([npz["single_seq"].tobytes()], [npz["single_events"]], npz["single_scalings"], k)
    a1 = align_db(ctx, b1)
    assert a1.n_pairs[0] == 7206 and sha(a1.read_pairs(0)) == gold["single_read"]["pairs_sha256"]
    st = ctx.read_stats(1)
    assert abs(st["sum_emission"][0] - gold["single_read"]["golden_sum_emission"]) < 2e-2
    be = ReadBatch(npz["ecoli_seq"], npz["ecoli_seq_ptr"], npz["ecoli_read_len"], npz["ecoli_events"],
                   npz["ecoli_event_ptr"], npz["ecoli_n_events"], npz["ecoli_scalings"],
                   np.ones(len(npz["ecoli_read_len"]), dtype=np.uint8), k)
    ae = align_db(ctx, be)
    assert [int(x) for x in ae.n_pairs] == gold["ecoli"]["n_pairs"]
    assert [sha(ae.read_pairs(i)) for i in range(be.n_reads)] == gold["ecoli"]["pairs_sha256"]
    st = ctx.read_stats(be.n_reads)
    for i, g in enumerate(gold["ecoli"]["golden_sum_emission"]):
        if g is not None:
            assert abs(st["sum_emission"][i] - g) < 5e-2


def test_long_read_no_cpu_fallback(ctx):
    """A read 30x longer than the batch mean (the reference would send it to the CPU, src/f5c.cu:440-452)."""
    short = synth.make_batch("r9", n_reads=31, mean_events=1000, sigma=0.2, epk=1.8, seed=41)
    long_ = synth.make_batch("r9", n_reads=1, mean_events=30000, sigma=0.01, epk=1.8, seed=42)
    seqs = [short.read_seq(i) for i in range(31)] + [long_.read_seq(0)]
    evs = [short.read_events(i) for i in range(31)] + [long_.read_events(0)]
    sc = np.concatenate([short.scalings, long_.scalings])
    b = ReadBatch.from_reads(seqs, evs, sc, short.kmer_size)
    got, _ = run_and_check(ctx, b, "r9", "long")
    assert got.n_pairs[31] > 25000


def check_alignment_properties(batch, aln):
    """Size-independent invariants of a passing alignment (reference src/align.c:452-543)."""
    K = batch.n_kmers
    for i in range(batch.n_reads):
        n = int(aln.n_pairs[i])
        if n == 0:
            continue
        p = aln.read_pairs(i)
        k, e = p["ref_pos"].astype(np.int64), p["read_pos"].astype(np.int64)
        assert k[0] == 0 and k[-1] == K[i] - 1                       # spanned
        assert e.min() >= 0 and e.max() < batch.n_events[i]
        dk, de = np.diff(k), np.diff(e)
        assert np.all((dk >= 0) & (dk <= 1) & (de >= 0) & (de <= 1) & (dk + de >= 1))   # D / U / L steps only
        skip = (de == 0)                                             # runs of FROM_L are bounded by max_gap 50
        if skip.any():
            runs = np.diff(np.flatnonzero(np.diff(np.concatenate(([0], skip.astype(np.int8), [0])))))[::2]
            assert runs.max() <= 50


def test_full_size_cfg2_properties(ctx):
    """BASELINE configs[1] at full size: invariants on all 4096 reads, oracle parity on a random sample, batch
    composition independence, idempotence."""
    b = synth.make_config("cfg2", seed=42)
    k, m = models.load_model("r9")
    m = ctx.set_model(m, k)
    a = align_db(ctx, b)
    assert (a.n_pairs > 0).mean() > 0.98
    check_alignment_properties(b, a)
    a2 = align_db(ctx, b)
    assert np.array_equal(a.n_pairs, a2.n_pairs) and sha(a.pairs[: 1 << 20]) == sha(a2.pairs[: 1 << 20])
    rng = np.random.default_rng(5)
    longest = np.argsort(b.n_events)[-8:]
    idx = np.concatenate([rng.choice(b.n_reads, 56, replace=False), longest])
    sub = b.subset(idx)
    want = ol.port_align(sub, m)
    alone = align_db(ctx, sub)
    ol.assert_same_alignment(alone, want, "cfg2 sample")
    for j, i in enumerate(idx):
        assert np.array_equal(a.read_pairs(int(i)), want.read_pairs(j))


def test_full_size_cfg4_long_read_stress(ctx):
    """BASELINE configs[3] (RNA004, mean 20k events/read) at reduced read count but full read lengths."""
    b = synth.make_config("cfg4", seed=43, n_reads=256)
    k, m = models.load_model("rna004")
    m = ctx.set_model(m, k)
    a = align_db(ctx, b)
    check_alignment_properties(b, a)
    idx = np.argsort(b.n_events)[-4:]
    sub = b.subset(idx)
    want = ol.port_align(sub, m)
    for j, i in enumerate(idx):
        assert np.array_equal(a.read_pairs(int(i)), want.read_pairs(j))


def test_parity_wide_kernel(built, monkeypatch):
    """The opt-in wide fill kernel (one CTA of 4 warps per read) on every read of a batch, long reads and edge cases."""
    monkeypatch.setenv("ABEA_WIDE", "1")   # batches with fewer reads than SMs run entirely wide
    with AbeaContext(0) as wctx:
        b = synth.make_config("cfg3", seed=78, n_reads=64)
        got, _ = run_and_check(wctx, b, "r10", "wide cfg3")
        assert got.timing["n_wide"] >= 32
        run_and_check(wctx, edge_batch(), "r9", "wide edge")
        b = synth.make_batch("r9", n_reads=128, mean_events=150, sigma=0.9, epk=1.8, seed=33, min_len=12)
        run_and_check(wctx, b, "r9", "wide short")


def test_parity_narrow_only(built, monkeypatch):
    """ABEA_WIDE=0: every read through the warp-per-read kernel, including the long-read / secondary-warp logic."""
    monkeypatch.setenv("ABEA_WIDE", "0")
    with AbeaContext(0) as nctx:
        b = synth.make_config("cfg3", seed=79, n_reads=256)
        got, _ = run_and_check(nctx, b, "r10", "narrow cfg3")
        assert got.timing["n_wide"] == 0
        run_and_check(nctx, edge_batch(), "r9", "narrow edge")


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not shipped")
def test_parity_against_reference_object_code(ctx):
    """Same comparison with the unmodified reference align() (oracle/_ref, built from /root/reference)."""
    b = synth.make_config("cfg3", seed=91, n_reads=48)
    k, m = models.load_model("r10")
    m = ctx.set_model(m, k)
    ol.assert_same_alignment(align_db(ctx, b), ol.ref_align(b, m), "cfg3 vs reference")


def full_size_check(ctx, cfg, n_longest, n_random, seed=42):
    """A BASELINE config at FULL size and full batch composition: invariants on every read; the staged, the streamed
    (pinned means) and the ragged path must agree on every read; the oracle on the longest reads (the ones that go
    wide and set the time) plus a random sample."""
    b = synth.make_config(cfg, seed=seed)
    k, m = models.load_model(b.meta["model"])
    m = ctx.set_model(m, k)
    a = align_db(ctx, b)
    assert (a.n_pairs > 0).mean() > 0.95
    check_alignment_properties(b, a)
    pb = ctx.pin_batch(b)
    out = ctx.alloc_output(b, pinned=True)
    s = ctx.align_batch(pb, out, means=ctx.pin_array(b.event_means()))      # the e2e path of bench.py
    assert s.timing["streamed"] == 5 and s.timing["n_wide"] == a.timing["n_wide"]
    assert np.array_equal(s.n_pairs, a.n_pairs)
    for i in range(b.n_reads):
        assert np.array_equal(s.read_pairs(i), a.read_pairs(i)), (cfg, "streamed", i)
    rng = np.random.default_rng(seed + 5)
    longest = np.argsort(b.n_events, kind="stable")[-n_longest:]
    idx = np.concatenate([rng.choice(np.setdiff1d(np.arange(b.n_reads), longest), n_random, replace=False), longest])
    want = ol.port_align(b.subset(idx), m)
    for j, i in enumerate(idx):
        assert int(a.n_pairs[i]) == int(want.n_pairs[j]), (cfg, i)
        assert np.array_equal(a.read_pairs(int(i)), want.read_pairs(j)), (cfg, "oracle", i)
    return b, a


def test_full_size_cfg3_all_reads(ctx):
    """BASELINE configs[2]: R10.4.1, 4096 reads, sigma 1.0 — its 160 k-event tail reads go wide."""
    b, a = full_size_check(ctx, "cfg3", 16, 48)
    assert a.timing["n_wide"] >= 2 and b.n_events.max() > 100000


def test_full_size_cfg4_all_reads(ctx):
    """BASELINE configs[3]: RNA004, 2048 reads, mean 20k events/read."""
    b, a = full_size_check(ctx, "cfg4", 16, 48)
    r = ctx.align_ragged(b, threads=8)
    assert np.array_equal(r.n_pairs, a.n_pairs)
    for i in range(b.n_reads):
        assert np.array_equal(r.read_pairs(i), a.read_pairs(i)), ("ragged", i)


def test_full_size_cfg5_target_config(ctx):
    """The north_star target (BASELINE configs[4] per GPU): R10.4.1, 4096 reads, mean 4k events/read. Also what a
    multi-GPU driver exchanges: the run's path codes in device memory (abea_device_codes), expanded on the device
    (abea_expand_codes), are the pair lists back to back."""
    import torch
    from f5c_b200.dist import compact_pairs
    b, a = full_size_check(ctx, "cfg5", 8, 56)
    _dp, dn, cap, n = ctx.device_results()
    dc, n_words = ctx.device_codes()
    cap_ptr = np.concatenate([[0], np.cumsum(b.pair_capacity().astype(np.int64))])
    assert n == b.n_reads and cap == cap_ptr[-1] and n_words == (cap >> 5) + 2 * n + 2
    dense = torch.zeros((int(cap), 2), dtype=torch.int32, device="cuda")
    total = ctx.expand_codes(dc, dn, cap_ptr, dense.data_ptr(), int(cap), sync=True)
    assert total == int(a.n_pairs.sum())
    got = dense[:total].cpu().numpy().reshape(-1).view(a.pairs.dtype)
    assert np.array_equal(got, compact_pairs(a.pairs, a.pair_ptr, a.n_pairs))


def test_ecoli_all_112_reads_from_blow5(ctx):
    """BASELINE configs[0] on the GPU, every read: the raw signals of tests/golden/ecoli/reads.blow5 (a copy of the
    reference's test/ecoli_2kb_region fixture) through event detection, the scaling estimate, the alignment and the
    recalibration, each stage against what the UNMODIFIED reference produced (tests/golden/ecoli_all.json)."""
    import blow5
    from f5c_b200.abea import scaling_db
    from f5c_b200.batch import EVENT_DTYPE, SCALINGS_DTYPE
    gold = json.load(open(os.path.join(HERE, "golden", "ecoli_all.json")))
    f = blow5.Blow5(os.path.join(HERE, "golden", "ecoli", "reads.blow5"))
    seqs = dict(blow5.read_fasta(os.path.join(HERE, "golden", "ecoli", "reads.fasta")))
    recs = {}
    for i in range(len(f)):
        rid, dig, off, rng, sr, sig = f.read(i)
        recs[rid] = (dig, off, rng, sig)
    names = [r["name"] for r in gold["reads"]]
    assert len(names) == 112
    sig = [recs[n][3] for n in names]
    n_samples = np.array([len(x) for x in sig], dtype=np.int32)
    raw_ptr = np.zeros(len(sig), dtype=np.int64)
    np.cumsum(n_samples[:-1], out=raw_ptr[1:])
    raw = np.concatenate(sig).astype(np.float32)
    cal = tuple(np.array([recs[n][j] for n in names], dtype=np.float32) for j in (1, 2, 0))   # offset, range, digitisation
    k, m = models.load_model("r9")
    ctx.set_model(m, k)
    ev, ev_ptr, nev, _ = ctx.getevents(raw, raw_ptr, n_samples, cal)
    assert [int(x) for x in nev] == [r["n_events"] for r in gold["reads"]]

    def event_sha(e):
        return sha(np.concatenate([e["start"].astype("<u8").view(np.uint8), e["length"].astype("<f4").view(np.uint8),
                                   e["mean"].astype("<f4").view(np.uint8), e["stdv"].astype("<f4").view(np.uint8)]))
    for i, r in enumerate(gold["reads"]):
        assert event_sha(ev[int(ev_ptr[i]):int(ev_ptr[i]) + int(nev[i])]) == r["events_sha256"], ("events", r["name"])
    b = ReadBatch.from_reads([seqs[n].encode() for n in names],
                             [ev[int(ev_ptr[i]):int(ev_ptr[i]) + int(nev[i])] for i in range(len(names))],
                             np.zeros(len(names), dtype=SCALINGS_DTYPE), k)
    ctx.upload(b, with_scalings=False)
    est, _ = ctx.estimate_scalings(b.n_reads)
    assert [int(x) for x in est["shift"].view(np.uint32)] == [r["shift_bits"] for r in gold["reads"]]
    assert [int(x) for x in est["scale"].view(np.uint32)] == [r["scale_bits"] for r in gold["reads"]]
    ctx.run()
    a = ctx.download(b)
    assert [int(x) for x in a.n_pairs] == [r["n_pairs"] for r in gold["reads"]]
    assert [sha(a.read_pairs(i)) for i in range(b.n_reads)] == [r["pairs_sha256"] for r in gold["reads"]]
    sc = scaling_db(ctx, b)
    for i, r in enumerate(gold["reads"]):
        assert int(sc.results["flags"][i]) == r["flags"] and int(sc.results["n_event_alignment"][i]) == r["n_event_alignment"]
        assert int(sc.results["scalings"]["shift"][i:i + 1].view(np.uint32)[0]) == r["recal_shift_bits"], r["name"]
        assert int(sc.results["scalings"]["scale"][i:i + 1].view(np.uint32)[0]) == r["recal_scale_bits"], r["name"]
        if r["map_sha256"] is not None:
            assert sha(sc.read_map(i)) == r["map_sha256"], r["name"]
