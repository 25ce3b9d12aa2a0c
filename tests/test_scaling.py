"""The stages either side of ABEA: N2 estimate_scalings_using_mom (reference src/align.c:58-106) and
N1 scaling_single = postalign + recalibrate_model + read flags (src/f5c.c:736-807, src/align.c:561-773).

CPU ("not gpu") tests pin the oracle: the restatement in oracle/abea_oracle.c against the fixtures generated from the
reference (tests/golden/make_scaling_golden.py — which also checked the reference's own est_scalings.exp,
recalib_scalings.exp and eventalign.summary.exp line by line) and against the reference object code where it is
built; and they run the CUDA kernels' exact control flow on the SIMT emulator. GPU tests are the parity tests proper,
through the C ABI. The bar is bit-exactness: these stages feed integer work (the alignment, the k-mer -> event map)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from edge_cases import edge_batch
from f5c_b200 import models, synth
from f5c_b200.abea import AbeaContext, scaling_db
from f5c_b200.batch import FAILED_ALIGNMENT, FAILED_CALIBRATION, ReadBatch, SCALINGS_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "scaling_golden.json")))
AGOLD = json.load(open(os.path.join(HERE, "golden", "abea_golden.json")))
NPZ = np.load(os.path.join(HERE, "golden", "abea_golden.npz"))
EMU = os.path.join(HERE, "simt", "libabea_emu.so")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return [int(x) for x in np.asarray(a, dtype=np.float32).view(np.uint32)]


def ecoli_batch():
    k, m = models.load_model("r9")
    b = ReadBatch(NPZ["ecoli_seq"], NPZ["ecoli_seq_ptr"], NPZ["ecoli_read_len"], NPZ["ecoli_events"],
                  NPZ["ecoli_event_ptr"], NPZ["ecoli_n_events"], NPZ["ecoli_scalings"],
                  np.ones(len(NPZ["ecoli_read_len"]), dtype=np.uint8), k)
    return b, ol.full_model(m)


def check_against_record(g, est, n_pairs, sc, what):
    """est: SCALINGS_DTYPE[n]; sc: anything with .res / .read_map (oracle_lib.ScalingResult)."""
    r = sc.res
    assert bits(est["shift"]) == g["est_shift_bits"] and bits(est["scale"]) == g["est_scale_bits"], what
    assert [int(x) for x in n_pairs] == g["n_pairs"], what
    assert bits(r["scalings"]["shift"]) == g["shift_bits"] and bits(r["scalings"]["scale"]) == g["scale_bits"], what
    for i in range(len(g["flags"])):
        assert int(r["flags"][i]) == g["flags"][i] and int(r["calibrated"][i]) == g["calibrated"][i], (what, i)
        assert int(r["n_event_alignment"][i]) == g["n_event_alignment"][i], (what, i)
        assert int(r["num_m_state"][i]) == g["num_m_state"][i], (what, i)
        assert float(r["events_per_base"][i]) == g["events_per_base"][i], (what, i)
        if g["calibrated"][i]:
            assert bits(r["scalings"]["var"][i:i + 1])[0] == g["var_bits"][i], (what, i)
            assert bits(r["scalings"]["log_var"][i:i + 1])[0] == g["log_var_bits"][i], (what, i)
        if g["map_sha256"][i] is not None:
            assert sha(sc.read_map(i)) == g["map_sha256"][i], (what, i)


# ---- the oracle is pinned (CPU) ---------------------------------------------------------------------------------

def test_generation_matched_every_reference_expectation():
    """Recorded by make_scaling_golden.py when it ran the reference over all 112 reads of test/ecoli_2kb_region:
    every line of eventalign.summary.exp (by read name, 3 decimals), every distinct line of recalib_scalings.exp and
    est_scalings.exp (2 decimals) was reproduced."""
    m = GOLD["ecoli_all"]
    assert m["n_reads"] == 112 and m["summary_lines_matched"] == 143
    assert m["recalib_matched"] == m["recalib_distinct"] == 111
    assert m["est_matched"] == m["est_distinct"] == 109


def test_port_reproduces_ecoli_fixtures_and_summary_exp():
    b, m = ecoli_batch()
    est = ol.port_estimate_scalings(b, m)
    aln = ol.port_align(b, m)
    sc = ol.port_scaling(b, m, aln, scalings=est)
    check_against_record(GOLD["ecoli"], est, aln.n_pairs, sc, "ecoli")
    for i, g in enumerate(GOLD["ecoli"]["summary_exp"]):   # the reference's own expected output, by read name
        if g is not None and sc.res["calibrated"][i]:
            s = sc.res["scalings"][i]
            assert abs(s["shift"] - g[0]) < 2e-3 and abs(s["scale"] - g[1]) < 2e-3 and abs(s["var"] - g[2]) < 2e-3


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4"])
def test_port_reproduces_synthetic_fixtures(cfg):
    g = GOLD["synthetic_" + cfg]
    b = synth.make_config(cfg, seed=g["seed"], n_reads=g["n_reads"])
    k, m = models.load_model(b.meta["model"])
    m = ol.full_model(m)
    est = ol.port_estimate_scalings(b, m)
    aln = ol.port_align(b, m)
    check_against_record(g, est, aln.n_pairs, ol.port_scaling(b, m, aln), cfg)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name,kw,min_events", [
    ("r9", dict(n_reads=24, mean_events=2500, sigma=0.6, epk=1.8, seed=21), 200),
    ("r10", dict(n_reads=12, mean_events=3000, sigma=1.0, epk=1.9, seed=22), 200),
    ("rna004", dict(n_reads=6, mean_events=5000, sigma=0.5, epk=2.5, seed=23), 200),
    ("r9", dict(n_reads=16, mean_events=90, sigma=0.9, epk=1.8, seed=25, min_len=12), 20),
])
def test_port_equals_reference_object_code(name, kw, min_events):
    b = synth.make_batch(name, **kw)
    k, m = models.load_model(name)
    m = ol.full_model(m)
    assert ol.port_estimate_scalings(b, m).tobytes() == ol.ref_estimate_scalings(b, m).tobytes()
    aln = ol.ref_align(b, m)
    ol.assert_same_scaling(ol.port_scaling(b, m, aln, min_events=min_events),
                           ol.ref_scaling(b, m, aln, min_events=min_events), name, check_var_d=False)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_equals_reference_on_edge_cases():
    b = edge_batch()
    k, m = models.load_model("r9")
    m = ol.full_model(m)
    aln = ol.ref_align(b, m)
    ol.assert_same_scaling(ol.port_scaling(b, m, aln, min_events=50), ol.ref_scaling(b, m, aln, min_events=50), "edge",
                           check_var_d=False)


# ---- the CUDA path --------------------------------------------------------------------------------------------------

def device_stages(ctx, batch, model_name, min_events=200, rna_signal_order=False):
    """upload without scalings -> estimate on the device -> align -> scaling stage; returns everything."""
    k, m = models.load_model(model_name)
    m = ctx.set_model(m, k)
    b = batch
    if rna_signal_order:   # hand the device the events in signal order, as event_single has them before the reversal
        evs = [batch.read_events(i)[::-1].copy() for i in range(batch.n_reads)]
        b = ReadBatch.from_reads([batch.read_seq(i) for i in range(batch.n_reads)], evs, batch.scalings.copy(),
                                 batch.kmer_size, good=batch.good.copy())
    ctx.upload(b, with_scalings=False)
    est, t_est = ctx.estimate_scalings(b.n_reads, reverse_events=rna_signal_order)
    ctx.run()
    aln = ctx.download(batch)
    sc = scaling_db(ctx, batch, min_events)
    return m, b, est, aln, sc


def check_device_stages(ctx, batch, model_name, what, min_events=200, rna_signal_order=False):
    m, b_in, est, aln, sc = device_stages(ctx, batch, model_name, min_events, rna_signal_order)
    usable = (batch.good != 0) & (batch.n_events >= 1) & (batch.read_len >= batch.kmer_size)
    want_est = np.zeros(batch.n_reads, dtype=SCALINGS_DTYPE)
    idx = np.flatnonzero(usable)
    if len(idx):
        want_est[idx] = ol.port_estimate_scalings(b_in.subset(idx), m)   # MoM sums run over the events as uploaded
    assert est.tobytes() == want_est.tobytes(), what + ": estimate_scalings_using_mom"
    with_est = ReadBatch(batch.seq, batch.seq_ptr, batch.read_len, batch.events, batch.event_ptr, batch.n_events,
                         want_est, batch.good, batch.kmer_size)
    want_aln = ol.port_align(with_est, m)
    ol.assert_same_alignment(aln, want_aln, what)
    want_sc = ol.port_scaling(with_est, m, want_aln, min_events=min_events)
    ol.assert_same_scaling(ol.ScalingResult(batch, sc.results, sc.maps), want_sc, what)
    return est, aln, sc


@pytest.fixture(scope="module")
def emu():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "simt")])
    return EMU


@pytest.mark.parametrize("name,kw,min_events", [
    ("r9", dict(n_reads=10, mean_events=600, sigma=0.6, epk=1.8, seed=21), 200),
    ("r10", dict(n_reads=5, mean_events=700, sigma=1.0, epk=1.9, seed=22), 100),
    ("r9", dict(n_reads=16, mean_events=90, sigma=0.9, epk=1.8, seed=25, min_len=12), 20),
])
def test_emulated_stages_match_oracle(emu, name, kw, min_events):
    with AbeaContext(0, lib_path=emu) as ctx:
        est, aln, sc = check_device_stages(ctx, synth.make_batch(name, **kw), name, name, min_events)
        assert sc.results["calibrated"].any()


def test_emulated_rna_reversal_and_edge_cases(emu):
    with AbeaContext(0, lib_path=emu) as ctx:
        b = synth.make_batch("rna_r9", n_reads=4, mean_events=500, sigma=0.4, epk=2.2, seed=24)
        check_device_stages(ctx, b, "rna_r9", "rna", 100, rna_signal_order=True)
        est, aln, sc = check_device_stages(ctx, edge_batch(), "r9", "edge", 50)
        assert sc.results["flags"][4] == FAILED_ALIGNMENT        # bad read: never aligned
        assert sc.results["flags"][3] == FAILED_ALIGNMENT        # over-segmented: filtered before ABEA


def test_emulated_fixture_reads(emu):
    """Two committed ecoli reads through the emulated kernels against the reference's recorded outputs."""
    b, m = ecoli_batch()
    order = np.argsort(b.n_events)[:2]
    sub = b.subset(order)
    with AbeaContext(0, lib_path=emu) as ctx:
        mm, b_in, est, aln, sc = device_stages(ctx, sub, "r9")
    g = GOLD["ecoli"]
    for j, i in enumerate(order):
        i = int(i)
        assert bits(est["shift"][j:j + 1])[0] == g["est_shift_bits"][i]
        assert int(aln.n_pairs[j]) == g["n_pairs"][i]
        assert int(sc.results["flags"][j]) == g["flags"][i]
        assert bits(sc.results["scalings"]["shift"][j:j + 1])[0] == g["shift_bits"][i]
        if g["map_sha256"][i] is not None:
            assert sha(sc.read_map(j)) == g["map_sha256"][i]


def test_run_without_scalings_is_refused(emu):
    from f5c_b200.abea import AbeaError
    b = synth.make_batch("r9", n_reads=2, mean_events=100, sigma=0.2, epk=1.8, seed=3)
    k, m = models.load_model("r9")
    with AbeaContext(0, lib_path=emu) as ctx:
        ctx.set_model(m, k)
        ctx.upload(b, with_scalings=False)
        with pytest.raises(AbeaError):
            ctx.run()
        with pytest.raises(AbeaError):
            ctx.scaling_stage()


@pytest.fixture(scope="module")
def gctx(built):
    c = AbeaContext(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n", [("cfg2", 192), ("cfg3", 96), ("cfg4", 24)])
def test_gpu_stages_synthetic_configs(gctx, cfg, n):
    b = synth.make_config(cfg, seed=77, n_reads=n)
    est, aln, sc = check_device_stages(gctx, b, b.meta["model"], cfg)
    assert (sc.results["flags"] == 0).mean() > 0.8


@pytest.mark.gpu
def test_gpu_stages_rna_reversal_short_and_edge(gctx):
    b = synth.make_batch("rna_r9", n_reads=32, mean_events=2500, sigma=0.5, epk=2.2, seed=31)
    check_device_stages(gctx, b, "rna_r9", "rna_r9", rna_signal_order=True)
    b = synth.make_batch("r9", n_reads=256, mean_events=120, sigma=0.9, epk=1.8, seed=32, min_len=12)
    check_device_stages(gctx, b, "r9", "short", min_events=20)
    est, aln, sc = check_device_stages(gctx, edge_batch(), "r9", "edge", min_events=50)
    assert sc.results["flags"][4] == FAILED_ALIGNMENT and sc.results["flags"][3] == FAILED_ALIGNMENT


@pytest.mark.gpu
def test_gpu_stages_golden_fixtures(gctx):
    """The committed ecoli reads: device estimate, alignment and recalibration against the reference's recorded
    outputs, and against eventalign.summary.exp (3 decimals) for the reads it lists."""
    b, m = ecoli_batch()
    mm, b_in, est, aln, sc = device_stages(gctx, b, "r9")
    check_against_record(GOLD["ecoli"], est, aln.n_pairs, ol.ScalingResult(b, sc.results, sc.maps), "ecoli on device")
    for i, g in enumerate(GOLD["ecoli"]["summary_exp"]):
        if g is not None and sc.results["calibrated"][i]:
            s = sc.results["scalings"][i]
            assert abs(s["shift"] - g[0]) < 2e-3 and abs(s["scale"] - g[1]) < 2e-3 and abs(s["var"] - g[2]) < 2e-3


@pytest.mark.gpu
def test_gpu_stages_full_size_cfg2(gctx):
    """BASELINE configs[1] at full size: oracle parity of the estimate on every read (cheap on the CPU), of the whole
    chain on a sample, and size-independent properties of the map on all 4096 reads."""
    b = synth.make_config("cfg2", seed=42)
    m, b_in, est, aln, sc = device_stages(gctx, b, "r9")
    assert est.tobytes() == ol.port_estimate_scalings(b, m).tobytes()
    assert np.array_equal(est["shift"], b.scalings["shift"]) and np.array_equal(est["scale"], b.scalings["scale"])
    r = sc.results
    K = b.n_kmers
    for i in range(b.n_reads):
        n = int(aln.n_pairs[i])
        if n == 0:
            assert r["flags"][i] == FAILED_ALIGNMENT
            continue
        mp = sc.read_map(i)
        has = mp["start"] >= 0
        assert np.all(mp["stop"][has] >= mp["start"][has]) and np.all(mp["stop"][~has] == -1)
        st, sp = mp["start"][has].astype(np.int64), mp["stop"][has].astype(np.int64)
        assert np.all(st[1:] == sp[:-1] + 1)                     # event ranges of consecutive k-mers tile the path
        assert int((sp - st + 1).sum()) == r["n_event_alignment"][i]
        p = aln.read_pairs(i)
        assert st[0] == p["read_pos"][0] and sp[-1] == p["read_pos"][-1]
        assert r["events_per_base"][i] == (int(p["read_pos"].max()) - int(p["read_pos"].min())) / K[i]
        assert 0 < r["num_m_state"][i] <= has.sum()
    idx = np.concatenate([np.random.default_rng(6).choice(b.n_reads, 40, replace=False), np.argsort(b.n_events)[-4:]])
    sub = b.subset(idx)
    want_aln = ol.port_align(sub, m)
    want = ol.port_scaling(sub, m, want_aln)
    for j, i in enumerate(idx):
        i = int(i)
        assert r[i].tobytes() == want.res[j].tobytes(), i
        if want.res["n_event_alignment"][j] > 0:
            assert np.array_equal(sc.read_map(i), want.read_map(j))
    assert (r["flags"] & FAILED_CALIBRATION).mean() < 0.05


def test_emulated_stages_with_many_skips(emu):
    """15 % of the k-mers without any event: hundreds of skip steps (pairs that repeat the previous event), runs of
    them, and k-mers whose only pair was reached by a skip — the cases postalign's map logic has to get right."""
    b = synth.make_batch("r9", n_reads=8, mean_events=500, sigma=0.4, epk=1.8, seed=5, p_skip=0.15)
    with AbeaContext(0, lib_path=emu) as ctx:
        est, aln, sc = check_device_stages(ctx, b, "r9", "skips", min_events=100)
    skips = empty = 0
    for i in range(b.n_reads):
        p = aln.read_pairs(i)
        skips += int(((np.diff(p["read_pos"]) == 0) & (np.diff(p["ref_pos"]) == 1)).sum())
        empty += int((sc.read_map(i)["start"] == -1).sum()) if aln.n_pairs[i] > 0 else 0
    assert skips > 50 and empty > 20



def test_emulated_scaling_download_layouts_and_device_results(emu):
    """abea_scaling_download with a caller layout that is not the canonical prefix sum (per-read copies), and
    abea_scaling_device_results (on the emulator the 'device' pointers are host pointers and can be read back)."""
    import ctypes
    from f5c_b200.batch import INDEX_PAIR_DTYPE, SCALING_RESULT_DTYPE
    b = synth.make_batch("r9", n_reads=5, mean_events=300, sigma=0.3, epk=1.8, seed=8)
    k, m = models.load_model("r9")
    with AbeaContext(0, lib_path=emu) as ctx:
        ctx.set_model(m, k)
        ctx.upload(b)
        ctx.run()
        ctx.scaling_stage(100)
        canon = ctx.scaling_download(b)
        mp = b.map_ptr()
        gap = 7                                              # every read's map 7 entries further apart
        mp2 = mp[:-1] + gap * np.arange(b.n_reads, dtype=np.int64)
        res = np.zeros(b.n_reads, dtype=SCALING_RESULT_DTYPE)
        maps = np.full(2 * (int(mp[-1]) + gap * b.n_reads), -7, dtype=np.int32).view(INDEX_PAIR_DTYPE)
        ctx._check(ctx.lib.abea_scaling_download(ctx._h, res.ctypes.data, maps.ctypes.data, mp2.ctypes.data), "download")
        assert res.tobytes() == canon.results.tobytes()
        K = b.n_kmers
        for i in range(b.n_reads):
            if canon.results["n_event_alignment"][i] > 0:
                assert np.array_equal(maps[int(mp2[i]):int(mp2[i]) + int(K[i])], canon.read_map(i))
        dres, dmaps, total = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        ctx._check(ctx.lib.abea_scaling_device_results(ctx._h, ctypes.byref(dres), ctypes.byref(dmaps), ctypes.byref(total)),
                   "device results")
        assert total.value == int(mp[-1])
        raw = (ctypes.c_char * (SCALING_RESULT_DTYPE.itemsize * b.n_reads)).from_address(dres.value)
        on_dev = np.frombuffer(raw, dtype=SCALING_RESULT_DTYPE)
        for f in ("flags", "n_event_alignment", "num_m_state", "events_per_base", "var_d"):
            assert np.array_equal(on_dev[f], canon.results[f]), f       # (log_var is filled by the download, on the host)
