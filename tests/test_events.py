"""Event detection: the reference's getevents (src/events.c:562-582) + event_single's pA
conversion (src/f5c.c:692-696).

CPU tests pin the oracle restatement (oracle/abea_oracle.c abea_oracle_getevents) to the reference's golden event
table test/ecoli_2kb_region/single_read/read1.events.exp and to committed real signals whose event tables reproduce
adaptive.exp (tests/golden/make_events_golden.py), compare it with the reference object code where that is built,
and run the CUDA kernel's control flow on the SIMT emulator. GPU tests are the parity tests proper, through the C
ABI. Bar: bit-exact event tables (start, length, mean, stdv)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from f5c_b200 import synth
from f5c_b200.abea import AbeaContext

HERE = os.path.dirname(os.path.abspath(__file__))
SIG = np.load(os.path.join(HERE, "golden", "events_golden.npz"))
NPZ = np.load(os.path.join(HERE, "golden", "abea_golden.npz"))
EMU = os.path.join(HERE, "simt", "libabea_emu.so")


def to_pa(sig, cal):
    """event_single, reference src/f5c.c:692-696 (all float)."""
    off, rng, dig = (np.float32(x) for x in cal)
    return ((sig.astype(np.float32) + off) * np.float32(rng / dig)).astype(np.float32)


def fixture_reads():
    """(int16 signal, (offset, range, digitisation), expected events) of the committed ecoli reads."""
    out = []
    for j in range(3):
        i = int(SIG[f"idx{j}"][0])
        p, n = int(NPZ["ecoli_event_ptr"][i]), int(NPZ["ecoli_n_events"][i])
        out.append((SIG[f"sig{j}"], SIG[f"cal{j}"], NPZ["ecoli_events"][p:p + n]))
    return out


def check_single_read(ev):
    """test/ecoli_2kb_region/single_read/read1.events.exp: 7165 events printed with %f."""
    g = NPZ["single_events"]
    assert len(ev) == len(g) == 7165
    assert np.array_equal(ev["start"], g["start"]) and np.array_equal(ev["length"], g["length"])
    assert np.abs(ev["mean"].astype(np.float64) - g["mean"]).max() < 1e-6
    assert np.abs(ev["stdv"].astype(np.float64) - g["stdv"]).max() < 1e-6


# ---- the oracle is pinned (CPU) ---------------------------------------------------------------------------------

def test_port_reproduces_reference_golden_event_table():
    check_single_read(ol.port_getevents(to_pa(SIG["single_sig"], SIG["single_cal"])))


def test_port_reproduces_committed_event_tables():
    for sig, cal, want in fixture_reads():
        assert ol._events_equal(ol.port_getevents(to_pa(sig, cal)), want)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("rna", [False, True])
def test_port_equals_reference_object_code(rna):
    sg = synth.make_signals(24, 1500, 0.7, seed=7 + rna, samples_per_event=12.0 if rna else 5.0)
    for i in range(24):
        pa = sg["pa"][sg["raw_ptr"][i]:sg["raw_ptr"][i] + sg["n_samples"][i]]
        assert ol._events_equal(ol.port_getevents(pa, rna), ol.ref_getevents(pa, rna)), i
    flat = np.full(500, 80.0, dtype=np.float32)          # a signal without a single boundary: one event here,
    assert len(ol.port_getevents(flat)) == 1              # undefined in the reference (it reads peaks[-1])
    assert len(ol.port_getevents(flat[:50])) == 0         # < 100 samples: the reference asserts


# ---- the CUDA path --------------------------------------------------------------------------------------------------

def check_device(ctx, sg, rna=False, calibrated=True):
    if calibrated:
        ev, ptr, nev, t = ctx.getevents(sg["raw"], sg["raw_ptr"], sg["n_samples"],
                                        (sg["offset"], sg["range"], sg["digitisation"]), rna=rna)
    else:
        ev, ptr, nev, t = ctx.getevents(sg["pa"], sg["raw_ptr"], sg["n_samples"], None, rna=rna)
    for i in range(len(nev)):
        pa = sg["pa"][sg["raw_ptr"][i]:sg["raw_ptr"][i] + sg["n_samples"][i]]
        want = ol.port_getevents(pa, rna)
        assert ol._events_equal(ev[ptr[i]:ptr[i] + nev[i]], want), (i, int(nev[i]), len(want))
    return ev, ptr, nev, t


def fixture_signals():
    reads = fixture_reads()
    sigs = [r[0] for r in reads] + [SIG["single_sig"]]
    cals = [r[1] for r in reads] + [SIG["single_cal"]]
    n = np.array([len(s) for s in sigs], dtype=np.int32)
    ptr = np.zeros(len(sigs), dtype=np.int64)
    np.cumsum(n[:-1], out=ptr[1:])
    raw = np.concatenate(sigs).astype(np.float32)
    off, rng, dig = (np.array([c[k] for c in cals], dtype=np.float32) for k in range(3))
    return raw, ptr, n, (off, rng, dig), [r[2] for r in reads]


def check_fixture_on(ctx):
    raw, ptr, n, cal, want = fixture_signals()
    ev, eptr, nev, t = ctx.getevents(raw, ptr, n, cal)
    for i, w in enumerate(want):
        assert ol._events_equal(ev[eptr[i]:eptr[i] + nev[i]], w), i
    check_single_read(ev[eptr[3]:eptr[3] + nev[3]])


@pytest.fixture(scope="module")
def emu():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "simt")])
    return EMU


@pytest.mark.parametrize("rna,chunk", [(False, None), (True, None), (False, "8"), (False, "64"), (True, "32")])
def test_emulated_kernel_matches_oracle(emu, monkeypatch, rna, chunk):
    """chunk: ABEA_EVT_CHUNK — tiny chunks make every read a long chain of speculative chunks, most of which never
    synchronise with the true walk (8 samples rarely contain a boundary), so both stitch paths are exercised."""
    if chunk:
        monkeypatch.setenv("ABEA_EVT_CHUNK", chunk)
    sg = synth.make_signals(9, 400, 0.6, seed=11 + rna, samples_per_event=12.0 if rna else 5.0)
    sg["n_samples"][3] = 60                                  # shorter than 100 samples: no events
    with AbeaContext(0, lib_path=emu) as ctx:
        ev, ptr, nev, t = check_device(ctx, sg, rna, calibrated=True)
        assert nev[3] == 0 and (np.delete(nev, 3) > 20).all()
        check_device(ctx, sg, rna, calibrated=False)


@pytest.mark.parametrize("chunk", [None, "48"])
def test_emulated_kernel_on_real_signals(emu, monkeypatch, chunk):
    if chunk:
        monkeypatch.setenv("ABEA_EVT_CHUNK", chunk)
    with AbeaContext(0, lib_path=emu) as ctx:
        check_fixture_on(ctx)


@pytest.fixture(scope="module")
def gctx(built):
    c = AbeaContext(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("rna,n,mean", [(False, 256, 4000), (True, 64, 6000), (False, 512, 150)])
def test_gpu_getevents_synthetic(gctx, rna, n, mean):
    sg = synth.make_signals(n, mean, 0.6, seed=21 + rna, samples_per_event=12.0 if rna else 5.0)
    check_device(gctx, sg, rna, calibrated=True)
    check_device(gctx, sg, rna, calibrated=False)


@pytest.mark.gpu
def test_gpu_getevents_golden_signals(gctx):
    check_fixture_on(gctx)


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", ["16", "200"])
def test_gpu_getevents_small_chunks(built, monkeypatch, chunk):
    monkeypatch.setenv("ABEA_EVT_CHUNK", chunk)
    with AbeaContext(0) as ctx:
        check_device(ctx, synth.make_signals(128, 2000, 0.6, seed=29), False, calibrated=True)
        check_fixture_on(ctx)


@pytest.mark.gpu
def test_gpu_getevents_full_size_properties(gctx):
    """4096 signals of cfg2's event counts (~81 M samples): size-independent properties on every read, oracle parity
    on a sample, and the events feed the alignment."""
    sg = synth.make_signals(4096, 4000, 0.5, seed=42)
    ev, ptr, nev, t = gctx.getevents(sg["raw"], sg["raw_ptr"], sg["n_samples"],
                                     (sg["offset"], sg["range"], sg["digitisation"]))
    assert (nev > 0).all()
    end = ev["start"].astype(np.int64) + ev["length"].astype(np.int64)
    first = ptr
    last = ptr + nev - 1
    assert (ev["start"][first] == 0).all() and np.array_equal(end[last], sg["n_samples"].astype(np.int64))
    inner = np.ones(len(ev), dtype=bool)
    inner[last] = False
    assert np.array_equal(end[inner], ev["start"][1:][inner[:-1]].astype(np.int64))   # events tile the signal
    assert (ev["length"] >= 1).all() and (ev["stdv"] >= 0).all()
    # sum over events of mean * length == sum of samples (within float accumulation of the check itself)
    i = int(np.argmax(nev))
    pa = sg["pa"][sg["raw_ptr"][i]:sg["raw_ptr"][i] + sg["n_samples"][i]]
    e = ev[ptr[i]:ptr[i] + nev[i]]
    assert abs(float((e["mean"].astype(np.float64) * e["length"]).sum()) - float(pa.astype(np.float64).sum())) < 1e-2 * len(e)
    for i in list(np.random.default_rng(3).choice(4096, 24, replace=False)) + [int(np.argmax(nev))]:
        pa = sg["pa"][sg["raw_ptr"][i]:sg["raw_ptr"][i] + sg["n_samples"][i]]
        assert ol._events_equal(ev[ptr[i]:ptr[i] + nev[i]], ol.port_getevents(pa)), i


# ---- raw signal in, alignment and recalibration out, everything resident on the device ---------------------------

def check_resident_chain(ctx, model_name, n_reads, mean_kmers, seed, min_events=50):
    """abea_getevents -> abea_upload_batch(events = NULL, scalings = NULL) -> abea_estimate_scalings -> abea_run ->
    abea_scaling_stage against the oracle's getevents -> estimate_scalings_using_mom -> align -> scaling_single."""
    from f5c_b200 import models
    from f5c_b200.abea import scaling_db
    from f5c_b200.batch import EVENT_DTYPE, SCALINGS_DTYPE, ReadBatch
    sg, seq, seq_ptr, read_len, k = synth.make_signal_batch(model_name, n_reads, mean_kmers, 0.5, seed)
    kk, m = models.load_model(model_name)
    m = ctx.set_model(m, kk)
    cal = (sg["offset"], sg["range"], sg["digitisation"])
    _, _, nev, t = ctx.getevents(sg["raw"], sg["raw_ptr"], sg["n_samples"], cal, download=False)
    assert (nev > 0).all()
    shell = ReadBatch(seq, seq_ptr, read_len, np.zeros(0, dtype=EVENT_DTYPE), np.zeros(n_reads, dtype=np.int64),
                      nev.astype(np.int32), np.zeros(n_reads, dtype=SCALINGS_DTYPE), np.ones(n_reads, dtype=np.uint8), k)
    ctx.upload(shell, with_scalings=False, device_events=True)
    est, _ = ctx.estimate_scalings(n_reads)
    ctx.run()
    aln = ctx.download(shell)
    sc = scaling_db(ctx, shell, min_events)
    # the oracle's chain on the host
    evs = [ol.port_getevents(sg["pa"][sg["raw_ptr"][i]:sg["raw_ptr"][i] + sg["n_samples"][i]]) for i in range(n_reads)]
    assert [len(e) for e in evs] == [int(x) for x in nev]
    seqs = [seq[seq_ptr[i]:seq_ptr[i] + read_len[i]].tobytes() for i in range(n_reads)]
    hb = ReadBatch.from_reads(seqs, evs, np.zeros(n_reads, dtype=SCALINGS_DTYPE), k)
    hb.scalings[:] = ol.port_estimate_scalings(hb, m)
    assert est.tobytes() == hb.scalings.tobytes()
    want = ol.port_align(hb, m)
    ol.assert_same_alignment(aln, want, "resident chain")
    ol.assert_same_scaling(ol.ScalingResult(hb, sc.results, sc.maps), ol.port_scaling(hb, m, want, min_events=min_events),
                           "resident chain")
    return aln, sc


def test_emulated_resident_chain(emu):
    with AbeaContext(0, lib_path=emu) as ctx:
        aln, sc = check_resident_chain(ctx, "r9", 6, 250, seed=5)
        assert (aln.n_pairs > 0).sum() >= 4      # signals generated from the sequences do align


@pytest.mark.gpu
@pytest.mark.parametrize("model,n,mean", [("r9", 192, 2500), ("r10", 64, 3000)])
def test_gpu_resident_chain(gctx, model, n, mean):
    aln, sc = check_resident_chain(gctx, model, n, mean, seed=17, min_events=200)
    assert (aln.n_pairs > 0).mean() > 0.8 and (sc.results["flags"] == 0).mean() > 0.6


def adversarial_signals():
    """Signals that stress the speculative detector's stitching: long flat stretches (chunks with no boundary at all,
    so the re-walk never meets the speculative walk), boundaries every two samples, repeated identical values, a
    staircase, heavy noise, a jump exactly at a chunk edge, magnitudes 1e-6 and 150 mixed (inexact cumulative sums),
    zeros and negative values."""
    rng = np.random.default_rng(99)
    sigs = []
    sigs.append(np.concatenate([np.full(900, 80.0), 80 + 0.01 * rng.standard_normal(700), np.full(50, 120.0), np.full(1200, 95.0)]))
    sigs.append(np.repeat(rng.normal(90, 15, 600), 2) + 0.3 * rng.standard_normal(1200))
    sigs.append(np.tile(np.array([70.0, 70.0, 70.0, 110.0, 110.0, 110.0]), 300))
    sigs.append(np.repeat(np.arange(60, 140, 0.5), 9).astype(np.float64))
    sigs.append(90 + 25 * rng.standard_normal(3000))
    s = np.full(2048, 85.0) + 0.5 * rng.standard_normal(2048)
    s[1024:] += 30.0
    sigs.append(s)
    # cumulative sums whose additions DO round (1e-6 next to 150: 62 bits needed) -> the ordered-chain path of the
    # sums kernel; the signals before this one all take its exact warp-scan path
    sigs.append(np.where(rng.random(1800) < 0.3, 1e-6 * (1 + rng.random(1800)), 150 + 12 * rng.standard_normal(1800)))
    sigs.append(np.concatenate([np.zeros(300), -40 + 5 * rng.standard_normal(900), np.zeros(200), 60 + 3 * rng.standard_normal(700)]))
    sigs.append(np.full(1500, 77.25))
    n = np.array([len(x) for x in sigs], dtype=np.int32)
    ptr = np.zeros(len(sigs), dtype=np.int64)
    np.cumsum(n[:-1], out=ptr[1:])
    pa = np.concatenate(sigs).astype(np.float32)
    return dict(pa=pa, raw=pa, raw_ptr=ptr, n_samples=n)


@pytest.mark.parametrize("chunk", ["8", "64", "1024"])
def test_emulated_kernel_adversarial_signals(emu, monkeypatch, chunk):
    monkeypatch.setenv("ABEA_EVT_CHUNK", chunk)
    sg = adversarial_signals()
    with AbeaContext(0, lib_path=emu) as ctx:
        for rna in (False, True):
            ev, ptr, nev, t = check_device(ctx, sg, rna, calibrated=False)
        assert nev[-1] == 1                      # a constant signal: one event (undefined in the reference)


@pytest.mark.gpu
def test_gpu_getevents_adversarial_signals(gctx):
    sg = adversarial_signals()
    for rna in (False, True):
        check_device(gctx, sg, rna, calibrated=False)


def test_exact_sum_condition_is_what_makes_the_scan_legal():
    """The sums kernel replaces the reference's ordered additions by a warp scan when
    (e_max + 1 + ceil(log2 n)) - (e_min - 23) <= 53 (and the analogous bound for the float squares). Under that
    condition every partial sum is exactly representable, so ANY association gives the sequential result; outside it
    the association matters (which is why the kernel then keeps the reference's order)."""
    rng = np.random.default_rng(8)
    x = rng.uniform(16.0, 256.0, 200000).astype(np.float32)             # e_min = 4, e_max = 7, n < 2^18: 49 bits
    assert (7 + 1 + 18) - (4 - 23) <= 53
    seq = np.cumsum(x.astype(np.float64))                                 # sequential, as compute_sum_sumsq
    blocks = x.astype(np.float64).reshape(-1, 100).sum(axis=1)            # a different association
    assert np.array_equal(np.cumsum(blocks), seq[99::100])
    assert float(np.sum(x.astype(np.float64)[::-1])) == float(seq[-1])    # reversed order
    sq = (x * x).astype(np.float64)                                       # float product, promoted afterwards
    assert (2 * 7 + 2 + 18) - (2 * 4 - 23) <= 53
    assert float(np.sum(sq[::-1])) == float(np.cumsum(sq)[-1])
    y = np.where(rng.random(200000) < 0.3, 1e-6, 150.0).astype(np.float32) * rng.uniform(1, 2, 200000).astype(np.float32)
    assert (8 + 1 + 18) - (-20 - 23) > 53                                 # 1e-6 next to 150: additions round
    assert float(np.sum(y.astype(np.float64)[::-1])) != float(np.cumsum(y.astype(np.float64))[-1])


def test_getevents_argument_errors(emu):
    from f5c_b200.abea import AbeaError
    with AbeaContext(0, lib_path=emu) as ctx:
        sg = synth.make_signals(2, 200, 0.3, seed=1)
        with pytest.raises(AbeaError):       # offset without range / digitisation
            import ctypes
            from f5c_b200.abea import CSignals, Timing
            cs = CSignals(2, sg["raw"].ctypes.data, sg["raw_ptr"].ctypes.data, sg["n_samples"].ctypes.data,
                          sg["offset"].ctypes.data, None, None)
            nev = np.zeros(2, dtype=np.int32)
            ctx._check(ctx.lib.abea_getevents(ctx._h, ctypes.byref(cs), 0, nev.ctypes.data, None), "abea_getevents")
        ev, ptr, nev, t = ctx.getevents(np.zeros(0, dtype=np.float32), np.zeros(0, dtype=np.int64),
                                        np.zeros(0, dtype=np.int32))          # an empty batch is fine
        assert len(ev) == 0 and len(nev) == 0
        from f5c_b200 import models
        from f5c_b200.batch import EVENT_DTYPE, SCALINGS_DTYPE, ReadBatch
        k, m = models.load_model("r9")
        ctx.set_model(m, k)
        b = synth.make_batch("r9", n_reads=2, mean_events=100, sigma=0.2, epk=1.8, seed=3)
        with pytest.raises(AbeaError):       # events == NULL, but the device holds no tables for these reads
            ctx.upload(b, device_events=True)
