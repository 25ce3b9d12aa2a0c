"""The reference-facing drop-in (f5c_b200/csrc/f5c_dropin.cu, compiled against the reference's own f5c.h): exports the
reference's C++-linkage entry points, and — on the GPU — init_cuda / align_cuda / free_cuda driven through real
core_t / db_t structs give the oracle's pairs."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from edge_cases import edge_batch
from f5c_b200 import models, synth
from f5c_b200.batch import CBatch, PAIR_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "f5c_b200", "lib", "libf5c_abea_dropin.so")
needs_so = pytest.mark.skipif(not os.path.exists(SO), reason="drop-in not built (needs the f5c tree at build time)")


@needs_so
def test_dropin_exports_reference_symbols():
    out = subprocess.check_output(["nm", "-D", "--defined-only", SO]).decode()
    # mangled names of void align_cuda(core_t*, db_t*), void init_cuda(core_t*), void free_cuda(core_t*)
    for sym in ("_Z10align_cudaP6core_tP4db_t", "_Z9init_cudaP6core_t", "_Z9free_cudaP6core_t"):
        assert sym in out, sym


@needs_so
@pytest.mark.gpu
@pytest.mark.parametrize("which", ["synthetic", "edge"])
def test_dropin_align_cuda_matches_oracle(built, which):
    lib = ctypes.CDLL(SO)
    lib.f5c_dropin_selftest.argtypes = [ctypes.POINTER(CBatch), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    b = synth.make_config("cfg2", seed=61, n_reads=64) if which == "synthetic" else edge_batch()
    k, m = models.load_model("r9")
    m = ol.full_model(m)
    pairs = np.zeros(int(b.pair_capacity().sum()), dtype=PAIR_DTYPE)
    n_pairs = np.full(b.n_reads, -1, dtype=np.int32)
    pp = b.pair_ptr()
    cb = b.as_c()
    assert lib.f5c_dropin_selftest(ctypes.byref(cb), m.ctypes.data, k, 0, pairs.ctypes.data, pp.ctypes.data,
                                   n_pairs.ctypes.data) == 0
    got = ol.AlignResult(b, pairs, n_pairs)
    ol.assert_same_alignment(got, ol.port_align(b, m), "drop-in " + which)


@needs_so
def test_dropin_exports_scaling_entry_points():
    out = subprocess.check_output(["nm", "-D", "--defined-only", SO]).decode()
    for sym in ("_Z12scaling_cudaP6core_tP4db_t", "_Z22estimate_scalings_cudaP6core_tP4db_t"):
        assert sym in out, sym


@needs_so
@pytest.mark.gpu
@pytest.mark.parametrize("which", ["synthetic", "edge"])
def test_dropin_estimate_align_scaling_matches_oracle(built, which):
    """estimate_scalings_cuda -> align_cuda -> scaling_cuda on real core_t / db_t against the oracle's
    estimate_scalings_using_mom -> align -> scaling_single."""
    from f5c_b200.batch import INDEX_PAIR_DTYPE, SCALING_RESULT_DTYPE, ReadBatch
    lib = ctypes.CDLL(SO)
    vp = ctypes.c_void_p
    lib.f5c_dropin_selftest_scaling.argtypes = [ctypes.POINTER(CBatch), vp, ctypes.c_uint32, ctypes.c_int,
                                                ctypes.c_int32, vp, vp, vp, vp]
    b = synth.make_config("cfg2", seed=62, n_reads=64) if which == "synthetic" else edge_batch()
    min_events = 200 if which == "synthetic" else 50
    k, m = models.load_model("r9")
    m = ol.full_model(m)
    mp = b.map_ptr()
    n_pairs = np.full(b.n_reads, -1, dtype=np.int32)
    res = np.zeros(b.n_reads, dtype=SCALING_RESULT_DTYPE)
    maps = np.full(2 * int(mp[-1]), -1, dtype=np.int32).view(INDEX_PAIR_DTYPE)
    cb = b.as_c()
    assert lib.f5c_dropin_selftest_scaling(ctypes.byref(cb), m.ctypes.data, k, 0, min_events, n_pairs.ctypes.data,
                                           res.ctypes.data, maps.ctypes.data, mp.ctypes.data) == 0
    usable = (b.good != 0) & (b.n_events >= 1) & (b.read_len >= k)
    est = np.zeros(b.n_reads, dtype=b.scalings.dtype)
    idx = np.flatnonzero(usable)
    est[idx] = ol.port_estimate_scalings(b.subset(idx), m)
    wb = ReadBatch(b.seq, b.seq_ptr, b.read_len, b.events, b.event_ptr, b.n_events, est, b.good, k)
    want_aln = ol.port_align(wb, m)
    want = ol.port_scaling(wb, m, want_aln, min_events=min_events)
    assert np.array_equal(n_pairs, want_aln.n_pairs)
    for f in ("flags", "n_event_alignment", "events_per_base"):
        assert np.array_equal(res[f], want.res[f]), f
    cal = want.res["calibrated"] != 0
    for f in ("shift", "scale"):
        assert np.array_equal(res["scalings"][f].view(np.uint32), want.res["scalings"][f].view(np.uint32)), f
    for f in ("var", "log_var"):
        assert np.array_equal(res["scalings"][f][cal].view(np.uint32), want.res["scalings"][f][cal].view(np.uint32)), f
    got = ol.ScalingResult(b, res, maps)
    for i in range(b.n_reads):
        if want.res["n_event_alignment"][i] > 0:
            assert np.array_equal(got.read_map(i), want.read_map(i)), i


@needs_so
def test_dropin_exports_getevents_entry_point():
    out = subprocess.check_output(["nm", "-D", "--defined-only", SO]).decode()
    assert "_Z14getevents_cudaP6core_tP4db_t" in out


@needs_so
@pytest.mark.gpu
def test_dropin_getevents_matches_oracle(built):
    """getevents_cuda on real core_t / db_t: event tables and the in-place pA conversion against the oracle."""
    from f5c_b200.abea import CSignals
    from f5c_b200.batch import EVENT_DTYPE
    lib = ctypes.CDLL(SO)
    vp = ctypes.c_void_p
    lib.f5c_dropin_selftest_events.argtypes = [ctypes.POINTER(CSignals), ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    sg = synth.make_signals(48, 1500, 0.6, seed=33)
    n = len(sg["n_samples"])
    cs = CSignals(n, sg["raw"].ctypes.data, sg["raw_ptr"].ctypes.data, sg["n_samples"].ctypes.data,
                  sg["offset"].ctypes.data, sg["range"].ctypes.data, sg["digitisation"].ctypes.data)
    cap = sg["n_samples"].astype(np.int64) // 2 + 2
    cap_ptr = np.zeros(n, dtype=np.int64)
    np.cumsum(cap[:-1], out=cap_ptr[1:])
    events = np.zeros(int(cap.sum()), dtype=EVENT_DTYPE)
    nev = np.zeros(n, dtype=np.int32)
    pa = np.zeros_like(sg["raw"])
    assert lib.f5c_dropin_selftest_events(ctypes.byref(cs), 0, 0, nev.ctypes.data, events.ctypes.data,
                                          cap_ptr.ctypes.data, pa.ctypes.data) == 0
    assert np.array_equal(pa, sg["pa"])
    for i in range(n):
        want = ol.port_getevents(sg["pa"][sg["raw_ptr"][i]:sg["raw_ptr"][i] + sg["n_samples"][i]])
        assert ol._events_equal(events[cap_ptr[i]:cap_ptr[i] + nev[i]], want), i


@needs_so
@pytest.mark.parametrize("threads", [1, 4])
def test_dropin_threaded_copies_roundtrip(threads):
    """The packer / unpacker copies of the drop-in (db_t's ragged arrays <-> flat staging) on N host threads; runs
    without a GPU (no CUDA call behind this door)."""
    lib = ctypes.CDLL(SO)
    vp = ctypes.c_void_p
    lib.f5c_dropin_selftest_pack.restype = ctypes.c_double
    lib.f5c_dropin_selftest_pack.argtypes = [ctypes.POINTER(CBatch), ctypes.c_int, vp, vp, vp, vp, vp, vp]
    b = synth.make_config("cfg2", seed=63, n_reads=300)
    k, m = models.load_model("r9")
    want = ol.port_align(b, ol.full_model(m))
    seq_out = np.full_like(b.seq, 0xEE)
    ev_out = np.zeros_like(b.events)
    pairs_rt = np.zeros_like(want.pairs)
    pp = b.pair_ptr()
    cb = b.as_c()
    ms = lib.f5c_dropin_selftest_pack(ctypes.byref(cb), threads, seq_out.ctypes.data, ev_out.ctypes.data,
                                      want.pairs.ctypes.data, pp.ctypes.data, want.n_pairs.ctypes.data,
                                      pairs_rt.ctypes.data)
    assert ms >= 0
    assert np.array_equal(seq_out, b.seq) and ev_out.tobytes() == b.events.tobytes()
    for i in range(b.n_reads):
        p, n = int(pp[i]), int(want.n_pairs[i])
        assert np.array_equal(pairs_rt[p:p + n], want.pairs[p:p + n])


@needs_so
@pytest.mark.gpu
def test_dropin_bench_door_ragged_db(built):
    """f5c_dropin_bench: align_cuda on a db_t whose reads are separate allocations (what load_db / event_single leave),
    several calls on the same core, through abea_align_ragged with 4 host threads — the e2e_dropin leg of bench.py."""
    lib = ctypes.CDLL(SO)
    vp = ctypes.c_void_p
    lib.f5c_dropin_bench.argtypes = [ctypes.POINTER(CBatch), vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, vp, vp, vp, vp]
    for b, name in ((synth.make_config("cfg3", seed=64, n_reads=300), "r10"), (edge_batch(), "r9")):
        k, m = models.load_model(name)
        m = ol.full_model(m)
        pairs = np.zeros(int(b.pair_capacity().sum()), dtype=PAIR_DTYPE)
        n_pairs = np.full(b.n_reads, -1, dtype=np.int32)
        pp = b.pair_ptr()
        ms = np.zeros(3, dtype=np.float64)
        cb = b.as_c()
        assert lib.f5c_dropin_bench(ctypes.byref(cb), m.ctypes.data, k, 0, 4, 1, 3, ms.ctypes.data, pairs.ctypes.data,
                                    pp.ctypes.data, n_pairs.ctypes.data) == 0
        assert (ms > 0).all()
        ol.assert_same_alignment(ol.AlignResult(b, pairs, n_pairs), ol.port_align(b, m), "drop-in bench door " + name)
