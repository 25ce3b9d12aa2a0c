"""ctypes doors into the two CPU checkers (TEST INFRASTRUCTURE — never imported by f5c_b200/):

* ``port``  — oracle/libabea_oracle.so, our C restatement of ABEA (oracle/abea_oracle.c)
* ``ref``   — oracle/_ref/libf5c_ref.so, the unmodified reference objects behind oracle/ref_shim.cpp
              (exists only where it was built from /root/reference; tests that need it skip otherwise)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from f5c_b200.batch import (MODEL_DTYPE, PAIR_DTYPE, SCALINGS_DTYPE, INDEX_PAIR_DTYPE, SCALING_RESULT_DTYPE,
                             MIN_NUM_EVENTS_TO_RESCALE, ReadBatch, CBatch)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "libabea_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libf5c_ref.so")


class OracleStats(ctypes.Structure):
    _fields_ = [("sum_emission", ctypes.c_double), ("end_score", ctypes.c_float),
                ("n_aligned", ctypes.c_int32), ("max_gap", ctypes.c_int32), ("end_event", ctypes.c_int32),
                ("spanned", ctypes.c_int32), ("n_bands", ctypes.c_int64), ("n_fills", ctypes.c_int64)]


STATS_DTYPE = np.dtype([("sum_emission", "<f8"), ("end_score", "<f4"), ("n_aligned", "<i4"),
                        ("max_gap", "<i4"), ("end_event", "<i4"), ("spanned", "<i4"), ("n_bands", "<i8"),
                        ("n_fills", "<i8")], align=True)
assert STATS_DTYPE.itemsize == ctypes.sizeof(OracleStats)

_port = None
_ref = None
# (pairs, n_pairs, seq, seq_len, events, n_events, model, k, min_num_events_to_rescale, scalings io, map, result)
_SCALING_ARGS = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
                 ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]


def build_port():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build_port()
        lib = ctypes.CDLL(PORT_SO)
        lib.abea_oracle_align_batch.restype = ctypes.c_double
        lib.abea_oracle_align_batch.argtypes = [ctypes.POINTER(CBatch), ctypes.c_void_p, ctypes.c_uint32,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_int32]
        lib.abea_oracle_fill_log_stdv.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        lib.abea_oracle_estimate_scalings.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_void_p,
                                                      ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64,
                                                      ctypes.c_void_p]
        lib.abea_oracle_scaling_single.argtypes = _SCALING_ARGS
        lib.abea_oracle_getevents.restype = ctypes.c_int64
        lib.abea_oracle_getevents.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int8, ctypes.c_void_p]
        lib.abea_oracle_transitions.argtypes = [ctypes.c_int64, ctypes.c_int64] + [ctypes.c_void_p] * 4
        _port = lib
    return _port


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(REF_SO)
        lib.f5cref_set_model.restype = ctypes.c_uint32
        lib.f5cref_set_model.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        lib.f5cref_align_batch.restype = ctypes.c_double
        lib.f5cref_align_batch.argtypes = [ctypes.POINTER(CBatch), ctypes.c_void_p, ctypes.c_uint32,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
        lib.f5cref_estimate_scalings.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_void_p,
                                                 ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64,
                                                 ctypes.c_void_p]
        lib.f5cref_scaling_single.argtypes = _SCALING_ARGS
        lib.f5cref_getevents.restype = ctypes.c_int64
        lib.f5cref_getevents.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int8, ctypes.c_void_p,
                                         ctypes.c_int64]
        _ref = lib
    return _ref


def full_model(model: np.ndarray) -> np.ndarray:
    """Return a copy with level_log_stdv filled the way set_model does (glibc logf)."""
    m = np.ascontiguousarray(model.copy())
    assert m.dtype == MODEL_DTYPE
    port().abea_oracle_fill_log_stdv(m.ctypes.data, m.shape[0])
    return m


def ref_model(model_id: int):
    buf = np.zeros(262144, dtype=MODEL_DTYPE)
    k = ref().f5cref_set_model(buf.ctypes.data, model_id)
    return int(k), buf[:4 ** k].copy()


class AlignResult:
    """Per-read pair lists in the flat capacity layout (pair_ptr = prefix sum of E+L)."""

    def __init__(self, batch: ReadBatch, pairs, n_pairs, stats=None, seconds=0.0):
        self.pair_ptr = batch.pair_ptr()
        self.pairs = pairs
        self.n_pairs = n_pairs
        self.stats = stats
        self.seconds = seconds

    def read_pairs(self, i: int) -> np.ndarray:
        p = int(self.pair_ptr[i])
        return self.pairs[p:p + int(self.n_pairs[i])]


def _alloc(batch: ReadBatch):
    cap = batch.pair_capacity()
    pairs = np.zeros(int(cap.sum()), dtype=PAIR_DTYPE)
    n_pairs = np.full(batch.n_reads, -1, dtype=np.int32)
    return pairs, n_pairs


def port_align(batch: ReadBatch, model: np.ndarray, threads: int = 0) -> AlignResult:
    pairs, n_pairs = _alloc(batch)
    stats = np.zeros(batch.n_reads, dtype=STATS_DTYPE)
    pp = batch.pair_ptr()
    cb = batch.as_c()
    t = port().abea_oracle_align_batch(ctypes.byref(cb), model.ctypes.data, batch.kmer_size, pairs.ctypes.data,
                                       pp.ctypes.data, n_pairs.ctypes.data, stats.ctypes.data,
                                       threads or (os.cpu_count() or 1))
    return AlignResult(batch, pairs, n_pairs, stats, t)


def ref_align(batch: ReadBatch, model: np.ndarray, threads: int = 0) -> AlignResult:
    pairs, n_pairs = _alloc(batch)
    pp = batch.pair_ptr()
    cb = batch.as_c()
    t = ref().f5cref_align_batch(ctypes.byref(cb), model.ctypes.data, batch.kmer_size, pairs.ctypes.data,
                                 pp.ctypes.data, n_pairs.ctypes.data, threads or (os.cpu_count() or 1))
    return AlignResult(batch, pairs, n_pairs, None, t)


def assert_same_alignment(a: AlignResult, b: AlignResult, what: str = ""):
    """Bit-exact parity on the contract outputs: per-read counts and the pair lists themselves."""
    np.testing.assert_array_equal(a.n_pairs, b.n_pairs, err_msg=f"{what}: n_event_align_pairs differ")
    for i in range(a.n_pairs.shape[0]):
        pa, pb = a.read_pairs(i), b.read_pairs(i)
        if not np.array_equal(pa, pb):
            d = np.nonzero((pa["ref_pos"] != pb["ref_pos"]) | (pa["read_pos"] != pb["read_pos"]))[0]
            raise AssertionError(f"{what}: read {i} pairs differ first at {d[:5]}: {pa[d[:3]]} vs {pb[d[:3]]}")


class ScalingResult:
    """Per-read outputs of the stage after ABEA (scaling_single, reference src/f5c.c:736-807)."""

    def __init__(self, batch: ReadBatch, res, maps):
        self.map_ptr = batch.map_ptr()
        self.res = res      # SCALING_RESULT_DTYPE [n]
        self.maps = maps    # INDEX_PAIR_DTYPE [sum K]; meaningful only for reads with pairs

    def read_map(self, i: int) -> np.ndarray:
        return self.maps[int(self.map_ptr[i]):int(self.map_ptr[i + 1])]


def _estimate(fn, batch: ReadBatch, model: np.ndarray) -> np.ndarray:
    out = np.zeros(batch.n_reads, dtype=SCALINGS_DTYPE)
    for i in range(batch.n_reads):
        ev = np.ascontiguousarray(batch.read_events(i))
        seq = batch.read_seq(i)
        fn(seq, len(seq), model.ctypes.data, batch.kmer_size, ev.ctypes.data, len(ev), out[i:i + 1].ctypes.data)
    return out


def port_estimate_scalings(batch, model):
    """estimate_scalings_using_mom per read through the restatement."""
    return _estimate(port().abea_oracle_estimate_scalings, batch, model)


def ref_estimate_scalings(batch, model):
    """... and through the reference's own object code."""
    return _estimate(ref().f5cref_estimate_scalings, batch, model)


def _scaling(fn, batch: ReadBatch, model: np.ndarray, aln: AlignResult, scalings=None,
             min_events: int = MIN_NUM_EVENTS_TO_RESCALE) -> ScalingResult:
    res = np.zeros(batch.n_reads, dtype=SCALING_RESULT_DTYPE)
    mp = batch.map_ptr()
    maps = np.full(int(mp[-1]), -1, dtype=np.int32).repeat(2).view(INDEX_PAIR_DTYPE).copy()
    sc = np.ascontiguousarray((batch.scalings if scalings is None else scalings).copy())
    for i in range(batch.n_reads):
        ev = np.ascontiguousarray(batch.read_events(i))
        seq = batch.read_seq(i)
        pairs = np.ascontiguousarray(aln.read_pairs(i))
        m = maps[int(mp[i]):int(mp[i + 1])]
        fn(pairs.ctypes.data if len(pairs) else None, int(aln.n_pairs[i]), seq, len(seq), ev.ctypes.data, len(ev),
           model.ctypes.data, batch.kmer_size, min_events, sc[i:i + 1].ctypes.data,
           m.ctypes.data if len(m) else None, res[i:i + 1].ctypes.data)
    return ScalingResult(batch, res, maps)


def port_scaling(batch, model, aln, scalings=None, min_events=MIN_NUM_EVENTS_TO_RESCALE):
    return _scaling(port().abea_oracle_scaling_single, batch, model, aln, scalings, min_events)


def ref_scaling(batch, model, aln, scalings=None, min_events=MIN_NUM_EVENTS_TO_RESCALE):
    return _scaling(ref().f5cref_scaling_single, batch, model, aln, scalings, min_events)


def assert_same_scaling(a: ScalingResult, b: ScalingResult, what: str = "", check_var_d: bool = True):
    """Bit-exact parity of the scaling stage: the stored floats, the flags, the counts and the k-mer -> event map."""
    for f in ("n_event_alignment", "num_m_state", "flags", "calibrated", "events_per_base"):
        np.testing.assert_array_equal(a.res[f], b.res[f], err_msg=f"{what}: {f}")
    if check_var_d:
        np.testing.assert_array_equal(a.res["var_d"], b.res["var_d"], err_msg=f"{what}: var_d")
    for f in ("shift", "scale", "var", "log_var"):
        cal = a.res["calibrated"] != 0   # var / log_var are undefined (never written) for reads that were not recalibrated
        x, y = a.res["scalings"][f], b.res["scalings"][f]
        if f in ("var", "log_var"):
            x, y = x[cal], y[cal]
        np.testing.assert_array_equal(x.view(np.uint32), y.view(np.uint32), err_msg=f"{what}: scalings.{f}")
    for i in range(a.res.shape[0]):
        if a.res["n_event_alignment"][i] > 0:
            assert np.array_equal(a.read_map(i), b.read_map(i)), f"{what}: base_to_event_map of read {i}"


def _events_equal(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in ("start", "length", "mean", "stdv"))


def port_getevents(pa: np.ndarray, rna: bool = False) -> np.ndarray:
    """getevents (reference src/events.c:562-582) through the restatement; pa = float32 samples in pA."""
    from f5c_b200.batch import EVENT_DTYPE
    pa = np.ascontiguousarray(pa, dtype=np.float32)
    ev = np.zeros(len(pa) // 2 + 2, dtype=EVENT_DTYPE)
    n = port().abea_oracle_getevents(len(pa), pa.ctypes.data, int(rna), ev.ctypes.data)
    return ev[:n].copy()


def ref_getevents(pa: np.ndarray, rna: bool = False) -> np.ndarray:
    from f5c_b200.batch import EVENT_DTYPE
    pa = np.ascontiguousarray(pa.copy(), dtype=np.float32)
    ev = np.zeros(len(pa) // 2 + 2, dtype=EVENT_DTYPE)
    n = ref().f5cref_getevents(len(pa), pa.ctypes.data, int(rna), ev.ctypes.data, len(ev))
    assert n <= len(ev)
    return ev[:n].copy()
