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

from f5c_b200.batch import MODEL_DTYPE, PAIR_DTYPE, ReadBatch, CBatch

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
