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
