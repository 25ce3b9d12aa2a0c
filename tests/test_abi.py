"""The C-ABI shared library loads on a box without a GPU, exports every symbol include/abea_b200.h declares, and the
product refuses to run (loudly) when there is no CUDA device — there is no CPU fallback to fall into."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "abea_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(abea_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built):
    from f5c_b200 import abea
    lib = abea.load_library()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.abea_version()


def test_types_match_reference_sizes():
    from f5c_b200 import batch
    assert batch.EVENT_DTYPE.itemsize == 24 and batch.MODEL_DTYPE.itemsize == 12
    assert batch.SCALINGS_DTYPE.itemsize == 16 and batch.PAIR_DTYPE.itemsize == 8
    assert ctypes.sizeof(batch.CBatch) == 80   # abea_batch_t: 8 pointers + n_reads + event_means


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    from f5c_b200.abea import AbeaContext, AbeaError
    with pytest.raises(AbeaError):
        AbeaContext(0)


def test_product_never_touches_the_oracle():
    """f5c_b200/ must not import, link or dlopen anything under oracle/ or tests/."""
    pkg = os.path.join(ROOT, "f5c_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle_lib" not in txt and "libabea_oracle" not in txt and "libf5c_ref" not in txt, fn
                assert "libabea_emu" not in txt, fn


def test_eligibility_filter_matches_reference_float_rule():
    from f5c_b200 import synth
    b = synth.make_batch("r9", n_reads=4, mean_events=300, sigma=0.2, epk=1.8, seed=1)
    b.good[2] = 0
    el = b.eligible()
    assert el.tolist() == [True, True, False, True]
    assert b.events_aligned() == int(b.n_events[[0, 1, 3]].sum())
    assert np.all(b.n_bands == b.n_events.astype(np.int64) + b.read_len - b.kmer_size + 1 + 2)
