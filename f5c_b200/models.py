"""Built-in nucleotide pore-model tables (input data of the ABEA path).

The tables under f5c_b200/data/ were dumped from the reference's set_model() (reference src/model.c:132-193,
model ids src/f5cmisc.h:24-30) by tools/extract_models.py. level_log_stdv is recomputed on the host by the C
library (glibc ``logf`` — the reference is C++, so ``log(float)`` at src/model.c:179 is the float overload), never by numpy.
"""
from __future__ import annotations

import os

import numpy as np

from .batch import MODEL_DTYPE

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# name -> (reference model id, k, file)
MODELS = {
    "r9": (1, 6, "r9.4_450bps.nucleotide.6mer.npy"),
    "rna_r9": (3, 5, "r9.4_70bps.u_to_t_rna.5mer.npy"),
    "r10": (4, 9, "r10.4.1_400bps.nucleotide.9mer.npy"),
    "rna004": (6, 9, "rna004_130bps.u_to_t_rna.9mer.npy"),
}


def load_model(name: str):
    """Return (kmer_size, model) with model a MODEL_DTYPE array of 4^k entries; level_log_stdv left at 0
    until the C library fills it (abea_model_fill_log_stdv)."""
    _, k, fn = MODELS[name]
    tab = np.load(os.path.join(_DATA, fn))
    assert tab.shape == (4 ** k, 2)
    model = np.zeros(4 ** k, dtype=MODEL_DTYPE)
    model["level_mean"] = tab[:, 0]
    model["level_stdv"] = tab[:, 1]
    return k, model


def synthetic_model(k: int, seed: int = 7):
    """A seeded stand-in table with the same shape and value ranges as a real one (for tests that must not
    depend on the data files)."""
    rng = np.random.default_rng(seed)
    model = np.zeros(4 ** k, dtype=MODEL_DTYPE)
    model["level_mean"] = rng.normal(90.0, 12.0, 4 ** k).astype(np.float32)
    model["level_stdv"] = rng.uniform(1.2, 4.0, 4 ** k).astype(np.float32)
    return k, model
