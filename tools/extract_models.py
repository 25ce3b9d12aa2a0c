#!/usr/bin/env python3
"""Dump the reference's built-in nucleotide pore-model tables to f5c_b200/data/*.npy.

The tables are input DATA of the ABEA path (k-mer -> level_mean, level_stdv), compiled into the reference
binary from src/model.h and reached only through set_model() (src/model.c:132-193). This script calls the
unmodified reference set_model() through oracle/_ref/libf5c_ref.so (so it runs only where /root/reference
was available to build that library) and stores float32 [4^k, 2] arrays. level_log_stdv is NOT stored: the
product recomputes it on the host with libm exactly as set_model does (src/model.c:179).
"""
import ctypes, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libf5c_ref.so"))
lib.f5cref_set_model.restype = ctypes.c_uint32
lib.f5cref_set_model.argtypes = [ctypes.c_void_p, ctypes.c_uint32]

NAMES = {1: "r9.4_450bps.nucleotide.6mer", 3: "r9.4_70bps.u_to_t_rna.5mer",
         4: "r10.4.1_400bps.nucleotide.9mer", 6: "rna004_130bps.u_to_t_rna.9mer"}
out_dir = os.path.join(ROOT, "f5c_b200", "data")
os.makedirs(out_dir, exist_ok=True)
for mid, name in NAMES.items():
    buf = np.zeros((262144, 3), dtype=np.float32)
    k = lib.f5cref_set_model(buf.ctypes.data, mid)
    n = 4 ** k
    tab = np.ascontiguousarray(buf[:n, :2])
    np.save(os.path.join(out_dir, name + ".npy"), tab)
    print(mid, name, "k=%d" % k, tab.shape, tab[:2].tolist(), "stdv range", tab[:, 1].min(), tab[:, 1].max())
