#!/usr/bin/env python3
"""align_cuda(core_t*, db_t*) through the drop-in's bench door on one config (ABEA_TIME_PACK=1 prints where the time goes).
Usage: dropin_run.py [cfg] [threads] [steps]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
from f5c_b200 import models, synth
from f5c_b200.batch import CBatch, PAIR_DTYPE
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
b = synth.make_config(cfg, seed=42)
k, m = models.load_model(b.meta["model"])
m = ol.full_model(m)
lib = ctypes.CDLL(os.path.join(ROOT, "f5c_b200", "lib", "libf5c_abea_dropin.so"))
vp = ctypes.c_void_p
lib.f5c_dropin_bench.argtypes = [ctypes.POINTER(CBatch), vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
pairs = np.zeros(int(b.pair_capacity().sum()), dtype=PAIR_DTYPE)
n_pairs = np.zeros(b.n_reads, dtype=np.int32)
pp = b.pair_ptr()
ms = np.zeros(steps, dtype=np.float64)
cb = b.as_c()
rc = lib.f5c_dropin_bench(ctypes.byref(cb), m.ctypes.data, k, 0, threads, 2, steps, ms.ctypes.data, pairs.ctypes.data, pp.ctypes.data, n_pairs.ctypes.data)
print(cfg, "threads", threads, "rc", rc, "ms per align_cuda call:", np.round(ms, 2).tolist(), "aligned", int((n_pairs > 0).sum()))
