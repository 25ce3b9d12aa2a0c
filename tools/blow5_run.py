#!/usr/bin/env python3
"""Timing of the device-side BLOW5 decode: the ecoli fixture's 112 zlib records replicated 16 times
(1792 records, 108 MB compressed -> 88 M samples) through abea_getevents_blow5, beside host zlib on one thread."""
import sys, time, os, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, blow5
from f5c_b200.abea import AbeaContext
name = sys.argv[1] if len(sys.argv) > 1 else "reads.blow5"
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 16
f = blow5.Blow5(os.path.join(ROOT, "tests", "golden", "ecoli", name))
idx = list(range(len(f))) * rep
chunks = [f.record_bytes(i) for i in idx]
rec_len = np.array([len(c) for c in chunks], dtype=np.int32)
rec_ptr = np.zeros(len(chunks), dtype=np.int64); np.cumsum(rec_len[:-1].astype(np.int64), out=rec_ptr[1:])
payload = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy()
with AbeaContext(0) as ctx:
    pp = ctx.pin_array(payload)
    for r in range(3):
        t0 = time.perf_counter(); nev, ns, t = ctx.getevents_blow5(pp, rec_ptr, rec_len, f.record_method, f.signal_method); dt = time.perf_counter() - t0
        print(name, "records", len(idx), "compressed MB %.1f" % (payload.nbytes / 1e6), "samples M %.1f" % (ns.sum() / 1e6), "events M %.2f" % (nev.sum() / 1e6),
              "wall ms %.1f" % (dt * 1e3), "decode (inflate+parse+signal) ms %.2f" % t["blow5_ms"], "events_ms %.2f" % t["events_ms"], "h2d_ms %.2f" % t["h2d_ms"], flush=True)
if f.record_method == 1:
    t0 = time.perf_counter()
    for c in chunks[:len(f)]: zlib.decompress(c)
    print("host zlib, 1 thread, %d records: %.1f ms" % (len(f), (time.perf_counter() - t0) * 1e3))
