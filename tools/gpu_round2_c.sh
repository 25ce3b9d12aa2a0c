#!/bin/bash
# round 2, run c: K-walker traceback, BLOW5 decode tests, per-read cycles
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r3_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
for c in cfg5 cfg3 cfg4; do timeout 300 python tools/prof_run.py $c - 3 > gpurun_out/r3_prof_$c.txt 2>&1; done
for m in 32 128; do ABEA_TB_MARGIN=$m timeout 300 python tools/prof_run.py cfg5 - 3 > gpurun_out/r3_prof_cfg5_m$m.txt 2>&1; done
timeout 300 python - > gpurun_out/r3_blow5.txt 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, blow5
from f5c_b200.abea import AbeaContext
f = blow5.Blow5("tests/golden/ecoli/reads.blow5")
idx = list(range(len(f))) * 16          # 1792 records, 108 MB compressed
chunks = [f.record_bytes(i) for i in idx]
rec_len = np.array([len(c) for c in chunks], dtype=np.int32)
rec_ptr = np.zeros(len(chunks), dtype=np.int64); np.cumsum(rec_len[:-1].astype(np.int64), out=rec_ptr[1:])
payload = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy()
with AbeaContext(0) as ctx:
    pp = ctx.pin_array(payload)
    for rep in range(3):
        t0 = time.perf_counter(); nev, ns, t = ctx.getevents_blow5(pp, rec_ptr, rec_len, f.record_method, f.signal_method); dt = time.perf_counter() - t0
        print("records", len(idx), "compressed MB %.1f" % (payload.nbytes / 1e6), "samples M %.1f" % (ns.sum() / 1e6), "events M %.2f" % (nev.sum() / 1e6),
              "wall ms %.1f" % (dt * 1e3), "blow5_ms %.2f" % t["blow5_ms"], "events_ms %.2f" % t["events_ms"], "h2d_ms %.2f" % t["h2d_ms"])
import zlib
t0 = time.perf_counter()
for c in chunks[:112]: zlib.decompress(c)
print("host zlib, 1 thread, 112 records: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
PY
tail -4 gpurun_out/r3_tests.log; cut -c1-330 gpurun_out/r3_bench.json; tail -8 gpurun_out/r3_prof_cfg5.txt; tail -7 gpurun_out/r3_prof_cfg3.txt; cat gpurun_out/r3_blow5.txt
