#!/usr/bin/env python3
"""BLOW5 + FASTA in, the reference's --print-banded-aln text out (src/f5c.c:989-1006), everything on the GPU:

    python tools/blow5_eventalign_dump.py reads.blow5 reads.fasta [out.txt] [--model r9|r10|rna004|rna_r9]

records -> (device: inflate, parse, signal decode) -> events -> method-of-moments scalings -> ABEA -> recalibration ->
per read ">name\\tN_ALGN_PAIR:n\\t{ref_pos,read_pos}" and its "{k,e}" pairs, skipping reads that failed the alignment,
for diffing against `f5c eventalign --print-banded-aln` run on the same reads. The only host work is walking the
file's framing (tests/blow5.py, a 60-line reader) and the FASTA."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import blow5
from f5c_b200 import models
from f5c_b200.abea import AbeaContext, scaling_db, write_pairs
from f5c_b200.batch import EVENT_DTYPE, SCALINGS_DTYPE, ReadBatch


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    model = "r9"
    if "--model" in sys.argv:
        model = sys.argv[sys.argv.index("--model") + 1]
        args = [a for a in args if a != model]
    f = blow5.Blow5(args[0])
    seqs = dict(blow5.read_fasta(args[1]))
    out = args[2] if len(args) > 2 else "-"
    names, idx = [], []
    for i in range(len(f)):   # read ids sit in the (compressed) records: the host decode of the id alone is the join key
        rid = f.read(i)[0]
        if rid in seqs:
            names.append(rid); idx.append(i)
    chunks = [f.record_bytes(i) for i in idx]
    rec_len = np.array([len(c) for c in chunks], dtype=np.int32)
    rec_ptr = np.zeros(len(chunks), dtype=np.int64)
    np.cumsum(rec_len[:-1].astype(np.int64), out=rec_ptr[1:])
    payload = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy()
    k, m = models.load_model(model)
    with AbeaContext(0) as ctx:
        ctx.set_model(m, k)
        nev, ns, _ = ctx.getevents_blow5(payload, rec_ptr, rec_len, f.record_method, f.signal_method, rna=model.startswith("rna"))
        seq_b = [seqs[n].encode() for n in names]
        shell = ReadBatch.from_reads(seq_b, [np.zeros(0, dtype=EVENT_DTYPE)] * len(names), np.zeros(len(names), dtype=SCALINGS_DTYPE), k)
        shell.n_events = np.maximum(nev, 0).astype(np.int32)
        ctx.upload(shell, with_scalings=False, device_events=True)
        ctx.estimate_scalings(len(names), reverse_events=model.startswith("rna"))
        ctx.run()
        aln = ctx.download(shell)
        sc = scaling_db(ctx, shell)
    write_pairs(out, names, aln, sc.results["flags"])


if __name__ == "__main__":
    main()
