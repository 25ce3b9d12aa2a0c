#!/usr/bin/env python3
"""BLOW5 + FASTA in, `f5c resquiggle` output out (reference src/resquiggle.c: TSV "read_id kmer_idx start_raw_idx
end_raw_idx", or PAF with -c), everything on the GPU — the BAM-free end-to-end driver of SURVEY.md 8(f) N2:

    python tools/blow5_resquiggle.py reads.blow5 reads.fasta [out.txt] [--model r9|r10|rna004|rna_r9] [-c]

records -> (device: inflate, parse, signal decode) -> events -> method-of-moments scalings -> ABEA -> postalign +
recalibration -> per k-mer the raw-signal range of its events. Host work: the file's framing, the FASTA and the text."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import blow5
from f5c_b200 import models
from f5c_b200.abea import AbeaContext, scaling_db, write_resquiggle
from f5c_b200.batch import EVENT_DTYPE, SCALINGS_DTYPE, ReadBatch


def main():
    paf = "-c" in sys.argv
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    model = "r9"
    if "--model" in sys.argv:
        model = sys.argv[sys.argv.index("--model") + 1]
        args = [a for a in args if a != model]
    rna = model.startswith("rna")
    f = blow5.Blow5(args[0])
    seqs = dict(blow5.read_fasta(args[1]))
    out = args[2] if len(args) > 2 else "-"
    names, idx = [], []
    for i in range(len(f)):
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
        nev, ns, _ = ctx.getevents_blow5(payload, rec_ptr, rec_len, f.record_method, f.signal_method, rna=rna)
        ev, ev_ptr = ctx.events_download(nev)          # signal order: start / length are what the output needs
        seq_b = [seqs[n].encode() for n in names]
        shell = ReadBatch.from_reads(seq_b, [np.zeros(0, dtype=EVENT_DTYPE)] * len(names), np.zeros(len(names), dtype=SCALINGS_DTYPE), k)
        shell.n_events = np.maximum(nev, 0).astype(np.int32)
        ctx.upload(shell, with_scalings=False, device_events=True)
        ctx.estimate_scalings(len(names), reverse_events=rna)
        ctx.run()
        sc = scaling_db(ctx, shell)
    if rna:  # the alignment ran on the reversed tables (src/f5c.c:713-721): the map's event indices refer to those
        ev = np.concatenate([ev[int(ev_ptr[i]):int(ev_ptr[i]) + int(nev[i])][::-1] for i in range(len(names))]) if len(names) else ev
    write_resquiggle(out, names, shell.read_len, ns, ev, ev_ptr, sc.results, sc.maps, sc.map_ptr, k,
                     fmt="paf" if paf else "tsv", rna=rna)


if __name__ == "__main__":
    main()
