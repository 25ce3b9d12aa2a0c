#!/usr/bin/env python3
"""Golden outputs of the UNMODIFIED reference for ALL reads of test/ecoli_2kb_region (BASELINE configs[0]).

Runs only where /root/reference exists. Copies of reads.blow5 and reads.fasta are committed under tests/golden/ecoli/
(data fixtures); for every read this script runs the reference's own getevents -> estimate_scalings_using_mom ->
align -> scaling_single (oracle/_ref) on the signal decoded from the BLOW5 file and records
    n_events, sha256 of the event table fields, shift / scale bit patterns, n_pairs, sha256 of the pair list,
    the recalibrated scalings' bit patterns and flags
in tests/golden/ecoli_all.json. The GPU tests re-derive every stage from the BLOW5 bytes and compare (112 reads).
"""
import hashlib, json, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
import blow5
from f5c_b200.batch import EVENT_DTYPE, SCALINGS_DTYPE, ReadBatch


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def to_pa(sig, dig, off, rng):
    """event_single, src/f5c.c:692-696: all float"""
    rawf = sig.astype(np.float32)
    raw_unit = np.float32(np.float32(rng) / np.float32(dig))
    return np.ascontiguousarray(((rawf + np.float32(off)) * raw_unit).astype(np.float32))


def event_sha(ev):
    return sha(np.concatenate([ev["start"].astype("<u8").view(np.uint8), ev["length"].astype("<f4").view(np.uint8),
                               ev["mean"].astype("<f4").view(np.uint8), ev["stdv"].astype("<f4").view(np.uint8)]))


def main():
    k, model = ol.ref_model(1)
    f = blow5.Blow5(os.path.join(HERE, "ecoli", "reads.blow5"))
    seqs = dict(blow5.read_fasta(os.path.join(HERE, "ecoli", "reads.fasta")))
    names, seq_l, ev_l, sc_l, nsamp = [], [], [], [], []
    for i in range(len(f)):
        rid, dig, off, rng, sr, sig = f.read(i)
        if rid not in seqs:
            continue
        pa = to_pa(sig, dig, off, rng)
        ev = ol.ref_getevents(pa)
        seq = seqs[rid].encode()
        sc = np.zeros(1, dtype=SCALINGS_DTYPE)
        ol.ref().f5cref_estimate_scalings(seq, len(seq), model.ctypes.data, k, ev.ctypes.data, len(ev), sc.ctypes.data)
        names.append(rid); seq_l.append(seq); ev_l.append(ev); sc_l.append(sc[0]); nsamp.append(len(sig))
    b = ReadBatch.from_reads(seq_l, ev_l, np.array(sc_l), k)
    aln = ol.ref_align(b, model)
    scl = ol.ref_scaling(b, model, aln)
    out = {"model": "r9", "kmer_size": k, "n_reads": b.n_reads, "n_events_total": int(b.n_events.sum()), "reads": []}
    for i in range(b.n_reads):
        r = scl.res[i]
        out["reads"].append({
            "name": names[i], "n_samples": int(nsamp[i]), "n_events": int(b.n_events[i]), "events_sha256": event_sha(ev_l[i]),
            "shift_bits": int(np.float32(sc_l[i]["shift"]).view(np.uint32)), "scale_bits": int(np.float32(sc_l[i]["scale"]).view(np.uint32)),
            "n_pairs": int(aln.n_pairs[i]), "pairs_sha256": sha(aln.read_pairs(i)),
            "recal_shift_bits": int(r["scalings"]["shift"].view(np.uint32)), "recal_scale_bits": int(r["scalings"]["scale"].view(np.uint32)),
            "flags": int(r["flags"]), "n_event_alignment": int(r["n_event_alignment"]),
            "map_sha256": sha(scl.read_map(i)) if r["n_event_alignment"] > 0 else None})
    json.dump(out, open(os.path.join(HERE, "ecoli_all.json"), "w"), indent=0)
    print("reads", b.n_reads, "events", int(b.n_events.sum()), "aligned", int((aln.n_pairs > 0).sum()))


if __name__ == "__main__":
    main()
