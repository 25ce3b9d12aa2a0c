#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ from the reference itself.

Runs only where /root/reference exists (this container). What it does:

1. single_read: parses the reference's own debug dumps test/ecoli_2kb_region/single_read/{read1.events.exp,
   read1.fasta, read1.scalings.exp, adaptive.exp}, runs the UNMODIFIED reference estimate_scalings_using_mom +
   align() (oracle/_ref) on them and checks the golden line (n_aligned_events 7206 exact; sum_emission to the
   precision the %f-printed inputs allow).
2. ecoli_2kb_region: builds the vendored slow5lib from a scratch copy, dumps reads.blow5, converts to pA exactly
   as event_single does (src/f5c.c:692-696), runs reference getevents -> MoM -> align for all reads and checks
   every result against test/ecoli_2kb_region/adaptive.exp (n_aligned_events exact, sum_emission |d|<0.05).
3. Writes tests/golden/abea_golden.npz: the single read plus a spread of ecoli reads (events, sequences, scalings)
   with the reference's outputs (n_pairs, sha256 of the pair bytes, the adaptive.exp values), and seeded
   synthetic batches' reference outputs (n_pairs + sha256 per read) for R9 / R10 / RNA004.

The reference never travels to the GPU box; these fixtures (and oracle/_ref/*.so) do.
"""
import hashlib, json, os, re, shutil, struct, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from f5c_b200.batch import EVENT_DTYPE, SCALINGS_DTYPE, ReadBatch
from f5c_b200 import synth, models

REF = "/root/reference"
ECOLI = os.path.join(REF, "test", "ecoli_2kb_region")


def sha(pairs):
    return hashlib.sha256(np.ascontiguousarray(pairs).tobytes()).hexdigest()


def parse_adaptive(path):
    rows = []
    for line in open(path):
        m = re.match(r"sum_emission (\S+), n_aligned_events (\S+), avg_log_emission (\S+)", line)
        if m:
            rows.append((float(m.group(1)), int(float(m.group(2))), float(m.group(3))))
    return rows


def read_fasta(path):
    out, name, buf = [], None, []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if name is not None:
                out.append((name, "".join(buf)))
            name, buf = line[1:].split()[0], []
        elif line:
            buf.append(line)
    if name is not None:
        out.append((name, "".join(buf)))
    return out


def ref_scalings(seq, events, model, k):
    sc = np.zeros(1, dtype=SCALINGS_DTYPE)
    ol.ref().f5cref_estimate_scalings(seq, len(seq), model.ctypes.data, k, events.ctypes.data, len(events),
                                      sc.ctypes.data)
    return sc[0]


def single_read(model, k):
    d = os.path.join(ECOLI, "single_read")
    txt = open(os.path.join(d, "read1.events.exp")).read()
    tup = re.findall(r"\{(\d+),([-\d.]+),([-\d.]+),([-\d.]+),-1,-1\}", txt)
    ev = np.zeros(len(tup), dtype=EVENT_DTYPE)
    ev["start"] = [int(t[0]) for t in tup]
    ev["length"] = [float(t[1]) for t in tup]
    ev["mean"] = [float(t[2]) for t in tup]
    ev["stdv"] = [float(t[3]) for t in tup]
    (name, seq), = read_fasta(os.path.join(d, "read1.fasta"))
    seq = seq.encode()
    sc = ref_scalings(seq, ev, model, k)
    gold = parse_adaptive(os.path.join(d, "adaptive.exp"))[0]
    print("single_read: events", len(ev), "bases", len(seq), "shift %.2f scale %.2f" % (sc["shift"], sc["scale"]),
          "golden", gold)
    assert "%.2f" % sc["shift"] == "1.95" and "%.2f" % sc["scale"] == "1.00"  # read1.scalings.exp
    return name, seq, ev, sc, gold


def ecoli_raw():
    """{read_id: (digitisation, offset, range, sample_rate, int16 signal)} of test/ecoli_2kb_region/reads.blow5."""
    tmp = tempfile.mkdtemp(prefix="slow5build")
    src = os.path.join(tmp, "slow5lib")
    shutil.copytree(os.path.join(REF, "slow5lib"), src)
    subprocess.check_call(["make", "-s", "-C", src, "lib/libslow5.a"], stdout=subprocess.DEVNULL)
    exe = os.path.join(tmp, "blow5_dump")
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(src, "include"), os.path.join(HERE, "blow5_dump.c"),
                           os.path.join(src, "lib", "libslow5.a"), "-lz", "-lm", "-lpthread", "-o", exe])
    dump = os.path.join(tmp, "reads.bin")
    subprocess.check_call([exe, os.path.join(ECOLI, "reads.blow5"), dump])
    raw = {}
    with open(dump, "rb") as f:
        while True:
            h = f.read(4)
            if not h:
                break
            (l,) = struct.unpack("<I", h)
            rid = f.read(l).decode()
            dig, off, rng, sr, n = struct.unpack("<ddddQ", f.read(40))
            sig = np.frombuffer(f.read(2 * n), dtype=np.int16)
            raw[rid] = (dig, off, rng, sr, sig)
    shutil.rmtree(tmp)
    return raw


def to_pa(sig, dig, off, rng):
    """event_single, src/f5c.c:692-696: all float"""
    rawf = sig.astype(np.float32)
    raw_unit = np.float32(np.float32(rng) / np.float32(dig))
    return np.ascontiguousarray(((rawf + np.float32(off)) * raw_unit).astype(np.float32))


def ecoli_reads(model, k):
    raw = ecoli_raw()
    reads = []
    for name, seq in read_fasta(os.path.join(ECOLI, "reads.fasta")):
        if name not in raw:
            continue
        dig, off, rng, sr, sig = raw[name]
        pa = to_pa(sig, dig, off, rng)
        ev = np.zeros(len(pa), dtype=EVENT_DTYPE)
        n = ol.ref().f5cref_getevents(len(pa), pa.ctypes.data, 0, ev.ctypes.data, len(ev))
        ev = ev[:n].copy()
        seqb = seq.encode()
        sc = ref_scalings(seqb, ev, model, k)
        reads.append((name, seqb, ev, sc))
    return reads


def main():
    k, model = ol.ref_model(1)
    out = {}
    meta = {}

    name, seq, ev, sc, gold = single_read(model, k)
    b1 = ReadBatch.from_reads([seq], [ev], np.array([sc]), k)
    r1 = ol.ref_align(b1, model, 1)
    p1 = ol.port_align(b1, model, 1)
    ol.assert_same_alignment(p1, r1, "single_read")
    print("  reference n_pairs", r1.n_pairs[0], "port sum_emission %.6f" % p1.stats["sum_emission"][0])
    assert r1.n_pairs[0] == gold[1] == 7206
    assert abs(p1.stats["sum_emission"][0] - gold[0]) < 0.02
    meta["single_read"] = dict(name=name, golden_sum_emission=gold[0], golden_n_aligned=gold[1],
                               n_pairs=int(r1.n_pairs[0]), pairs_sha256=sha(r1.read_pairs(0)))
    out["single_seq"] = np.frombuffer(seq, dtype=np.uint8)
    out["single_events"] = ev
    out["single_scalings"] = np.array([sc])

    reads = ecoli_reads(model, k)
    be = ReadBatch.from_reads([r[1] for r in reads], [r[2] for r in reads], np.array([r[3] for r in reads]), k)
    re_ = ol.ref_align(be, model)
    pe = ol.port_align(be, model)
    ol.assert_same_alignment(pe, re_, "ecoli")
    gold_rows = parse_adaptive(os.path.join(ECOLI, "adaptive.exp"))
    gold_by_n = {}
    for s, n, a in gold_rows:
        gold_by_n.setdefault(n, []).append(s)
    matched = 0
    per_read_gold = []
    for i in range(be.n_reads):
        n = int(pe.stats["n_aligned"][i])
        s = float(pe.stats["sum_emission"][i])
        cands = gold_by_n.get(n, [])
        ok = any(abs(s - g) < 0.05 for g in cands)
        per_read_gold.append(min(cands, key=lambda g: abs(g - s)) if ok else None)
        matched += ok
    distinct = len({(n, round(s, 3)) for s, n, a in gold_rows})
    print("ecoli: reads", be.n_reads, "events", int(be.n_events.sum()), "golden lines", len(gold_rows),
          "distinct", distinct, "reads matched to a golden line:", matched)
    assert matched >= distinct, (matched, distinct)
    meta["ecoli_all"] = dict(n_reads=be.n_reads, n_events=int(be.n_events.sum()), golden_lines=len(gold_rows),
                             golden_distinct=distinct, reads_matched=int(matched),
                             ref_vs_port_pairs_identical=True)

    # keep a spread of reads (shortest, longest, quartiles, a QC failure if any) small enough to commit
    order = np.argsort(be.n_events)
    pick = sorted({int(order[j]) for j in np.linspace(0, len(order) - 1, 8).astype(int)})
    fails = [i for i in range(be.n_reads) if re_.n_pairs[i] == 0]
    pick = sorted(set(pick) | set(fails[:2]))
    sub = be.subset(pick)
    rs = ol.ref_align(sub, model, 1)
    out["ecoli_seq"], out["ecoli_seq_ptr"], out["ecoli_read_len"] = sub.seq, sub.seq_ptr, sub.read_len
    out["ecoli_events"], out["ecoli_event_ptr"], out["ecoli_n_events"] = sub.events, sub.event_ptr, sub.n_events
    out["ecoli_scalings"] = sub.scalings
    meta["ecoli"] = dict(names=[reads[i][0] for i in pick], n_pairs=[int(x) for x in rs.n_pairs],
                         pairs_sha256=[sha(rs.read_pairs(j)) for j in range(sub.n_reads)],
                         golden_sum_emission=[per_read_gold[i] for i in pick],
                         golden_n_aligned=[int(pe.stats["n_aligned"][i]) if per_read_gold[i] is not None else None
                                           for i in pick])
    print("  kept reads:", pick, "events", int(sub.n_events.sum()))

    # synthetic batches: reference outputs per read
    for cfg, n, seed in (("cfg2", 24, 101), ("cfg3", 12, 102), ("cfg4", 4, 103)):
        b = synth.make_config(cfg, seed=seed, n_reads=n)
        kk, m = ol.ref_model(models.MODELS[b.meta["model"]][0])
        r = ol.ref_align(b, m)
        p = ol.port_align(b, m)
        ol.assert_same_alignment(p, r, cfg)
        meta["synthetic_" + cfg] = dict(n_reads=n, seed=seed, n_events=[int(x) for x in b.n_events],
                                        events_sha256=sha(b.events), seq_sha256=sha(b.seq),
                                        n_pairs=[int(x) for x in r.n_pairs],
                                        pairs_sha256=[sha(r.read_pairs(j)) for j in range(n)])
        print(cfg, "reads", n, "events", int(b.n_events.sum()), "pairs", int(r.n_pairs.sum()))

    np.savez_compressed(os.path.join(HERE, "abea_golden.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "abea_golden.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "abea_golden.npz"), os.path.getsize(os.path.join(HERE, "abea_golden.npz")))


if __name__ == "__main__":
    main()
