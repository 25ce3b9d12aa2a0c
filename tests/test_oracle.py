"""The oracle is pinned: our C restatement (oracle/abea_oracle.c) against the reference's golden vectors, the
committed fixtures generated from the reference (tests/golden/make_golden.py), and — where it was built — the
unmodified reference align() itself (oracle/_ref)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from f5c_b200 import models, synth
from f5c_b200.batch import ReadBatch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "abea_golden.json")))
NPZ = np.load(os.path.join(HERE, "golden", "abea_golden.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def r9_model():
    k, m = models.load_model("r9")
    return k, ol.full_model(m)


def single_read_batch():
    k, m = r9_model()
    b = ReadBatch.from_reads([NPZ["single_seq"].tobytes()], [NPZ["single_events"]], NPZ["single_scalings"], k)
    return b, m


def ecoli_batch():
    k, m = r9_model()
    b = ReadBatch(NPZ["ecoli_seq"], NPZ["ecoli_seq_ptr"], NPZ["ecoli_read_len"], NPZ["ecoli_events"],
                  NPZ["ecoli_event_ptr"], NPZ["ecoli_n_events"], NPZ["ecoli_scalings"],
                  np.ones(len(NPZ["ecoli_read_len"]), dtype=np.uint8), k)
    return b, m


def test_single_read_golden_vector():
    """reference test/ecoli_2kb_region/single_read/adaptive.exp: n_aligned_events 7206, sum_emission -20697.529925
    (inputs were printed with %f, hence the 2e-2 absolute slack on the sum)."""
    b, m = single_read_batch()
    r = ol.port_align(b, m, 1)
    g = GOLD["single_read"]
    assert r.n_pairs[0] == g["golden_n_aligned"] == 7206
    assert abs(r.stats["sum_emission"][0] - g["golden_sum_emission"]) < 2e-2
    assert sha(r.read_pairs(0)) == g["pairs_sha256"]  # pair list identical to the reference's


def test_ecoli_reads_golden_vectors():
    """reference test/ecoli_2kb_region/adaptive.exp lines for the committed subset of reads."""
    b, m = ecoli_batch()
    r = ol.port_align(b, m)
    g = GOLD["ecoli"]
    assert [int(x) for x in r.n_pairs] == g["n_pairs"]
    for i in range(b.n_reads):
        assert sha(r.read_pairs(i)) == g["pairs_sha256"][i]
        if g["golden_sum_emission"][i] is not None:
            assert int(r.stats["n_aligned"][i]) == g["golden_n_aligned"][i]
            assert abs(r.stats["sum_emission"][i] - g["golden_sum_emission"][i]) < 5e-2
    meta = GOLD["ecoli_all"]  # recorded when the fixtures were generated: all 111 distinct golden lines matched
    assert meta["reads_matched"] >= meta["golden_distinct"] == 111


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4"])
def test_synthetic_fixtures(cfg):
    """Seeded synthetic batches: the generator is deterministic and the oracle reproduces the reference's outputs."""
    g = GOLD["synthetic_" + cfg]
    b = synth.make_config(cfg, seed=g["seed"], n_reads=g["n_reads"])
    assert [int(x) for x in b.n_events] == g["n_events"]
    assert sha(b.events) == g["events_sha256"] and sha(b.seq) == g["seq_sha256"]
    k, m = models.load_model(b.meta["model"])
    r = ol.port_align(b, ol.full_model(m))
    assert [int(x) for x in r.n_pairs] == g["n_pairs"]
    assert [sha(r.read_pairs(i)) for i in range(b.n_reads)] == g["pairs_sha256"]


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name,kw", [
    ("r9", dict(n_reads=24, mean_events=2500, sigma=0.6, epk=1.8, seed=21)),
    ("r10", dict(n_reads=12, mean_events=3000, sigma=1.0, epk=1.9, seed=22)),
    ("rna004", dict(n_reads=6, mean_events=5000, sigma=0.5, epk=2.5, seed=23)),
    ("rna_r9", dict(n_reads=8, mean_events=1500, sigma=0.5, epk=2.2, seed=24)),
    ("r9", dict(n_reads=16, mean_events=90, sigma=0.9, epk=1.8, seed=25, min_len=12)),
])
def test_port_equals_reference(name, kw):
    b = synth.make_batch(name, **kw)
    mid = models.MODELS[name][0]
    k, mref = ol.ref_model(mid)
    k2, m = models.load_model(name)
    m = ol.full_model(m)
    assert k == k2 and m.tobytes() == mref.tobytes()  # table + logf(level_stdv) identical to set_model's
    ol.assert_same_alignment(ol.port_align(b, m), ol.ref_align(b, m), name)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_equals_reference_on_fixture_reads():
    b, m = ecoli_batch()
    ol.assert_same_alignment(ol.port_align(b, m), ol.ref_align(b, m), "ecoli")
    b, m = single_read_batch()
    ol.assert_same_alignment(ol.port_align(b, m), ol.ref_align(b, m), "single")


def test_transition_constants():
    import ctypes
    out = (ctypes.c_double * 4)()
    ol.port().abea_oracle_transitions(4000, 2223, *[ctypes.addressof(out) + 8 * i for i in range(4)])
    lp_skip, lp_stay, lp_step, lp_trim = list(out)
    assert lp_skip == np.log(1e-10) and lp_trim == np.log(0.01)
    assert abs(lp_stay - np.log(1 - 1 / (4000 / 2223 + 1))) < 1e-15
    assert abs(np.exp(lp_skip) + np.exp(lp_stay) + np.exp(lp_step) - 1.0) < 1e-12


def test_port_reproduces_reference_on_all_112_ecoli_reads():
    """BASELINE configs[0]: our restatement through every stage (getevents, estimate, align, scaling_single) on every read of
    tests/golden/ecoli/reads.blow5 against what the UNMODIFIED reference produced (tests/golden/ecoli_all.json, written by
    tests/golden/make_ecoli_all.py where the f5c tree is present). The GPU test of the same name checks the CUDA path."""
    import hashlib, json
    import blow5
    from f5c_b200.batch import SCALINGS_DTYPE

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "ecoli_all.json")))
    f = blow5.Blow5(os.path.join(here, "golden", "ecoli", "reads.blow5"))
    seqs = dict(blow5.read_fasta(os.path.join(here, "golden", "ecoli", "reads.fasta")))
    recs = {}
    for i in range(len(f)):
        rid, dig, off, rng, sr, sig = f.read(i)
        recs[rid] = (dig, off, rng, sig)
    k, m = models.load_model("r9")
    m = ol.full_model(m)
    evs, names = [], [r["name"] for r in gold["reads"]]
    for n in names:
        dig, off, rng, sig = recs[n]
        pa = ((sig.astype(np.float32) + np.float32(off)) * np.float32(np.float32(rng) / np.float32(dig))).astype(np.float32)
        evs.append(ol.port_getevents(pa))
    assert [len(e) for e in evs] == [r["n_events"] for r in gold["reads"]]
    b = ReadBatch.from_reads([seqs[n].encode() for n in names], evs, np.zeros(len(names), dtype=SCALINGS_DTYPE), k)
    b.scalings[:] = ol.port_estimate_scalings(b, m)
    assert [int(x) for x in b.scalings["shift"].view(np.uint32)] == [r["shift_bits"] for r in gold["reads"]]
    a = ol.port_align(b, m)
    assert [int(x) for x in a.n_pairs] == [r["n_pairs"] for r in gold["reads"]]
    assert [sha(a.read_pairs(i)) for i in range(b.n_reads)] == [r["pairs_sha256"] for r in gold["reads"]]
