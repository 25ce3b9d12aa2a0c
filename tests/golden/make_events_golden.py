#!/usr/bin/env python3
"""Golden fixture for event detection, generated from the reference itself.

Runs only where /root/reference exists. Dumps test/ecoli_2kb_region/reads.blow5 (vendored slow5lib, as
make_golden.py does), converts to pA exactly as event_single does (src/f5c.c:692-696) and runs the UNMODIFIED
reference getevents (oracle/_ref) on all 112 reads; checks that (1) our restatement (oracle/abea_oracle.c
abea_oracle_getevents) gives the identical event tables, and (2) getevents on the raw signal of the
reference's single_read reproduces its golden event table single_read/read1.events.exp (7165 events: starts and
lengths exact, mean / stdv to the %f print precision), and (3) for the reads already committed in abea_golden.npz the
tables are the very events that fixture holds — the ones that reproduce the reference's adaptive.exp lines. Writes tests/golden/events_golden.npz with the raw int16
signals + calibration of three of those reads, so that the check travels to the GPU box without the reference.
"""
import json, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
import oracle_lib as ol
import make_golden as mg


def main():
    raw = mg.ecoli_raw()
    G = json.load(open(os.path.join(HERE, "abea_golden.json")))
    NPZ = np.load(os.path.join(HERE, "abea_golden.npz"))
    n_samples = 0
    for rid, (dig, off, rng, sr, sig) in raw.items():
        pa = mg.to_pa(sig, dig, off, rng)
        a, b = ol.ref_getevents(pa), ol.port_getevents(pa)
        assert ol._events_equal(a, b), rid
        n_samples += len(pa)
    print("restatement == reference getevents on", len(raw), "reads,", n_samples, "samples")
    # the reference's own golden event table: single_read/read1.events.exp (printed with %f) against getevents on
    # that read's raw signal from reads.blow5
    rid = G["single_read"]["name"]
    dig1, off1, rng1, sr1, sig1 = raw[rid]
    ev1 = ol.ref_getevents(mg.to_pa(sig1, dig1, off1, rng1))
    g1 = NPZ["single_events"]
    assert len(ev1) == len(g1) == 7165
    assert np.array_equal(ev1["start"], g1["start"]) and np.array_equal(ev1["length"], g1["length"])
    assert np.abs(ev1["mean"].astype(np.float64) - g1["mean"]).max() < 1e-6
    assert np.abs(ev1["stdv"].astype(np.float64) - g1["stdv"]).max() < 1e-6
    print("single_read/read1.events.exp reproduced: 7165 events, starts and lengths exact, mean / stdv to print precision")
    names = G["ecoli"]["names"]
    ne = NPZ["ecoli_n_events"]
    order = np.argsort(ne)[:3]
    out = {}
    for j, i in enumerate(order):
        i = int(i)
        dig, off, rng, sr, sig = raw[names[i]]
        ev = ol.ref_getevents(mg.to_pa(sig, dig, off, rng))
        p = int(NPZ["ecoli_event_ptr"][i])
        fix = NPZ["ecoli_events"][p:p + int(ne[i])]
        assert ol._events_equal(ev, fix), names[i]
        out[f"sig{j}"] = sig.astype(np.int16)
        out[f"cal{j}"] = np.array([off, rng, dig], dtype=np.float64)
        out[f"idx{j}"] = np.array([i], dtype=np.int64)
    out["single_sig"] = sig1.astype(np.int16)
    out["single_cal"] = np.array([off1, rng1, dig1], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "events_golden.npz"), **out)
    meta = dict(reads_checked=len(raw), samples=n_samples, single_read_events_exp_reproduced=True,
                committed=[names[int(i)] for i in order])
    json.dump(meta, open(os.path.join(HERE, "events_golden.json"), "w"), indent=1)
    print("wrote events_golden.npz", [(names[int(i)], int(ne[int(i)])) for i in order])


if __name__ == "__main__":
    main()
