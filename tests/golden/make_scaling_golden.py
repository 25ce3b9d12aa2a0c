#!/usr/bin/env python3
"""Golden fixtures for the stages either side of ABEA, generated from the reference itself.

Runs only where /root/reference exists. For all 112 reads of test/ecoli_2kb_region (reads.blow5 + reads.fasta ->
pA -> reference getevents, exactly as make_golden.py does) it runs the UNMODIFIED reference
estimate_scalings_using_mom -> align -> postalign -> recalibrate_model (oracle/_ref) and pins the results to the
reference's own expected outputs:

* test/ecoli_2kb_region/est_scalings.exp        (DEBUG_ESTIMATED_SCALING dump: shift %.2f, scale %.2f per read)
* test/ecoli_2kb_region/recalib_scalings.exp    (DEBUG_RECALIB_SCALING dump: shift / scale / var %.2f per read)
* test/ecoli_2kb_region/eventalign.summary.exp  (per read NAME: shift, scale, var with three decimals)

and then writes tests/golden/scaling_golden.json: for the committed subset of reads (the ones already in
abea_golden.npz) and for seeded synthetic batches, the reference's per-read outputs — scalings as float bit
patterns, flags, counts, events_per_base, sha256 of the base_to_event_map.
"""
import hashlib, json, os, re, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
import oracle_lib as ol
from f5c_b200.batch import ReadBatch
from f5c_b200 import synth, models
import make_golden as mg

ECOLI = mg.ECOLI


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return [int(x) for x in np.asarray(a, dtype=np.float32).view(np.uint32)]


def record(batch, est, aln, sc):
    r = sc.res
    return dict(est_shift_bits=bits(est["shift"]), est_scale_bits=bits(est["scale"]),
                n_pairs=[int(x) for x in aln.n_pairs],
                shift_bits=bits(r["scalings"]["shift"]), scale_bits=bits(r["scalings"]["scale"]),
                var_bits=[b if c else None for b, c in zip(bits(r["scalings"]["var"]), r["calibrated"])],
                log_var_bits=[b if c else None for b, c in zip(bits(r["scalings"]["log_var"]), r["calibrated"])],
                flags=[int(x) for x in r["flags"]], calibrated=[int(x) for x in r["calibrated"]],
                n_event_alignment=[int(x) for x in r["n_event_alignment"]],
                num_m_state=[int(x) for x in r["num_m_state"]],
                events_per_base=[float(x) for x in r["events_per_base"]],
                map_sha256=[sha(sc.read_map(i)) if r["n_event_alignment"][i] > 0 else None
                            for i in range(batch.n_reads)])


def main():
    k, model = ol.ref_model(1)
    reads = mg.ecoli_reads(model, k)
    names = [r[0] for r in reads]
    b = ReadBatch.from_reads([r[1] for r in reads], [r[2] for r in reads], np.array([r[3] for r in reads]), k)
    est = ol.ref_estimate_scalings(b, model)
    assert est.tobytes() == ol.port_estimate_scalings(b, model).tobytes()
    assert np.array_equal(est["shift"], b.scalings["shift"]) and np.array_equal(est["scale"], b.scalings["scale"])
    aln = ol.ref_align(b, model)
    sc = ol.ref_scaling(b, model, aln)
    ol.assert_same_scaling(ol.port_scaling(b, model, aln), sc, "ecoli", check_var_d=False)

    # (1) eventalign.summary.exp, by read name
    n_sum = 0
    for line in open(os.path.join(ECOLI, "eventalign.summary.exp")):
        f = line.rstrip("\n").split("\t")
        if f[0] == "read_index" or f[1] not in names:
            continue
        i = names.index(f[1])
        g_shift, g_scale, g_var = float(f[10]), float(f[11]), float(f[13])
        s = sc.res["scalings"][i]
        assert abs(s["shift"] - g_shift) < 2e-3 and abs(s["scale"] - g_scale) < 2e-3 and abs(s["var"] - g_var) < 2e-3, \
            (f[1], s, g_shift, g_scale, g_var)
        n_sum += 1
    # (2) recalib_scalings.exp, as a multiset of %.2f triples
    gold = [tuple(re.findall(r"[-\d.]+", l)) for l in open(os.path.join(ECOLI, "recalib_scalings.exp"))]
    mine = {("%.2f" % s["shift"], "%.2f" % s["scale"], "%.2f" % s["var"])
            for s, c in zip(sc.res["scalings"], sc.res["calibrated"]) if c}
    n_recal = sum(1 for g in set(gold) if g in mine)
    # (3) est_scalings.exp: "... shift: X" / "... scale: Y" line pairs
    lines = open(os.path.join(ECOLI, "est_scalings.exp")).read().splitlines()
    gold_est = {(lines[j].split("shift: ")[1], lines[j + 1].split("scale: ")[1]) for j in range(0, len(lines) - 1, 2)}
    mine_est = {("%.2f" % e["shift"], "%.2f" % e["scale"]) for e in est}
    n_est = len(gold_est & mine_est)
    print("ecoli: reads", b.n_reads, "| summary lines matched by name:", n_sum, "| recalib_scalings.exp distinct",
          len(set(gold)), "matched", n_recal, "| est_scalings.exp distinct", len(gold_est), "matched", n_est)
    assert n_sum >= 100 and n_recal == len(set(gold)) and n_est == len(gold_est)

    out = {"ecoli_all": dict(n_reads=b.n_reads, summary_lines_matched=n_sum, recalib_distinct=len(set(gold)),
                             recalib_matched=n_recal, est_distinct=len(gold_est), est_matched=n_est)}
    G = json.load(open(os.path.join(HERE, "abea_golden.json")))
    pick = [names.index(n) for n in G["ecoli"]["names"]]
    sub = b.subset(pick)
    a2 = ol.ref_align(sub, model, 1)
    assert [int(x) for x in a2.n_pairs] == G["ecoli"]["n_pairs"]
    out["ecoli"] = record(sub, ol.ref_estimate_scalings(sub, model), a2, ol.ref_scaling(sub, model, a2))
    # golden summary values for the committed reads
    summ = {}
    for line in open(os.path.join(ECOLI, "eventalign.summary.exp")):
        f = line.rstrip("\n").split("\t")
        if f[1] in G["ecoli"]["names"]:
            summ[f[1]] = [float(f[10]), float(f[11]), float(f[13])]
    out["ecoli"]["summary_exp"] = [summ.get(n) for n in G["ecoli"]["names"]]

    for cfg, n, seed in (("cfg2", 24, 101), ("cfg3", 12, 102), ("cfg4", 4, 103)):
        sb = synth.make_config(cfg, seed=seed, n_reads=n)
        kk, m = ol.ref_model(models.MODELS[sb.meta["model"]][0])
        e = ol.ref_estimate_scalings(sb, m)
        a = ol.ref_align(sb, m)
        s = ol.ref_scaling(sb, m, a)
        ol.assert_same_scaling(ol.port_scaling(sb, m, a), s, cfg, check_var_d=False)
        out["synthetic_" + cfg] = dict(record(sb, e, a, s), n_reads=n, seed=seed)
    json.dump(out, open(os.path.join(HERE, "scaling_golden.json"), "w"), indent=1)
    print("wrote scaling_golden.json")


if __name__ == "__main__":
    main()
