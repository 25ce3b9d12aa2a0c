"""BLOW5 records decoded on the device (f5c_b200/csrc/blow5_kernels.cuh) and the int16 front door of the
event detection. The checker is tests/blow5.py (Python's zlib + numpy restatements of slow5lib's svb-zd and ex-zd) on
fixtures written by the reference's own slow5lib (tests/golden/ecoli: a copy of test/ecoli_2kb_region/reads.blow5 — zlib,
v0.1.0 — and its first 8 records re-encoded with svb-zd / ex-zd signal compression by tests/golden/make_blow5_fixtures.py).
The CPU tests run the product kernels on the SIMT emulator; the GPU tests run all 112 records and compare the events
with the reference's golden event tables (tests/golden/ecoli_all.json)."""
import hashlib
import json
import os
import subprocess
import zlib

import numpy as np
import pytest

import blow5
import oracle_lib as ol
from f5c_b200 import models
from f5c_b200.abea import AbeaContext, AbeaError

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "simt", "libabea_emu.so")
ECOLI = os.path.join(HERE, "golden", "ecoli")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "simt")], stderr=subprocess.DEVNULL)
    return EMU


def payload_of(f: blow5.Blow5, idx):
    """The stored bytes of the chosen records back to back + (rec_ptr, rec_len): what a host reader hands over after
    walking the file's framing."""
    chunks = [f.record_bytes(i) for i in idx]
    rec_len = np.array([len(c) for c in chunks], dtype=np.int32)
    rec_ptr = np.zeros(len(chunks), dtype=np.int64)
    np.cumsum(rec_len[:-1].astype(np.int64), out=rec_ptr[1:])
    return np.frombuffer(b"".join(chunks), dtype=np.uint8).copy(), rec_ptr, rec_len


def check_decode(ctx, f, idx, f_truth=None, check_events=True):
    f_truth = f_truth or f
    payload, rec_ptr, rec_len = payload_of(f, idx)
    nev, ns, t = ctx.getevents_blow5(payload, rec_ptr, rec_len, f.record_method, f.signal_method)
    raw, raw_ptr = ctx.raw_download(ns)
    ev, ev_ptr = ctx.events_download(nev)
    for j, i in enumerate(idx):
        rid, dig, off, rng, sr, sig = f_truth.read(i)
        assert int(ns[j]) == len(sig), (i, "sample count")
        got = raw[int(raw_ptr[j]):int(raw_ptr[j]) + int(ns[j])]
        assert np.array_equal(got, sig.astype(np.float32)), (i, "decoded signal")
        if check_events:
            pa = ((sig.astype(np.float32) + np.float32(off)) * np.float32(np.float32(rng) / np.float32(dig))).astype(np.float32)
            want = ol.port_getevents(pa)
            assert int(nev[j]) == len(want), (i, "event count")
            assert ol._events_equal(ev[int(ev_ptr[j]):int(ev_ptr[j]) + int(nev[j])], want), (i, "events")
    return t


def test_python_checker_reads_the_slow5lib_fixtures():
    f0 = blow5.Blow5(os.path.join(ECOLI, "reads.blow5"))
    assert len(f0) == 112 and f0.record_method == blow5.RECORD_ZLIB and f0.signal_method == 0
    for name in ("ecoli8_zlib_svbzd.blow5", "ecoli8_none_svbzd.blow5", "ecoli8_zlib_exzd.blow5"):
        f = blow5.Blow5(os.path.join(ECOLI, name))
        assert len(f) == 8 and f.signal_method == (2 if "exzd" in name else 1)
        for i in range(8):
            a, b = f.read(i), f0.read(i)
            assert a[0] == b[0] and a[1:5] == b[1:5] and np.array_equal(a[5], b[5])


def test_emulated_inflate_and_events_zlib_records(emu):
    """zlib records (dynamic Huffman blocks as slow5lib's deflate level writes them): two short reads on the emulator."""
    f = blow5.Blow5(os.path.join(ECOLI, "reads.blow5"))
    order = np.argsort([r[1] for r in f.records])
    with AbeaContext(0, lib_path=emu) as ctx:
        check_decode(ctx, f, [int(order[0]), int(order[1])])


@pytest.mark.parametrize("name", ["ecoli8_zlib_svbzd.blow5", "ecoli8_none_svbzd.blow5"])
def test_emulated_svbzd_signal_decode(emu, name):
    f = blow5.Blow5(os.path.join(ECOLI, name))
    order = np.argsort([r[1] for r in f.records])
    with AbeaContext(0, lib_path=emu) as ctx:
        check_decode(ctx, f, [int(order[0]), int(order[2])], check_events=False)


def test_emulated_exzd_signal_decode(emu):
    """ex-zd signal compression (slow5lib >= 1.2; slow5_press.c:1262-1842) on a file written by slow5lib itself: hundreds of
    exceptions per read, their positions and values in streamvbyte."""
    f = blow5.Blow5(os.path.join(ECOLI, "ecoli8_zlib_exzd.blow5"))
    order = np.argsort([r[1] for r in f.records])
    with AbeaContext(0, lib_path=emu) as ctx:
        check_decode(ctx, f, [int(order[0]), int(order[2])], check_events=False)


def exzd_records(signals, rid=b"synthetic-read"):
    """Uncompressed BLOW5 records whose signal field is ex-zd (tests/blow5.py::ex_zd_encode, checked here against the
    decoder that reads slow5lib's own files)."""
    recs = []
    for sig in signals:
        enc = blow5.ex_zd_encode(sig)
        assert np.array_equal(blow5.ex_zd_decode(enc), sig)
        recs.append(np.uint16(len(rid) + 1).tobytes() + rid + b"\0" + np.uint32(0).tobytes() +
                    np.array([8192.0, 10.0, 1400.0, 4000.0], dtype="<f8").tobytes() + np.uint64(len(enc)).tobytes() + enc + b"aux")
    return recs


def test_emulated_exzd_edge_cases_and_errors(emu):
    """The branches slow5lib's real files rarely take: no exception at all, exactly one (stored inline), a common shift
    (all samples multiples of 8), a single sample, steps that wrap int16, exceptions at the first and last position;
    a truncated stream and one whose exception count lies must be reported."""
    rng = np.random.default_rng(5)
    smooth = (500 + np.cumsum(rng.integers(-20, 21, 700))).astype(np.int16)                     # no exception
    one = smooth.copy(); one[300:] += 400                                                       # exactly one
    shifted = ((60 + np.cumsum(rng.integers(-5, 6, 500))) * 8).astype(np.int16); shifted[100:] += 8 * 300
    wrap = np.array([32000, -32000, 32000, 5, -5, 30000, -30000], dtype=np.int16)               # every delta an exception
    ends = smooth[:200].copy(); ends[1:] += 900; ends[-1] -= 900
    sigs = [smooth, one, shifted, np.array([-77], dtype=np.int16), wrap, ends]
    recs = exzd_records(sigs)
    rec_len = np.array([len(r) for r in recs], dtype=np.int32)
    rec_ptr = np.zeros(len(recs), dtype=np.int64)
    np.cumsum(rec_len[:-1].astype(np.int64), out=rec_ptr[1:])
    payload = np.frombuffer(b"".join(recs), dtype=np.uint8).copy()
    with AbeaContext(0, lib_path=emu) as ctx:
        nev, ns, _ = ctx.getevents_blow5(payload, rec_ptr, rec_len, 0, 2)
        assert [int(x) for x in ns] == [len(s) for s in sigs]
        raw, raw_ptr = ctx.raw_download(ns)
        for j, s_ in enumerate(sigs):
            assert np.array_equal(raw[int(raw_ptr[j]):int(raw_ptr[j]) + len(s_)], s_.astype(np.float32)), j
        body = bytearray(recs[1])
        o = body.index(b"\0", 2) + 1 + 4 + 32      # the signal's length field; below: ten of the plain bytes missing
        short = bytes(body[:o]) + np.uint64(len(blow5.ex_zd_encode(one)) - 10).tobytes() + bytes(body[o + 8:-13]) + b"aux"
        with pytest.raises(AbeaError):
            ctx.getevents_blow5(np.frombuffer(short, dtype=np.uint8).copy(), np.zeros(1, dtype=np.int64),
                                np.array([len(short)], dtype=np.int32), 0, 2)
        lie = bytearray(recs[0]); e0 = o + 8 + 12                              # the exception count of a stream that has none
        lie[e0:e0 + 4] = np.uint32(3).tobytes()
        with pytest.raises(AbeaError):
            ctx.getevents_blow5(np.frombuffer(bytes(lie), dtype=np.uint8).copy(), np.zeros(1, dtype=np.int64),
                                np.array([len(lie)], dtype=np.int32), 0, 2)


def synth_records(levels, signals, rid=b"synthetic-read"):
    """BLOW5 v0.1.0-style records (no signal compression) built here, deflated at the given zlib levels — level 0 gives
    stored blocks, level 1 with Z_FIXED fixed-Huffman blocks."""
    recs = []
    for (level, strategy), sig in zip(levels, signals):
        body = (np.uint16(len(rid) + 1).tobytes() + rid + b"\0" + np.uint32(0).tobytes() +
                np.array([8192.0, 10.0, 1400.0, 4000.0], dtype="<f8").tobytes() + np.uint64(len(sig)).tobytes() +
                sig.astype("<i2").tobytes() + b"aux-fields-are-skipped")
        co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        recs.append(co.compress(body) + co.flush())
    return recs


def test_emulated_inflate_block_types_and_errors(emu):
    """Stored, fixed-Huffman and dynamic blocks, long matches (a constant stretch), an empty signal; a truncated stream
    and a corrupted one must be reported, not decoded."""
    rng = np.random.default_rng(3)
    sigs = [rng.integers(300, 700, 900).astype(np.int16),
            np.concatenate([np.full(700, 512, dtype=np.int16), rng.integers(-200, 900, 300).astype(np.int16)]),
            (500 + 40 * np.sin(np.arange(1500) / 7.0) + rng.normal(0, 3, 1500)).astype(np.int16),
            np.zeros(0, dtype=np.int16)]
    levels = [(0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY)]
    recs = synth_records(levels, sigs)
    rec_len = np.array([len(r) for r in recs], dtype=np.int32)
    rec_ptr = np.zeros(len(recs), dtype=np.int64)
    np.cumsum(rec_len[:-1].astype(np.int64), out=rec_ptr[1:])
    payload = np.frombuffer(b"".join(recs), dtype=np.uint8).copy()
    with AbeaContext(0, lib_path=emu) as ctx:
        nev, ns, _ = ctx.getevents_blow5(payload, rec_ptr, rec_len, 1, 0)
        assert [int(x) for x in ns] == [len(s) for s in sigs]
        raw, raw_ptr = ctx.raw_download(ns)
        for j, s in enumerate(sigs):
            assert np.array_equal(raw[int(raw_ptr[j]):int(raw_ptr[j]) + len(s)], s.astype(np.float32)), j
        bad = payload.copy()
        with pytest.raises(AbeaError):
            ctx.getevents_blow5(bad[:int(rec_len[0]) - 7], rec_ptr[:1], np.array([rec_len[0] - 7], dtype=np.int32), 1, 0)
        bad[int(rec_ptr[2]) + 40: int(rec_ptr[2]) + 60] ^= 0x5A
        with pytest.raises(AbeaError):
            ctx.getevents_blow5(bad, rec_ptr, rec_len, 1, 0)
        with pytest.raises(AbeaError):     # zstd records are not supported
            ctx.getevents_blow5(payload, rec_ptr, rec_len, 3, 0)


def test_emulated_int16_front_door(emu):
    """abea_signals_t.raw_i16: ADC counts as int16 give the same event tables as the float samples."""
    f = blow5.Blow5(os.path.join(ECOLI, "reads.blow5"))
    order = np.argsort([r[1] for r in f.records])
    recs = [f.read(int(order[j])) for j in (0, 1)]
    sig = np.concatenate([r[5] for r in recs])
    ns = np.array([len(r[5]) for r in recs], dtype=np.int32)
    ptr = np.array([0, ns[0]], dtype=np.int64)
    cal = tuple(np.array([r[j] for r in recs], dtype=np.float32) for j in (2, 3, 1))
    with AbeaContext(0, lib_path=emu) as ctx:
        e16, p16, n16, _ = ctx.getevents(sig, ptr, ns, cal)
        e32, p32, n32, _ = ctx.getevents(sig.astype(np.float32), ptr, ns, cal)
    assert np.array_equal(n16, n32) and e16.tobytes() == e32.tobytes()


# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gctx(built):
    c = AbeaContext(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_blow5_all_112_records_to_golden_events(gctx):
    """BLOW5 bytes -> events on the GPU for every record of the ecoli fixture, against the UNMODIFIED reference's event
    tables (tests/golden/ecoli_all.json: f5cref_getevents on slow5lib-decoded signals) and the Python-decoded signals."""
    gold = {r["name"]: r for r in json.load(open(os.path.join(HERE, "golden", "ecoli_all.json")))["reads"]}
    f = blow5.Blow5(os.path.join(ECOLI, "reads.blow5"))
    idx = list(range(len(f)))
    payload, rec_ptr, rec_len = payload_of(f, idx)
    nev, ns, t = gctx.getevents_blow5(payload, rec_ptr, rec_len, f.record_method, f.signal_method)
    raw, raw_ptr = gctx.raw_download(ns)
    ev, ev_ptr = gctx.events_download(nev)

    def event_sha(e):
        return hashlib.sha256(np.concatenate([e["start"].astype("<u8").view(np.uint8), e["length"].astype("<f4").view(np.uint8),
                                              e["mean"].astype("<f4").view(np.uint8), e["stdv"].astype("<f4").view(np.uint8)]).tobytes()).hexdigest()
    for j in idx:
        rid, dig, off, rng, sr, sig = f.read(j)
        assert np.array_equal(raw[int(raw_ptr[j]):int(raw_ptr[j]) + int(ns[j])], sig.astype(np.float32)), rid
        g = gold[rid]
        assert int(ns[j]) == g["n_samples"] and int(nev[j]) == g["n_events"], rid
        assert event_sha(ev[int(ev_ptr[j]):int(ev_ptr[j]) + int(nev[j])]) == g["events_sha256"], rid
    assert t["blow5_ms"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ecoli8_zlib_svbzd.blow5", "ecoli8_none_svbzd.blow5", "ecoli8_zlib_exzd.blow5"])
def test_gpu_blow5_svbzd(gctx, name):
    f = blow5.Blow5(os.path.join(ECOLI, name))
    check_decode(gctx, f, list(range(len(f))))


@pytest.mark.gpu
def test_gpu_int16_front_door_pinned(gctx):
    f = blow5.Blow5(os.path.join(ECOLI, "reads.blow5"))
    recs = [f.read(i) for i in range(0, 112, 7)]
    sig = np.concatenate([r[5] for r in recs])
    ns = np.array([len(r[5]) for r in recs], dtype=np.int32)
    ptr = np.zeros(len(recs), dtype=np.int64)
    np.cumsum(ns[:-1].astype(np.int64), out=ptr[1:])
    cal = tuple(np.array([r[j] for r in recs], dtype=np.float32) for j in (2, 3, 1))
    e16, p16, n16, _ = gctx.getevents(gctx.pin_array(sig), ptr, ns, cal)
    e32, p32, n32, _ = gctx.getevents(sig.astype(np.float32), ptr, ns, cal)
    assert np.array_equal(n16, n32) and e16.tobytes() == e32.tobytes()


def banded_aln_text(names, aln, flags):
    """The text of the reference's --print-banded-aln block (src/f5c.c:989-1006), formatted independently in Python."""
    out = []
    for i, nm in enumerate(names):
        if flags is not None and (int(flags[i]) & 0x002):
            continue
        out.append(">%s\tN_ALGN_PAIR:%d\t{ref_pos,read_pos}\n" % (nm, int(aln.n_pairs[i])))
        p = aln.read_pairs(i)
        out.append("".join("{%d,%d}\t" % (int(a), int(b)) for a, b in zip(p["ref_pos"], p["read_pos"])) + "\n")
    return "".join(out)


def test_pair_dump_matches_reference_format(emu, tmp_path):
    """abea_write_pairs on the oracle's pair lists of a few synthetic reads (one failing QC, one flagged as failed)."""
    from f5c_b200 import synth
    from f5c_b200.abea import write_pairs
    from edge_cases import edge_batch
    b = edge_batch()
    k, m = models.load_model("r9")
    want = ol.port_align(b, ol.full_model(m))
    names = ["read/%d" % i for i in range(b.n_reads)]
    flags = np.where(want.n_pairs > 0, 0, 2).astype(np.uint32)
    path = str(tmp_path / "aln.txt")
    write_pairs(path, names, want, flags, lib_path=emu)
    assert open(path).read() == banded_aln_text(names, want, flags)
    write_pairs(path, names, want, None, lib_path=emu)
    assert open(path).read() == banded_aln_text(names, want, None)


def resquiggle_text(names, read_len, n_samples, batch, sc, k, fmt, rna):
    """f5c resquiggle's output (src/resquiggle.c:322-447), formatted independently in Python from the oracle's maps."""
    out = ["read_id\tkmer_idx\tstart_raw_idx\tend_raw_idx\n"] if fmt == "tsv" else []
    for i, nm in enumerate(names):
        if int(sc.res["flags"][i]):
            continue
        ev = batch.read_events(i)
        nk = int(read_len[i]) - k + 1
        mp = [(int(a), int(b)) for a, b in sc.read_map(i)[:nk]]
        if rna:
            mp = [(b, a) for a, b in reversed(mp)]
        first, ci, d, cnt, matches, ss = True, 0, 0, 0, 0, []
        s2 = e2 = rs = re_ = -1
        for j, (a, b) in enumerate(mp):
            if a == -1:
                st = en = -1
                if not first:
                    d += 1
            else:
                st = int(ev["start"][a])
                if first:
                    s2, rs, ci, first = st, j, st, False
                en = e2 = int(ev["start"][b]) + int(ev["length"][b])
                re_ = j
                if fmt == "paf":
                    if d:
                        ss.append("%dD" % d); d = 0
                    if j == 0:
                        ci = st
                    mi = st - ci; ci += mi
                    if mi:
                        ss.append("%dI" % mi); cnt += mi
                    mi = en - st; ci += mi
                    if mi:
                        matches += 1; ss.append("%d," % mi); cnt += mi
            if fmt == "tsv":
                out.append("%s\t%d\t%s\t%s\n" % (nm, nk - j - 1 if rna else j, "." if st < 0 else str(st), "." if en < 0 else str(en)))
        if fmt == "paf":
            assert cnt == e2 - s2
            out.append("%s\t%d\t%d\t%d\t+\t%s\t%d\t%d\t%d\t%d\t%d\t255\tsc:f:%f\tsh:f:%f\tss:Z:%s\n" % (
                nm, int(n_samples[i]), s2, e2, nm, nk, nk - rs if rna else rs, nk - 1 - re_ if rna else re_ + 1, matches, nk,
                float(sc.res["scalings"]["scale"][0]), float(sc.res["scalings"]["shift"][0]), "".join(ss)))
    return "".join(out)


@pytest.mark.parametrize("fmt", ["tsv", "paf"])
@pytest.mark.parametrize("rna", [False, True])
def test_resquiggle_output_matches_reference_format(emu, tmp_path, fmt, rna):
    """abea_write_resquiggle on the oracle's chain (align -> scaling_single) of synthetic reads with skipped k-mers
    (deletions in the ss tag / '.' rows), one read that fails QC (not printed), both formats, DNA and the RNA reversal."""
    from f5c_b200 import synth
    from f5c_b200.abea import write_resquiggle
    b = synth.make_batch("r9", n_reads=6, mean_events=900, sigma=0.3, epk=1.8, seed=23, p_skip=0.08)
    b.events["mean"][b.event_ptr[2]:b.event_ptr[2] + b.n_events[2]] += 40.0      # read 2: far off the model -> fails
    k, m = models.load_model("r9")
    fm = ol.full_model(m)
    aln = ol.port_align(b, fm)
    sc = ol.port_scaling(b, fm, aln)
    assert int(sc.res["flags"][2]) != 0 and (sc.res["flags"] == 0).sum() >= 4
    names = ["read-%d" % i for i in range(b.n_reads)]
    n_samples = (b.events["start"][b.event_ptr + b.n_events - 1] + 30).astype(np.int64)
    if rna:   # an RNA event table is reversed before the alignment (src/f5c.c:713-721): time runs backwards along it
        for i in range(b.n_reads):
            e = b.events[int(b.event_ptr[i]):int(b.event_ptr[i]) + int(b.n_events[i])]
            e["start"] = (int(n_samples[i]) - e["start"].astype(np.int64) - e["length"].astype(np.int64)).astype(np.uint64)
    path = str(tmp_path / "rsq.txt")
    write_resquiggle(path, names, b.read_len, n_samples, b.events, b.event_ptr, sc.res, sc.maps, sc.map_ptr, k, fmt=fmt,
                     rna=rna, lib_path=emu)
    got = open(path).read()
    assert got == resquiggle_text(names, b.read_len, n_samples, b, sc, k, fmt, rna)
    assert got.count("\n") > (1000 if fmt == "tsv" else 3) and ("." in got or fmt == "paf") and ("D" in got or fmt == "tsv")


@pytest.mark.gpu
def test_gpu_blow5_to_banded_aln_dump(built, tmp_path):
    """tools/blow5_eventalign_dump.py: BLOW5 + FASTA -> the --print-banded-aln text, every stage on the GPU; every read's
    pair list in the text must be the one the UNMODIFIED reference produced (tests/golden/ecoli_all.json)."""
    import sys
    from f5c_b200.batch import PAIR_DTYPE
    out = str(tmp_path / "dump.txt")
    subprocess.check_call([sys.executable, os.path.join(os.path.dirname(HERE), "tools", "blow5_eventalign_dump.py"),
                           os.path.join(ECOLI, "reads.blow5"), os.path.join(ECOLI, "reads.fasta"), out])
    gold = {r["name"]: r for r in json.load(open(os.path.join(HERE, "golden", "ecoli_all.json")))["reads"]}
    lines = open(out).read().split("\n")
    seen = 0
    for hdr, body in zip(lines[0::2], lines[1::2]):
        if not hdr:
            continue
        name, npair, _ = hdr[1:].split("\t")
        n = int(npair.split(":")[1])
        toks = [t for t in body.split("\t") if t]
        assert len(toks) == n == gold[name]["n_pairs"], name
        arr = np.array([tuple(int(x) for x in t[1:-1].split(",")) for t in toks], dtype=PAIR_DTYPE)
        assert hashlib.sha256(arr.tobytes()).hexdigest() == gold[name]["pairs_sha256"], name
        seen += 1
    assert seen == sum(1 for r in gold.values() if not (r["flags"] & 2))


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["tsv", "paf"])
def test_gpu_blow5_to_resquiggle(built, tmp_path, fmt):
    """tools/blow5_resquiggle.py: BLOW5 + FASTA -> f5c resquiggle's TSV / PAF, every stage on the GPU; the rows of the six
    shortest reads must be what the CPU chain (oracle getevents -> estimate -> align -> scaling_single) formats to."""
    import sys
    from f5c_b200.batch import ReadBatch, SCALINGS_DTYPE
    out = str(tmp_path / "rsq.txt")
    subprocess.check_call([sys.executable, os.path.join(os.path.dirname(HERE), "tools", "blow5_resquiggle.py"),
                           os.path.join(ECOLI, "reads.blow5"), os.path.join(ECOLI, "reads.fasta"), out] + (["-c"] if fmt == "paf" else []))
    f = blow5.Blow5(os.path.join(ECOLI, "reads.blow5"))
    seqs = dict(blow5.read_fasta(os.path.join(ECOLI, "reads.fasta")))
    k, m = models.load_model("r9")
    fm = ol.full_model(m)
    order = [int(i) for i in np.argsort([r[1] for r in f.records])[:6]]
    recs = [f.read(i) for i in order]
    evs = []
    for rid, dig, off, rng, sr, sig in recs:
        pa = ((sig.astype(np.float32) + np.float32(off)) * np.float32(np.float32(rng) / np.float32(dig))).astype(np.float32)
        evs.append(ol.port_getevents(pa))
    b = ReadBatch.from_reads([seqs[r[0]].encode() for r in recs], evs, np.zeros(len(recs), dtype=SCALINGS_DTYPE), k)
    b.scalings = ol.port_estimate_scalings(b, fm)
    aln = ol.port_align(b, fm)
    sc = ol.port_scaling(b, fm, aln)
    names = [r[0] for r in recs]
    want = resquiggle_text(names, b.read_len, [len(r[5]) for r in recs], b, sc, k, fmt, False)
    got = open(out).read()
    if fmt == "tsv":
        assert got.startswith("read_id\tkmer_idx\tstart_raw_idx\tend_raw_idx\n")
        rows = {}
        for line in got.split("\n")[1:]:
            if line:
                rows.setdefault(line.split("\t")[0], []).append(line)
        for nm in names:
            mine = [l for l in want.split("\n")[1:] if l.startswith(nm + "\t")]
            assert rows.get(nm, []) == mine, nm
    else:
        lines = {l.split("\t")[0]: l for l in got.split("\n") if l}
        for l in want.split("\n"):
            if l:    # the sc / sh fields are the BATCH's first read's (a reference quirk): compare everything else
                a, b_ = l.split("\t"), lines[l.split("\t")[0]].split("\t")
                assert a[:12] == b_[:12] and a[14] == b_[14], a[0]

