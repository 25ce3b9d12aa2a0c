"""CPU tests of the PRODUCT sources: f5c_b200/csrc/{abea_host.cu,abea_kernels.cuh} compiled unchanged for the
lock-step SIMT emulator (tests/simt, test infrastructure) must agree bit-exactly with the oracle. This covers the
kernels' lane mapping, shuffles, window sliding, trace packing, traceback and the host packer/unpacker without a GPU.
The sizes are tiny because every warp shuffle costs ~100 fiber switches."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from edge_cases import edge_batch
from f5c_b200 import models, synth
from f5c_b200.abea import AbeaContext

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "simt", "libabea_emu.so")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "simt")], stderr=subprocess.DEVNULL)
    return EMU


def check(emu, batch, model_name, what):
    k, m = models.load_model(model_name)
    with AbeaContext(0, lib_path=emu) as ctx:
        m = ctx.set_model(m, k)
        got = ctx.align_batch(batch)
        st = ctx.read_stats(batch.n_reads)
        # the same batch handed over as a flat array of event means (abea_batch_t.event_means, 4 B per event)
        gm = ctx.align_batch(batch, means=batch.event_means())
        stm = ctx.read_stats(batch.n_reads)
    want = ol.port_align(batch, m)
    ol.assert_same_alignment(got, want, what)
    ol.assert_same_alignment(gm, want, what + " (means only)")
    sched = want.stats["n_bands"] > 0
    for s_ in (st, stm):
        assert np.array_equal(s_["sum_emission"][sched], want.stats["sum_emission"][sched])
        assert np.array_equal(s_["end_event"][sched], want.stats["end_event"][sched])
        assert np.array_equal(s_["max_gap"][sched], want.stats["max_gap"][sched])
    assert gm.timing["h2d_bytes"] < got.timing["h2d_bytes"] or batch.events.shape[0] == 0
    return got


@pytest.mark.parametrize("wide", ["0", "1"])
@pytest.mark.parametrize("name,kw", [
    ("r9", dict(n_reads=6, mean_events=600, sigma=0.5, epk=1.8, seed=3)),
    ("r10", dict(n_reads=4, mean_events=700, sigma=1.0, epk=1.9, seed=5)),
    ("rna004", dict(n_reads=2, mean_events=1200, sigma=0.5, epk=2.5, seed=6)),
    ("r9", dict(n_reads=9, mean_events=120, sigma=0.8, epk=1.8, seed=7, min_len=20)),
])
def test_emulated_kernels_match_oracle(emu, monkeypatch, name, kw, wide):
    monkeypatch.setenv("ABEA_STREAM", "3" if kw["seed"] % 2 else "0")
    """wide=0: everything through the narrow (warp per read) kernel; wide=1: the scheduler's own split (a batch with
    fewer reads than SMs runs wide, a larger one mixes wide, long-narrow and regular reads)."""
    monkeypatch.setenv("ABEA_WIDE", wide)
    got = check(emu, synth.make_batch(name, **kw), name, name)
    assert (got.n_pairs > 0).any()
    if wide == "0":
        assert got.timing["n_wide"] == 0


@pytest.mark.parametrize("stream", ["0", "1", "2", "3", "4", "5", "7"])
@pytest.mark.parametrize("wide", ["0", "1"])
def test_emulated_edge_cases(emu, monkeypatch, wide, stream):
    """ABEA_STREAM bit 0: events through abea_load_kernel (which also range-checks them: read 0 has an event that must
    send it to the exact instantiation); bit 1: pair lists written whole to the caller's buffer by the traceback; bit 2
    (wins over bit 1): pair lists as path codes, expanded by host threads; 0: staged copies."""
    monkeypatch.setenv("ABEA_WIDE", wide)
    monkeypatch.setenv("ABEA_STREAM", stream)
    b = edge_batch()
    got = check(emu, b, "r9", "edge")
    assert (got.timing["streamed"] & 6) == (4 if int(stream) & 4 else int(stream) & 2)
    if b.events.ctypes.data % 16 == 0:
        assert (got.timing["streamed"] & 1) == (int(stream) & 1)
    assert got.n_pairs[1] == 0 and got.n_pairs[3] == 0 and got.n_pairs[4] == 0
    assert got.n_pairs[0] > 0 and got.n_pairs[2] > 0


def test_emulated_resident_phases(emu):
    """upload / run (twice) / download give the same answer as the one-shot call."""
    b = synth.make_batch("r9", n_reads=3, mean_events=300, sigma=0.4, epk=1.8, seed=13)
    k, m = models.load_model("r9")
    with AbeaContext(0, lib_path=emu) as ctx:
        m = ctx.set_model(m, k)
        one = ctx.align_batch(b)
        t = ctx.upload(b)
        assert t["n_scheduled"] == 3 and t["n_events"] == int(b.n_events.sum())
        ctx.run()
        t = ctx.run()
        assert t["kernel_launches"] >= 3
        two = ctx.download(b)
    ol.assert_same_alignment(one, two, "resident")


def test_emulated_wide_kernel_all_reads(emu, monkeypatch):
    """Force every read through the wide (4 warps per read) fill kernel, including the edge cases."""
    monkeypatch.setenv("ABEA_WIDE", "1")
    b = synth.make_batch("r10", n_reads=3, mean_events=800, sigma=0.8, epk=1.9, seed=15)
    k, m = models.load_model("r10")
    with AbeaContext(0, lib_path=emu) as ctx:
        m = ctx.set_model(m, k)
        got = ctx.align_batch(b)
    assert got.timing["n_wide"] >= 1
    ol.assert_same_alignment(got, ol.port_align(b, m), "wide")
    check(emu, edge_batch(), "r9", "wide edge")


def test_emulated_narrow_only(emu, monkeypatch):
    monkeypatch.setenv("ABEA_WIDE", "0")
    b = synth.make_batch("r9", n_reads=4, mean_events=1500, sigma=0.5, epk=1.8, seed=16)
    got = check(emu, b, "r9", "narrow only")
    assert got.timing["n_wide"] == 0


def test_emulated_degenerate_batches(emu):
    """No read scheduled at all (everything filtered), and an empty batch: counts are zero, nothing crashes."""
    import numpy as np
    from f5c_b200.batch import ReadBatch
    b = synth.make_batch("r9", n_reads=3, mean_events=200, sigma=0.2, epk=1.8, seed=19)
    b.good[:] = 0
    k, m = models.load_model("r9")
    with AbeaContext(0, lib_path=emu) as ctx:
        ctx.set_model(m, k)
        got = ctx.align_batch(b)
        assert got.n_pairs.tolist() == [0, 0, 0] and got.timing["n_scheduled"] == 0
        empty = ReadBatch.from_reads([], [], np.zeros(0, dtype=b.scalings.dtype), k)
        got = ctx.align_batch(empty)
        assert got.n_pairs.shape == (0,)
        b.good[:] = 1            # the same context keeps working afterwards
        got = ctx.align_batch(b)
        assert (got.n_pairs > 0).all()


def test_smoke_entry_point_on_the_emulator(emu):
    """__graft_entry__.smoke() — the driver's GPU smoke test — with the emulator build in place of the CUDA library:
    both of its legs (alignment; raw signal -> events -> scalings -> alignment -> recalibration) stay runnable."""
    import __graft_entry__ as g
    g.smoke(lib_path=emu)


@pytest.mark.parametrize("stream", ["5", "3", "0"])
@pytest.mark.parametrize("threads", [1, 3])
def test_emulated_ragged_front_door(emu, monkeypatch, stream, threads):
    """abea_align_ragged: the batch as db_t holds it (one allocation per read); packer threads publish the means piece
    by piece to the loader, unpacker threads copy each pair list out when its count appears. Same answer as the flat
    call, with and without the overlap (ABEA_STREAM=0: plain pack -> copy engine -> unpack)."""
    monkeypatch.setenv("ABEA_STREAM", stream)
    monkeypatch.setenv("ABEA_LOAD_PIECE_KB", "1")   # many pieces per read, so that the list order matters
    k, m = models.load_model("r9")
    for b in (synth.make_batch("r9", n_reads=7, mean_events=500, sigma=0.7, epk=1.8, seed=21), edge_batch()):
        with AbeaContext(0, lib_path=emu) as ctx:
            m = ctx.set_model(m, k)
            got = ctx.align_ragged(b, threads=threads)
            again = ctx.align_ragged(b, threads=threads)      # staging is reused
        want = ol.port_align(b, m)
        ol.assert_same_alignment(got, want, "ragged")
        ol.assert_same_alignment(again, want, "ragged, second batch")
        assert got.timing["streamed"] == int(stream)


@pytest.mark.parametrize("tb", ["1", "0"])
def test_emulated_path_codes(emu, monkeypatch, tb):
    """Pair lists leave the device as path codes (first pair + two bit planes per 32 steps) and host threads expand them:
    lists long enough for several 32-word stores plus a tail, from both traceback forms, with 1 and 3 threads; with 0
    threads the lists are copied back whole. 8 bytes per 32 pairs cross the boundary."""
    monkeypatch.setenv("ABEA_TB", tb)
    b = synth.make_batch("r9", n_reads=5, mean_events=2600, sigma=0.5, epk=1.8, seed=77)
    k, m = models.load_model("r9")
    with AbeaContext(0, lib_path=emu) as ctx:
        m = ctx.set_model(m, k)
        want = ol.port_align(b, m)
        assert int(want.n_pairs.max()) > 3 * 1024 + 40
        for threads in (1, 3):
            assert ctx.host_threads(threads) == threads
            got = ctx.align_batch(b)
            ol.assert_same_alignment(got, want, f"path codes, {threads} threads")
            assert got.timing["streamed"] & 4
            assert got.timing["d2h_bytes"] < int(want.n_pairs.sum()) * 8 // 16
        # the same codes stay in device memory for consumers on the GPU (the multi-GPU exchange): expanded back to dense
        # lists by abea_expand_codes they are the compacted pair lists (under the emulator device memory is host memory)
        from f5c_b200.dist import compact_pairs
        _dp, dn, cap, n = ctx.device_results()
        dc, n_words = ctx.device_codes()
        cap_ptr = np.concatenate([[0], np.cumsum(b.pair_capacity().astype(np.int64))])
        assert n == b.n_reads and cap == cap_ptr[-1] and n_words == (cap >> 5) + 2 * n + 2
        dense = np.zeros((int(cap), 2), dtype=np.int32)
        total = ctx.expand_codes(dc, dn, cap_ptr, dense.ctypes.data, int(cap), sync=True)
        assert total == int(want.n_pairs.sum())
        assert np.array_equal(dense[:total].reshape(-1).view(want.pairs.dtype), compact_pairs(want.pairs, want.pair_ptr, want.n_pairs))
        assert ctx.host_threads(0) == 0 and ctx.host_threads() == 0
        got = ctx.align_batch(b)
        ol.assert_same_alignment(got, want, "whole lists")
        assert (got.timing["streamed"] & 4) == 0 and got.timing["d2h_bytes"] >= int(want.n_pairs.sum()) * 8


@pytest.mark.parametrize("tb,margin", [("0", "64"), ("1", "0"), ("1", "7"), ("1", "200"), ("3", "64")])
def test_emulated_traceback_forms(emu, monkeypatch, tb, margin):
    """The traceback in its three forms must give the same pairs, the same QC sum (bit for bit), the same max_gap:
    ABEA_TB=0 the serial walk; 1 the segment-parallel walk (a walk per lane from a speculative entry cell, verified
    against the true path) with margins from 0 (every segment mis-speculates and is corrected) to longer than the
    reads (every lane walks from the end cell); 3 the parallel walk with the emissions summed in traceback order (the
    path taken when the partial sums could round)."""
    monkeypatch.setenv("ABEA_TB", tb)
    monkeypatch.setenv("ABEA_TB_MARGIN", margin)
    for wide in ("0", "1"):
        monkeypatch.setenv("ABEA_WIDE", wide)
        check(emu, synth.make_batch("r9", n_reads=5, mean_events=900, sigma=0.6, epk=1.8, seed=41), "r9", f"tb {tb}/{margin}")
        check(emu, synth.make_batch("r9", n_reads=8, mean_events=130, sigma=0.8, epk=1.8, seed=42, min_len=20), "r9", "short")
        check(emu, edge_batch(), "r9", "edge")
    # skip-heavy reads: long runs of FROM_L cross segment borders (max_gap is assembled from per-lane runs)
    monkeypatch.setenv("ABEA_WIDE", "0")
    check(emu, synth.make_batch("r9", n_reads=4, mean_events=700, sigma=0.3, epk=1.8, seed=43, p_skip=0.35), "r9", "skips")
