"""N > 1 host logic on CPU: world_size-2 gloo run of the read-wise sharding + result gather that bench.py uses over
NCCL. Each rank materialises only its shard (LPT partition of one global length list), aligns it — here with the CPU
emulation build of the library, test infrastructure — and rank 0 receives every shard's pairs and checks them against
the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import oracle_lib as ol
    from f5c_b200 import models, synth
    from f5c_b200.abea import AbeaContext
    from f5c_b200.dist import compact_pairs, gather_results
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    synth.CONFIGS["tiny"] = dict(model="r9", n_reads=5, mean_events=300, sigma=0.6, epk=1.8)
    b = synth.make_config_shard("tiny", rank, world, seed=17)
    k, m = models.load_model("r9")
    with AbeaContext(0, lib_path=os.path.join(HERE, "simt", "libabea_emu.so")) as ctx:
        m = ctx.set_model(m, k)
        a = ctx.align_batch(b)
        # the exchange bench.py uses over NCCL: path codes + counts into persistent buffers, expanded on rank 0
        from f5c_b200.dist import ResultExchange
        ex = ResultExchange(rank, world, b.pair_capacity(), "cpu")
        res2 = [ex.gather(ctx) for _ in range(2)][-1]     # twice: the buffers are reused
        if rank == 0:
            res2 = [(c.numpy().copy(), p.numpy().copy().reshape(-1).view(a.pairs.dtype)) for c, p in res2]
    res = gather_results(a.n_pairs, compact_pairs(a.pairs, a.pair_ptr, a.n_pairs), rank, world, "cpu")
    if rank == 0:
        ok = len(res) == world and len(res2) == world
        for r in range(world):
            ok &= bool(np.array_equal(res[r][0], res2[r][0])) and bool(np.array_equal(res[r][1], res2[r][1]))
        total_reads = 0
        for r in range(world):
            br = synth.make_config_shard("tiny", r, world, seed=17)   # rank 0 can regenerate any shard to check it
            want = ol.port_align(br, m)
            c, p = res[r]
            ok &= bool(np.array_equal(c, want.n_pairs))
            ok &= bool(np.array_equal(p, compact_pairs(want.pairs, want.pair_ptr, want.n_pairs)))
            total_reads += br.n_reads
        ok &= total_reads == 5 * world
        open(out_path, "w").write("OK" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_align_and_gather(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "simt")], stderr=subprocess.DEVNULL)
    out = tmp_path / "result.txt"
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text() == "OK"


def test_lpt_shards_balance_band_counts():
    sys.path.insert(0, ROOT)
    from f5c_b200 import synth
    rng = np.random.default_rng(3)
    w = synth.draw_lengths(4096 * 8, 4000, 0.5, rng)
    shards = synth.lpt_shards(w, 8)
    loads = np.array([w[s].sum() for s in shards])
    assert sorted(np.concatenate(shards).tolist()) == list(range(4096 * 8))
    assert loads.max() / loads.min() < 1.001           # LPT balances the sum of band counts almost perfectly
    assert max(len(s) for s in shards) - min(len(s) for s in shards) < 200
