#!/usr/bin/env python3
"""bench.py — events aligned/sec of the ABEA hot path on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps K --warmup W                     # our CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus N --steps K --warmup W    # the reference's own CPU align() on host cores

A "step" is one pass of the hot path over one synthetic batch. The default workload is the north_star target (BASELINE
configs[4], "cfg5"): R10.4.1, 4096 reads per GPU, mean 4k events/read, log-normal lengths, read-sharded over the GPUs
(at N = 8 that is the 32768-read batch of configs[4]); --config selects cfg2 / cfg3 / cfg4 instead. `value` is
device-timed (CUDA events inside the library, on the stream the kernels are launched on) with the batch already
resident in HBM. `e2e` is the same metric through the C-ABI call abea_align_batch with HOST (pinned) buffers in and out —
everything inside the timed region: the event means (4 bytes per event, abea_batch_t.event_means) are pulled over PCIe
by abea_load_kernel while the fill runs and the pair lists are written straight into the caller's pinned buffer by the
traceback — plus, for N > 1, the NCCL exchange of all ranks' results to rank 0. `e2e_dropin` is the same batch through
the reference's own plug-in call align_cuda(core_t*, db_t*) on a ragged db_t (every read its own allocations).
At N = 1 the line also carries `configs` (cfg2 / cfg3 / cfg4 at full size, device-timed, each with its own roofline
fraction and a parity check against the CPU reference), `gpu_reference` (the reference's own CUDA kernels compiled for
sm_100a, same batch, same box) and `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "events aligned/sec (ABEA kernel, device-timed)"
UNIT = "events/s"
DTYPE = "f32 emission + f64 transition sums (bit-exact vs CPU)"
DROPIN_SO = os.path.join(ROOT, "f5c_b200", "lib", "libf5c_abea_dropin.so")
REFGPU_SO = os.path.join(ROOT, "oracle", "_ref", "libf5c_refgpu.so")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg5", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--reads-per-gpu", type=int, default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-sample-events", type=float, default=0, help="events in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip configs / gpu_reference / e2e_dropin / stages (profiling runs)")
    ap.add_argument("--dropin-threads", type=int, default=0, help="core->opt.num_thread of the e2e_dropin leg (0 = min(16, cores))")
    return ap.parse_args()


def config_dict(a, world: int) -> dict:
    """The `config` object — identical in both arms (ours and --impl reference), so that the driver can match them."""
    from f5c_b200 import synth
    p = synth.CONFIGS[a.config]
    per = a.reads_per_gpu if a.reads_per_gpu is not None else p["n_reads"]
    model = {"r9": "R9.4.1 DNA (k=6)", "r10": "R10.4.1 DNA (k=9)", "rna004": "RNA004 (k=9)"}[p["model"]]
    return {"workload": f"{a.config}: synthetic {model} batch, {per * world} reads ({per}/GPU), mean {p['mean_events']} "
                        f"events/read log-normal sigma {p['sigma']}, {p['epk']} events per base, bandwidth 100",
            "global_reads": per * world, "reads_per_gpu": per,
            "parallelism": f"read-sharded x{world} (LPT by band count), NCCL result exchange only",
            "l2": "inputs larger than L2 (trace 32 B/band + pairs 8 B/event + means 4 B/event per step >> 126 MB); no flush",
            "seed": a.seed}


# ---------------------------------------------------------------------------------------------------------------------
def committed_capture(config: str):
    """DRAM bytes and issue statistics of the dominant kernel from the committed `ncu --set full` capture of THIS
    config (profiles/traffic_r02.json, written by tools/ncu_summary.py from the .ncu-rep of the same bench command).
    Not measured in this run — ncu cannot run inside a timed bench — so the line says where it comes from."""
    p = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d if d.get("config") == config else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_sample(batch, target_events: float):
    """A bounded prefix-by-count sample of the same workload for the CPU arm."""
    order = np.arange(batch.n_reads)
    csum = np.cumsum(batch.n_events[order].astype(np.int64))
    n = int(np.searchsorted(csum, target_events) + 1)
    n = max(8, min(batch.n_reads, n))
    return batch.subset(order[:n]), n


def run_cpu(batch, model, threads: int):
    """The reference's CPU branch of align_db (oracle/_ref when it was built from the reference, else our C port)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    if ol.have_ref():
        r = ol.ref_align(batch, model, threads)
        return r, "reference"
    r = ol.port_align(batch, model, threads)
    return r, "port"


def same_pairs(got_pairs, got_ptr, got_n, want, idx_got, idx_want=None) -> bool:
    """Bit-exact comparison of the pair lists of reads idx_got of one result with reads idx_want of another."""
    idx_want = range(len(idx_got)) if idx_want is None else idx_want
    for ig, iw in zip(idx_got, idx_want):
        n = int(got_n[ig])
        if n != int(want.n_pairs[iw]):
            return False
        p = int(got_ptr[ig])
        if not np.array_equal(got_pairs[p:p + n], want.read_pairs(iw)):
            return False
    return True


def parity_sample(batch, n_longest: int, n_random: int, seed: int):
    rng = np.random.default_rng(seed)
    longest = np.argsort(batch.n_events, kind="stable")[-n_longest:]
    rest = np.setdiff1d(np.arange(batch.n_reads), longest)
    rnd = rng.choice(rest, min(n_random, len(rest)), replace=False)
    return np.concatenate([rnd, longest]).astype(np.int64)


# ---------------------------------------------------------------------------------------------------------------------
def time_resident(ctx, steps: int, warmup: int):
    for _ in range(warmup):
        ctx.run()
    tot = dict(kernel_ms=0.0, fill_ms=0.0, kmer_ms=0.0, launches=0)
    for _ in range(steps):
        t = ctx.run()
        tot["kernel_ms"] += t["kernel_ms"]; tot["fill_ms"] += t["fill_ms"]; tot["kmer_ms"] += t["kmer_ms"]
        tot["launches"] += t["kernel_launches"]
    return tot, t


def other_configs(ctx_factory, a, peak, cores):
    """cfg2 / cfg3 / cfg4 (BASELINE configs[1..3]) at full size on this GPU: device-timed, roofline fraction, and the
    pair lists of the 16 longest + 16 random reads against the CPU reference (bit-exact)."""
    from f5c_b200 import models, synth
    out = []
    for name in ("cfg2", "cfg3", "cfg4"):
        if name == a.config:
            continue
        b = synth.make_config(name, seed=a.seed)
        k, model = models.load_model(b.meta["model"])
        with ctx_factory() as ctx:
            model = ctx.set_model(model, k)
            pb = ctx.pin_batch(b)
            ctx.upload(pb)
            tot, last = time_resident(ctx, 3, 2)
            res = ctx.download(pb)
            idx = parity_sample(b, 16, 16, a.seed + 1)
            sub = b.subset(idx)
            want, kind = run_cpu(sub, model, cores)
            ok = same_pairs(res.pairs, res.pair_ptr, res.n_pairs, want, idx)
        ms = tot["kernel_ms"] / 3
        alg = float(b.algorithmic_bytes(res.n_pairs))
        ev = b.events_aligned()
        out.append({"config": name, "workload": config_dict(argparse.Namespace(config=name, reads_per_gpu=None, seed=a.seed), 1)["workload"],
                    "events_per_step": ev, "bands_per_step": int(b.n_bands[b.eligible()].sum()),
                    "ms_per_step": ms, "value": ev / (ms * 1e-3), "unit": UNIT, "steps": 3, "warmup": 2,
                    "n_wide": last["n_wide"], "longest_read_bands": int(b.n_bands.max()),
                    "roofline": {"achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "frac": alg / (ms * 1e-3) / 1e9 / peak,
                                 "unit": "GB/s", "algorithmic_bytes": int(alg)},
                    "parity_on_cpu_sample": bool(ok), "parity_sample": f"16 longest + 16 random reads vs {kind} align()",
                    "reads_aligned_fraction": float((res.n_pairs > 0).mean())})
        del b, pb, res
    return out


def gpu_reference(batch, model, k, device, steps, warmup, ours_pairs, ours_ptr, ours_n):
    """The reference's own CUDA kernels (src/align.cu) compiled for sm_100a, same batch, same GPU."""
    if not os.path.exists(REFGPU_SO):
        return {"unavailable": "oracle/_ref/libf5c_refgpu.so not built (needs the f5c tree at build time)"}
    from f5c_b200.batch import CBatch, PAIR_DTYPE
    lib = ctypes.CDLL(REFGPU_SO)
    vp = ctypes.c_void_p
    lib.f5cref_gpu_bytes.restype = ctypes.c_int64
    lib.f5cref_gpu_bytes.argtypes = [ctypes.POINTER(CBatch)]
    lib.f5cref_gpu_align_batch.argtypes = [ctypes.POINTER(CBatch), vp, ctypes.c_uint32, ctypes.c_int, vp, vp, vp,
                                           ctypes.c_int, ctypes.c_int, vp, vp]
    cb = batch.as_c()
    need = int(lib.f5cref_gpu_bytes(ctypes.byref(cb)))
    pairs = np.zeros(int(batch.pair_capacity().sum()), dtype=PAIR_DTYPE)
    n_pairs = np.zeros(batch.n_reads, dtype=np.int32)
    pp = batch.pair_ptr()
    kms = np.zeros(steps, dtype=np.float64)
    ems = np.zeros(steps, dtype=np.float64)
    rc = lib.f5cref_gpu_align_batch(ctypes.byref(cb), model.ctypes.data, k, device, pairs.ctypes.data, pp.ctypes.data,
                                    n_pairs.ctypes.data, warmup, steps, kms.ctypes.data, ems.ctypes.data)
    if rc != 0:
        return {"unavailable": f"reference kernels failed (rc {rc}; the layout needs {need / 1e9:.1f} GB of device memory)"}
    ev = batch.events_aligned()
    el = np.flatnonzero(batch.eligible())
    eq = sum(1 for i in el if int(n_pairs[i]) == int(ours_n[i]) and
             np.array_equal(pairs[int(pp[i]):int(pp[i]) + int(n_pairs[i])], ours_pairs[int(ours_ptr[i]):int(ours_ptr[i]) + int(ours_n[i])]))
    return {"value": ev / (kms.mean() * 1e-3), "unit": UNIT, "ms_per_step": float(kms.mean()),
            "e2e_value": ev / (ems.mean() * 1e-3), "e2e_ms_per_step": float(ems.mean()), "steps": steps, "warmup": warmup,
            "device_bytes": need, "pairs_equal_fraction": eq / max(1, len(el)),
            "kernels": "align_kernel_pre_2d + align_kernel_core_2d_shm + align_kernel_post (reference src/align.cu, nvcc -O2 "
                       "sm_100a, launch shapes of src/f5c.cu:910-960), every read on the GPU (no CPU side-pool), whole batch in one launch",
            "note": "device-timed with CUDA events first kernel -> last kernel; e2e adds the reference's memset, seven blocking "
                    "H2D copies from pageable memory, two D2H copies and the host-side reversal (src/f5c.cu:832-1030). Its pairs "
                    "are not bit-equal to the CPU align() (float transition constants, FMA contraction: SURVEY 2b)"}


def e2e_dropin(batch, model, k, device, threads, steps, warmup, want, sample_n):
    """align_cuda(core_t*, db_t*) — the reference's plug-in call — on a ragged db_t built once (every read's sequence,
    event table and pair buffer its own allocation), wall-clocked per call inside the drop-in's bench door."""
    if not os.path.exists(DROPIN_SO):
        return {"unavailable": "drop-in not built (needs the f5c headers at build time)"}
    from f5c_b200.batch import CBatch, PAIR_DTYPE
    lib = ctypes.CDLL(DROPIN_SO)
    vp = ctypes.c_void_p
    lib.f5c_dropin_bench.argtypes = [ctypes.POINTER(CBatch), vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, vp, vp, vp, vp]
    pairs = np.zeros(int(batch.pair_capacity().sum()), dtype=PAIR_DTYPE)
    n_pairs = np.zeros(batch.n_reads, dtype=np.int32)
    pp = batch.pair_ptr()
    ms = np.zeros(steps, dtype=np.float64)
    cb = batch.as_c()
    rc = lib.f5c_dropin_bench(ctypes.byref(cb), model.ctypes.data, k, device, threads, warmup, steps, ms.ctypes.data,
                              pairs.ctypes.data, pp.ctypes.data, n_pairs.ctypes.data)
    if rc != 0:
        return {"unavailable": f"f5c_dropin_bench failed ({rc})"}
    ev = batch.events_aligned()
    d = {"value": ev / (ms.mean() * 1e-3), "unit": UNIT, "ms_per_step": float(ms.mean()), "ms_min": float(ms.min()),
         "steps": steps, "warmup": warmup, "opt.num_thread": threads,
         "call": "align_cuda(core_t*, db_t*) of libf5c_abea_dropin.so on a ragged db_t (per-read malloc'd sequences, event tables "
                 "and pair buffers); timed per call with the reference's realtime()",
         "h2d_bytes_per_step": int(4 * batch.n_events.astype(np.int64).sum() + batch.seq.shape[0]),
         # the lists cross PCIe as path codes (one 8-byte word for the first pair + one per 32 steps) unless ABEA_STREAM says otherwise
         "d2h_bytes_per_step": int((8 * (1 + (np.maximum(n_pairs.astype(np.int64) - 1, 0) + 31) // 32) * (n_pairs > 0)).sum()
                                   + 4 * batch.n_reads) if (int(os.environ.get("ABEA_STREAM", "7")) & 4)
                               else int(8 * n_pairs.astype(np.int64).sum() + 4 * batch.n_reads)}
    if want is not None:
        d["parity_on_cpu_sample"] = bool(same_pairs(pairs, pp, n_pairs, want, range(sample_n)))
    return d, pairs, n_pairs


# ---------------------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        return reference_arm(a, rank, world)

    import torch
    import torch.distributed as dist
    from f5c_b200 import models, synth
    from f5c_b200.abea import AbeaContext
    from f5c_b200.dist import ResultExchange

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; f5c_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch = synth.make_config_shard(a.config, rank, world, seed=a.seed, reads_per_gpu=a.reads_per_gpu)
    k, model = models.load_model(batch.meta["model"])
    ctx = AbeaContext(local_rank)
    # host threads that expand the path codes: the box's CPUs are shared by the ranks of this node (one process per GPU),
    # so a rank takes its share minus the thread that drives the GPU (the library's own default, min(8, CPUs), is for a
    # process that has the box to itself); ABEA_HOST_THREADS overrides
    # A host core expands ~0.8 G pairs/s, so fewer than four threads cannot keep up with the kernels
    # (profiles/path_codes_r02.txt): a rank whose share is smaller has the traceback write whole lists instead (0 threads)
    if "ABEA_HOST_THREADS" not in os.environ:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        share = len(os.sched_getaffinity(0)) // max(1, local_world) - 1
        ctx.host_threads(min(8, share) if share >= 4 else 0)
    model = ctx.set_model(model, k)
    sm_count, dev_name = ctx.device_info()
    pinned = ctx.pin_batch(batch)
    pinned_means = ctx.pin_array(batch.event_means())
    out = ctx.alloc_output(batch, pinned=True)
    my_events = batch.events_aligned()
    cores = os.cpu_count() or 1

    # ---- resident (device-timed) ------------------------------------------------------------------------------
    ctx.upload(pinned, means=pinned_means)
    for _ in range(a.warmup):
        ctx.run()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.perf_counter()
    dev_ms = fill_ms = kmer_ms = 0.0
    launches = 0
    for _ in range(a.steps):
        t = ctx.run()
        dev_ms += t["kernel_ms"]; fill_ms += t["fill_ms"]; kmer_ms += t["kmer_ms"]
        launches += t["kernel_launches"]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    n_wide = t["n_wide"]
    res = ctx.download(pinned)
    n_pairs_local = res.n_pairs.copy()
    sched_model = ctx.scheduler_model()

    # ---- the stages either side of the alignment, device-timed on the same resident batch ----
    # (outside the metric's timed region: reported beside it, not in it)
    stages = None
    if rank == 0 and not a.no_extras:
        mom_ms = scl_ms = 0.0
        reps = 3
        for i in range(reps + 1):
            ctx.upload(pinned, with_scalings=False, means=pinned_means)
            est, t1 = ctx.estimate_scalings(batch.n_reads)
            ctx.run()
            t3 = ctx.scaling_stage()
            if i > 0:   # first round is the warm-up
                mom_ms += t1["mom_ms"]; scl_ms += t3["scaling_ms"]
        el = batch.eligible()
        E = batch.n_events.astype(np.int64)[el]; L = batch.read_len.astype(np.int64)[el]; K = batch.n_kmers[el]
        P = n_pairs_local.astype(np.int64)[el]
        # algorithmic bytes, with the event means as the device holds them now (4 B per event): N2 reads
        # every mean twice, the sequence and one model entry per k-mer; N1 reads the pairs, writes and re-reads the
        # k-mer -> event map, and gathers one model entry + one mean per k-mer row in each of its two passes
        b_n2 = int((2 * 4 * E + (L + 1) + 12 * K + 16).sum())
        b_n1 = int((8 * P + 3 * 8 * K + 2 * (L + 1) + 2 * (12 + 4) * K + 48).sum())
        same = bool(np.array_equal(est["shift"], batch.scalings["shift"]) and np.array_equal(est["scale"], batch.scalings["scale"]))
        stages = {"estimate_scalings_mom": {"ms": mom_ms / reps, "algorithmic_bytes": b_n2,
                                            "achieved_gbs": b_n2 / (mom_ms / reps * 1e-3) / 1e9,
                                            "matches_host_scalings_bit_exact": same},
                  "scaling_single": {"ms": scl_ms / reps, "algorithmic_bytes": b_n1,
                                     "achieved_gbs": b_n1 / (scl_ms / reps * 1e-3) / 1e9},
                  "note": "abea_mom_kernel / abea_scaling_kernel on rank 0's batch, CUDA events, 3 runs after 1 warm-up; "
                          "latency-bound by the ordered double sums of the longest read, not by HBM"}
        # N3: event detection on synthetic raw signals with the same event counts (there are no raw signals behind
        # the synthetic event tables, so this stage runs on its own input: ~5 samples per event)
        sg = synth.make_signals(batch.n_reads, batch.meta["mean_events"], batch.meta["sigma"], seed=a.seed)
        cal = (sg["offset"], sg["range"], sg["digitisation"])
        ev_ms = 0.0
        for i in range(reps + 1):
            _ev, _ptr, nev, t5 = ctx.getevents(sg["raw"], sg["raw_ptr"], sg["n_samples"], cal, download=False)
            if i > 0:
                ev_ms += t5["events_ms"]
        ns = int(sg["n_samples"].astype(np.int64).sum())
        stages["getevents"] = {"ms": ev_ms / reps, "samples": ns, "events_detected": int(nev.sum()),
                               "samples_per_s": ns / (ev_ms / reps * 1e-3), "algorithmic_bytes": 57 * ns,
                               "achieved_gbs": 57 * ns / (ev_ms / reps * 1e-3) / 1e9}
        del sg
        ctx.upload(pinned, means=pinned_means)

    # ---- end to end through the C ABI with host buffers (+ NCCL result exchange for N > 1) ------------------------
    exch = None
    if world > 1:
        exch = ResultExchange(rank, world, batch.pair_capacity(), torch.device("cuda", local_rank))
    gathered = None

    def e2e_step():
        nonlocal gathered
        r = ctx.align_batch(pinned, out, means=pinned_means)
        g_ms = 0.0
        if world > 1:
            g0 = time.perf_counter()
            gathered = exch.gather(ctx)   # dense pair lists of every rank -> rank 0's HBM over NCCL/NVLink
            torch.cuda.synchronize()
            g_ms = (time.perf_counter() - g0) * 1e3
        return r, g_ms

    for _ in range(max(1, a.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    gather_ms = 0.0
    h2d = d2h = 0
    e2e_launches = 0
    for _ in range(a.steps):
        r, g = e2e_step()
        gather_ms += g
        h2d, d2h = r.timing["h2d_bytes"], r.timing["d2h_bytes"]
        e2e_launches += r.timing["kernel_launches"] + (2 if world > 1 else 0)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None   # sampled across both timed regions (and the stage timings between them)
    e2e_parts = {kk: r.timing[kk] for kk in ("pack_ms", "h2d_ms", "load_ms", "kernel_ms", "d2h_ms", "unpack_ms")}
    e2e_parts["streamed"] = r.timing["streamed"]
    e2e_parts["n_wide"] = r.timing["n_wide"]
    # what the last timed e2e step left in the caller's pinned buffers must be what the resident runs produced
    e2e_same_as_resident = bool(np.array_equal(out[2], n_pairs_local) and all(
        np.array_equal(out[0][int(out[1][i]):int(out[1][i]) + int(out[2][i])], res.read_pairs(i)) for i in range(batch.n_reads)))

    # ---- reduce over ranks: MAX of times, SUM of units -------------------------------------------------------------
    def reduce(x, op):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=op)
        return float(tt.item())

    MAX = dist.ReduceOp.MAX if world > 1 else None
    SUM = dist.ReduceOp.SUM if world > 1 else None
    MIN = dist.ReduceOp.MIN if world > 1 else None
    dev_ms_max = reduce(dev_ms, MAX)
    e2e_ms_max = reduce(e2e_ms, MAX)
    wall_ms_max = reduce(wall_ms, MAX)
    total_events = reduce(float(my_events), SUM)
    total_reads = reduce(float(batch.n_reads), SUM)
    e2e_ok_all = reduce(1.0 if e2e_same_as_resident else 0.0, MIN) > 0.5
    load_ms_max = reduce(float(r.timing["load_ms"]), MAX)
    alg_bytes = float(batch.algorithmic_bytes(n_pairs_local))

    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        peak, peak_src = peaks()
        value = total_events * a.steps / (dev_ms_max * 1e-3)
        e2e_val = total_events * a.steps / (e2e_ms_max * 1e-3)
        kern_s = (dev_ms / a.steps) * 1e-3      # rank 0's kernels, per step
        achieved = alg_bytes / kern_s / 1e9
        cap = committed_capture(a.config) if (a.reads_per_gpu is None and world == 1) else None
        cfg = config_dict(a, world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": cfg,
            "device": {"name": dev_name, "sm_count": sm_count},
            "events_per_step": int(total_events), "bands_per_step_rank0": int(batch.n_bands[batch.eligible()].sum()),
            "reads_wide_rank0": int(n_wide), "scheduler_model_cycles": sched_model,
            "wall_ms_per_step": wall_ms_max / a.steps,
            "kernels_ms_per_step_rank0": {"kmer_params": kmer_ms / a.steps, "fill_and_traceback": fill_ms / a.steps},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms_max / a.steps, "nccl_gather_ms_per_step": gather_ms / a.steps,
                    "load_ms_max_over_ranks": load_ms_max,
                    "call": "abea_align_batch (C ABI) with pinned host buffers: event means 4 B/event in; pair lists out "
                            + ("as path codes (first pair + 2 bits per step) that the library's host threads expand into the "
                               "caller's buffer while the kernels run" if (r.timing["streamed"] & 4) else
                               "whole, written by the traceback into the caller's mapped buffer (too few host CPUs per "
                               "rank for the code expansion)") + "; d2h bytes are what crossed PCIe",
                    "host_threads": ctx.host_threads(),
                    "last_step_parts_ms_rank0": e2e_parts,
                    "last_step_output_equals_resident_result_all_ranks": e2e_ok_all},
            "gpu_launches": int(launches + e2e_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (cap["dram_bytes_read"] + cap["dram_bytes_write"]) if cap else None,
                         "traffic_source": (cap.get("source") if cap else None),
                         "peak_source": peak_src,
                         "kernel": "abea_fill_kernel<true> = band fill + fused traceback (dominant; bytes and time cover the whole step incl. abea_prepare_kernel and the concurrent wide kernel)",
                         "kernel_share": fill_ms / dev_ms if dev_ms else None,
                         "algorithmic_bytes_per_step_rank0": int(alg_bytes),
                         "bytes_per_event": alg_bytes / max(1.0, float(my_events)),
                         "binding_resource": ({kk: cap[kk] for kk in ("issue_active_per_cycle_active", "warp_instructions_per_band",
                                                                      "pipes_pct_of_peak_active", "source") if kk in cap} if cap else None)},
            "clocks": clocks,
            "stages": stages,
        }
        if world > 1 and gathered is not None:
            # one gathered shard (the last rank's) against the CPU reference on reads regenerated here
            import oracle_lib as ol   # noqa: F401  (run_cpu imports it too)
            rr = world - 1
            br = synth.make_config_shard(a.config, rr, world, seed=a.seed, reads_per_gpu=a.reads_per_gpu)
            idx = parity_sample(br, 4, 12, a.seed + 2)
            want, kind = run_cpu(br.subset(idx), model, cores)
            cnt, dense = gathered[rr]
            cnt = cnt.cpu().numpy()
            off = np.zeros(len(cnt) + 1, dtype=np.int64)
            np.cumsum(cnt.astype(np.int64), out=off[1:])
            dense = dense.cpu().numpy().reshape(-1).view(res.pairs.dtype)
            line["parity_gathered_shard"] = {"rank": rr, "reads_checked": int(len(idx)), "against": kind,
                                             "ok": bool(len(cnt) == br.n_reads and same_pairs(dense, off, cnt, want, idx))}
        if world == 1 and not a.no_cpu_baseline:
            target = a.cpu_sample_events or min(float(my_events), cores * 0.3e6 * 6.0)
            sample, n_s = cpu_sample(batch, target)
            rcpu, kind = run_cpu(sample, model, cores)
            ev = sample.events_aligned()
            line["cpu_baseline"] = {"value": ev / rcpu.seconds, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"first {n_s} reads of the same batch ({ev} events), "
                                              f"{rcpu.seconds:.2f} s wall, all host threads"}
            # the OUTPUT BUFFER OF THE LAST TIMED e2e STEP against the CPU arm's pairs (the sample is a prefix of the batch)
            line["parity_on_cpu_sample"] = bool(same_pairs(out[0], out[1], out[2], rcpu, range(n_s)))
            if not a.no_extras:
                thr = a.dropin_threads or min(16, cores)
                dr = e2e_dropin(batch, model, k, local_rank, thr, max(3, min(a.steps, 10)), 2, rcpu, n_s)
                line["e2e_dropin"] = dr[0] if isinstance(dr, tuple) else dr
                if isinstance(dr, tuple):   # every read's list, against what the C-ABI call left in the pinned buffer
                    line["e2e_dropin"]["equals_c_abi_result"] = bool(np.array_equal(dr[2], out[2]) and all(
                        np.array_equal(dr[1][int(out[1][i]):int(out[1][i]) + int(out[2][i])],
                                       out[0][int(out[1][i]):int(out[1][i]) + int(out[2][i])]) for i in range(batch.n_reads)))
                line["gpu_reference"] = gpu_reference(batch, model, k, local_rank, 3, 1, res.pairs, res.pair_ptr, res.n_pairs)
        if world == 1 and not a.no_extras:
            ctx.close()
            ctx = None
            line["configs"] = other_configs(lambda: AbeaContext(local_rank), a, peak, cores)
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:   # last thing written, on a line of its own (NCCL and torch print to the same pipe under torchrun)
        sys.stderr.flush()
        sys.stdout.write("\n" + json.dumps(line) + "\n")
        sys.stdout.flush()


def reference_arm(a, rank, world):
    """The reference's own CPU implementation of the path (align_db's CPU branch) on this box's host cores."""
    if rank != 0:
        return
    from f5c_b200 import models, synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    world = max(1, a.gpus)
    batch = synth.make_config_shard(a.config, 0, world, seed=a.seed, reads_per_gpu=a.reads_per_gpu)
    k, model = models.load_model(batch.meta["model"])
    model = ol.full_model(model)
    cores = os.cpu_count() or 1
    # each step: a bounded sample sized for ~6 s on all cores at ~0.3 M events/s/thread
    target = a.cpu_sample_events or min(float(batch.events_aligned()), cores * 0.3e6 * 6.0)
    sample, n_s = cpu_sample(batch, target)
    ev = sample.events_aligned()
    for _ in range(a.warmup):
        run_cpu(sample, model, cores)
    t = 0.0
    kind = "port"
    for _ in range(a.steps):
        r, kind = run_cpu(sample, model, cores)
        t += r.seconds
    value = ev * a.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": t / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 emission + f64 transition sums", "data": "synthetic",
            "config": config_dict(a, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"first {n_s} of rank 0's {batch.n_reads} reads ({ev} events) per step, all {cores} host "
                                       "threads, reference align() behind a dynamic work queue (align_db CPU branch)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
