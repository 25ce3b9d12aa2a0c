#!/usr/bin/env python3
"""bench.py — events aligned/sec of the ABEA hot path on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps K --warmup W                     # our CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus N --steps K --warmup W    # the reference's own CPU align() on host cores

A "step" is one pass of the hot path over one synthetic batch (BASELINE.json configs[1], "cfg2": R9.4.1, 4096 reads
per GPU, mean 4k events/read; other configs via --config). `value` is device-timed (CUDA events inside the library,
on the stream the kernels are launched on) with the batch already resident in HBM; `e2e` is the same metric through
the C-ABI call with HOST (pinned) buffers in and out — everything inside the timed region: the events are pulled over
PCIe by abea_load_kernel while the fill runs and the pair lists are written straight into the caller's pinned buffer
by the traceback (timing["streamed"] == 3; ABEA_STREAM=0 stages through the copy engine instead) — plus, for N > 1,
the NCCL gather of all ranks' device-resident results to rank 0. Reads are partitioned read-wise across ranks (weak
scaling: 4096 reads per GPU); there is no collective on the data path, only the final result gather.
"""
from __future__ import annotations

import argparse
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--reads-per-gpu", type=int, default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-sample-events", type=float, default=0, help="events in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
def measured_traffic():
    """DRAM bytes of the dominant kernel from the committed ncu --set full capture (profiles/), per launch."""
    p = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    return None


def measured_issue():
    """What actually bounds the dominant kernel, from the same committed capture: issue slots, not HBM."""
    p = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if "issue_active_per_cycle_active" in d:
            return {"issue_slots_used_while_active": d["issue_active_per_cycle_active"],
                    "sub_partitions_active_fraction": d["smsp_cycles_active_avg"] / d["sm_cycles_elapsed_max"],
                    "pipes_pct_of_peak_active": d.get("pipes_pct_of_peak_active"),
                    "source": "profiles/fill_narrow_final_r01.txt (ncu --set full of the same kernel and workload)"}
    return None


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
    from f5c_b200.dist import gather_device_results

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
    model = ctx.set_model(model, k)
    sm_count, dev_name = ctx.device_info()
    pinned = ctx.pin_batch(batch)
    out = ctx.alloc_output(batch, pinned=True)
    my_events = batch.events_aligned()

    # ---- resident (device-timed) ------------------------------------------------------------------------------
    ctx.upload(pinned)
    for _ in range(a.warmup):
        ctx.run()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    fill_ms = trace_ms = kmer_ms = 0.0
    launches = 0
    for _ in range(a.steps):
        t = ctx.run()
        dev_ms += t["kernel_ms"]
        fill_ms += t["fill_ms"]; trace_ms += t["trace_ms"]; kmer_ms += t["kmer_ms"]
        launches += t["kernel_launches"]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    res = ctx.download(pinned, out)
    n_pairs_local = res.n_pairs.copy()

    # ---- the stages either side of the alignment, device-timed on the same resident batch ------
    # (outside the metric's timed region: reported beside it, not in it)
    stages = None
    if rank == 0:
        mom_ms = scl_ms = 0.0
        reps = 3
        for i in range(reps + 1):
            ctx.upload(pinned, with_scalings=False)
            est, t1 = ctx.estimate_scalings(batch.n_reads)
            ctx.run()
            t3 = ctx.scaling_stage()
            if i > 0:   # first round is the warm-up
                mom_ms += t1["mom_ms"]; scl_ms += t3["scaling_ms"]
        el = batch.eligible()
        E = batch.n_events.astype(np.int64)[el]; L = batch.read_len.astype(np.int64)[el]; K = batch.n_kmers[el]
        P = n_pairs_local.astype(np.int64)[el]
        # algorithmic bytes: N2 reads every event twice (24-B AoS as delivered), the sequence and one
        # model entry per k-mer; N1 reads the pairs, writes and re-reads the k-mer -> event map, and gathers one model
        # entry + one event per k-mer row in each of its two passes
        b_n2 = int((2 * 24 * E + (L + 1) + 12 * K + 16).sum())
        b_n1 = int((8 * P + 3 * 8 * K + 2 * (L + 1) + 2 * (12 + 24) * K + 48).sum())
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
            _ev, _ptr, nev, t5 = ctx.getevents(sg["raw"], sg["raw_ptr"], sg["n_samples"], cal)
            if i > 0:
                ev_ms += t5["events_ms"]
        ns = int(sg["n_samples"].astype(np.int64).sum())
        stages["getevents"] = {"ms": ev_ms / reps, "samples": ns, "events_detected": int(nev.sum()),
                               "samples_per_s": ns / (ev_ms / reps * 1e-3), "algorithmic_bytes": 57 * ns,
                               "achieved_gbs": 57 * ns / (ev_ms / reps * 1e-3) / 1e9}
        del sg, _ev
        ctx.upload(pinned)

    # ---- end to end through the C ABI with host buffers (+ NCCL result gather for N > 1) --------------------------
    def e2e_step():
        r = ctx.align_batch(pinned, out)
        g_ms = 0.0
        if world > 1:
            g0 = time.perf_counter()
            gather_device_results(ctx, rank, world)   # device-resident results -> rank 0's HBM over NCCL/NVLink
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
        e2e_launches += r.timing["kernel_launches"]
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None   # sampled across both timed regions (and the stage timings between them)
    e2e_parts = {kk: r.timing[kk] for kk in ("pack_ms", "h2d_ms", "load_ms", "kernel_ms", "d2h_ms", "unpack_ms")}
    e2e_parts["streamed"] = r.timing["streamed"]

    # ---- reduce over ranks: MAX of times, SUM of units -------------------------------------------------------------
    def reduce(x, op):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=op)
        return float(tt.item())

    MAX = dist.ReduceOp.MAX if world > 1 else None
    SUM = dist.ReduceOp.SUM if world > 1 else None
    dev_ms_max = reduce(dev_ms, MAX)
    e2e_ms_max = reduce(e2e_ms, MAX)
    wall_ms_max = reduce(wall_ms, MAX)
    total_events = reduce(float(my_events), SUM)
    total_reads = reduce(float(batch.n_reads), SUM)
    alg_bytes = float(batch.algorithmic_bytes(n_pairs_local))

    if rank == 0:
        peak, peak_src = peaks()
        value = total_events * a.steps / (dev_ms_max * 1e-3)
        e2e_val = total_events * a.steps / (e2e_ms_max * 1e-3)
        kern_s = (dev_ms / a.steps) * 1e-3      # rank 0's three kernels, per step
        achieved = alg_bytes / kern_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 emission + f64 transition sums (bit-exact vs CPU)", "data": "synthetic",
            "config": {"workload": f"{a.config}: synthetic {batch.meta['model']} DNA/RNA batch, "
                                   f"{int(total_reads)} reads ({batch.n_reads}/GPU), mean {batch.meta['mean_events']} "
                                   f"events/read log-normal sigma {batch.meta['sigma']}, bandwidth 100",
                       "global_reads": int(total_reads), "events_per_step": int(total_events),
                       "bands_per_step_rank0": int(batch.n_bands[batch.eligible()].sum()),
                       "parallelism": f"read-sharded x{world} (LPT by band count), NCCL result gather only",
                       "l2": "inputs larger than L2 (events 24 B/event + trace 32 B/band per step >> 126 MB); no flush",
                       "seed": a.seed, "device": dev_name, "sm_count": sm_count},
            "wall_ms_per_step": wall_ms_max / a.steps,
            "kernels_ms_per_step_rank0": {"kmer_params": kmer_ms / a.steps, "fill": fill_ms / a.steps,
                                          "traceback": trace_ms / a.steps},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms_max / a.steps, "nccl_gather_ms_per_step": gather_ms / a.steps,
                    "last_step_parts_ms_rank0": e2e_parts},
            "gpu_launches": int(launches + e2e_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (measured_traffic() if a.config == "cfg2" and a.reads_per_gpu is None else None),
                         "peak_source": peak_src,
                         "kernel": "abea_fill_kernel<true> = band fill + fused traceback (dominant; bytes and time cover the whole step incl. abea_prepare_kernel)",
                         "kernel_share": fill_ms / dev_ms if dev_ms else None,
                         "algorithmic_bytes_per_step_rank0": int(alg_bytes),
                         "bytes_per_event": alg_bytes / max(1.0, float(my_events)),
                         "binding_resource": (measured_issue() if a.config == "cfg2" and a.reads_per_gpu is None else None)},
            "clocks": clocks,
            "stages": stages,
        }
        if world == 1 and not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            target = a.cpu_sample_events or min(float(my_events), cores * 0.3e6 * 6.0)
            sample, n_s = cpu_sample(batch, target)
            rcpu, kind = run_cpu(sample, model, cores)
            ev = sample.events_aligned()
            line["cpu_baseline"] = {"value": ev / rcpu.seconds, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"first {n_s} reads of the same batch ({ev} events), "
                                              f"{rcpu.seconds:.2f} s wall, all host threads"}
            # the sample doubles as a live parity check of the benchmarked path
            sub = ctx.align_batch(sample)
            ok = bool(np.array_equal(sub.n_pairs, rcpu.n_pairs)) and all(
                np.array_equal(sub.read_pairs(i), rcpu.read_pairs(i)) for i in range(sample.n_reads))
            line["parity_on_cpu_sample"] = ok
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def reference_arm(a, rank, world):
    """The reference's own CPU implementation of the path (align_db's CPU branch) on this box's host cores."""
    if rank != 0:
        return
    from f5c_b200 import models, synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    batch = synth.make_config_shard(a.config, 0, max(1, a.gpus), seed=a.seed, reads_per_gpu=a.reads_per_gpu)
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
            "config": {"workload": f"{a.config}: bounded sample of the same synthetic batch "
                                   f"(first {n_s} of {batch.n_reads} reads, {ev} events per step)", "seed": a.seed},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"first {n_s} reads ({ev} events) per step, all {cores} host threads, "
                                       "reference align() behind a dynamic work queue (align_db CPU branch)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
