"""Seeded synthetic read batches for the BASELINE.json configurations (recipe: SURVEY.md §8d).

Per read: target events E* ~ LogNormal(ln(mean) - sigma^2/2, sigma); L = floor(E*/epk) + k (min 200);
sequence iid uniform over ACGT; per k-mer: with p=0.03 no event (skip) else Geometric so the mean number of
events per k-mer is epk; event mean = scale_r*level_mean[rank] + shift_r + level_stdv[rank]*N(0,1) with
shift_r ~ N(0,10), scale_r ~ N(1,0.05); start cumulative, length 3..22, stdv 1.0 (unused by ABEA).
Scalings are a method-of-moments estimate (same formula as reference src/align.c:58-106, numpy float64).
"""
from __future__ import annotations

import numpy as np

from .batch import EVENT_DTYPE, SCALINGS_DTYPE, ReadBatch
from .models import load_model

# BASELINE.json configs -> generator parameters
CONFIGS = {
    "cfg2": dict(model="r9", n_reads=4096, mean_events=4000, sigma=0.5, epk=1.8),
    "cfg3": dict(model="r10", n_reads=4096, mean_events=8000, sigma=1.0, epk=1.9),
    "cfg4": dict(model="rna004", n_reads=2048, mean_events=20000, sigma=0.5, epk=2.5),
    # BASELINE configs[4] / the north_star target: R10.4.1, mean 4k events/read, 4096 reads PER GPU, read-sharded — the
    # 32768-read batch of configs[4] is this config at world == 8; at world == N it is N x 4096 reads (weak scaling)
    "cfg5": dict(model="r10", n_reads=4096, mean_events=4000, sigma=0.5, epk=1.9),
}
GLOBAL_READS_AT_8 = {"cfg5": 32768}


def kmer_ranks_flat(bases: np.ndarray, k: int) -> np.ndarray:
    """rank[i] of the k-mer starting at flat position i (first base most significant; reference
    src/align.c:36-47). Positions within k-1 of the end are garbage and must be masked by the caller."""
    n = bases.shape[0]
    r = np.zeros(n, dtype=np.int64)
    for j in range(k):
        shifted = np.zeros(n, dtype=np.int64)
        shifted[:n - j] = bases[j:]
        r = (r << 2) | shifted
    return r


def mom_scalings(ev_mean: np.ndarray, levels: np.ndarray):
    """estimate_scalings_using_mom for one read (reference src/align.c:58-106), float64 accumulators."""
    ev = ev_mean.astype(np.float64)
    lv = levels.astype(np.float64)
    shift = ev.sum() / ev.size - lv.sum() / lv.size
    scale = (((ev - shift) ** 2).sum() / ev.size) / ((lv * lv).sum() / lv.size)
    return np.float32(scale), np.float32(shift)


def draw_lengths(n_reads: int, mean_events: float, sigma: float, rng) -> np.ndarray:
    """Target events per read: LogNormal(ln(mean) - sigma^2/2, sigma)."""
    return rng.lognormal(np.log(mean_events) - 0.5 * sigma * sigma, sigma, n_reads)


def lpt_shards(weights: np.ndarray, n_shards: int):
    """Longest-processing-time-first partition: reads sorted by weight descending, each to the lightest shard
    (balance the sum of band counts, not the read count). Returns a list of index arrays."""
    order = np.argsort(-np.asarray(weights, dtype=np.float64), kind="stable")
    load = np.zeros(n_shards)
    bins = [[] for _ in range(n_shards)]
    for i in order:
        j = int(np.argmin(load))
        bins[j].append(int(i))
        load[j] += weights[i]
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def make_batch(model: str = "r9", n_reads: int = 64, mean_events: float = 4000.0, sigma: float = 0.5,
               epk: float = 1.8, seed: int = 42, model_table=None, min_len: int = 200,
               e_star: np.ndarray | None = None, p_skip: float = 0.03) -> ReadBatch:
    """Generate a ragged batch. model_table=(k, MODEL_DTYPE array) overrides the named built-in table;
    e_star (target event counts per read) overrides the log-normal draw (used for read-wise sharding); p_skip is the
    share of k-mers without any event (0.03 in the validated recipe; tests raise it to stress skip runs)."""
    k, mt = model_table if model_table is not None else load_model(model)
    rng = np.random.default_rng(seed)
    if e_star is None:
        e_star = draw_lengths(n_reads, mean_events, sigma, rng)
    else:
        e_star = np.asarray(e_star, dtype=np.float64)
        n_reads = int(e_star.shape[0])
    L = np.maximum((e_star / epk).astype(np.int64) + k, min_len)
    K = L - k + 1

    # sequences (flat, NUL after each read)
    seq_ptr = np.zeros(n_reads, dtype=np.int64)
    np.cumsum(L[:-1] + 1, out=seq_ptr[1:])
    total_seq = int((L + 1).sum())
    base_idx = rng.integers(0, 4, total_seq, dtype=np.int64)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[base_idx].copy()
    seq[seq_ptr + L] = 0
    ranks_flat = kmer_ranks_flat(base_idx, k)

    # flat k-mer index -> (read, position)
    kptr = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(K, out=kptr[1:])
    total_k = int(kptr[-1])
    read_of_kmer = np.repeat(np.arange(n_reads, dtype=np.int64), K)
    pos_in_read = np.arange(total_k, dtype=np.int64) - kptr[read_of_kmer]
    rank = ranks_flat[seq_ptr[read_of_kmer] + pos_in_read]

    # events per k-mer: 3% skips, otherwise Geometric (>=1) with overall mean epk
    cnt = rng.geometric(min(1.0, 0.97 / epk), total_k).astype(np.int64)
    cnt[rng.random(total_k) < p_skip] = 0
    # every read needs at least one event on its first and last k-mer so it can span
    cnt[kptr[:-1]] = np.maximum(cnt[kptr[:-1]], 1)
    cnt[kptr[1:] - 1] = np.maximum(cnt[kptr[1:] - 1], 1)

    ev_rank = np.repeat(rank, cnt)
    ev_read = np.repeat(read_of_kmer, cnt)
    n_ev_total = int(ev_rank.shape[0])
    n_events = np.bincount(ev_read, minlength=n_reads).astype(np.int32)
    event_ptr = np.zeros(n_reads, dtype=np.int64)
    np.cumsum(n_events[:-1].astype(np.int64), out=event_ptr[1:])

    shift_r = rng.normal(0.0, 10.0, n_reads)
    scale_r = rng.normal(1.0, 0.05, n_reads)
    lm = mt["level_mean"].astype(np.float64)
    ls = mt["level_stdv"].astype(np.float64)
    mean = scale_r[ev_read] * lm[ev_rank] + shift_r[ev_read] + ls[ev_rank] * rng.standard_normal(n_ev_total)

    events = np.zeros(n_ev_total, dtype=EVENT_DTYPE)
    events["mean"] = mean.astype(np.float32)
    length = rng.integers(3, 23, n_ev_total)
    events["length"] = length.astype(np.float32)
    cs = np.cumsum(length) - length
    events["start"] = (cs - cs[event_ptr][ev_read]).astype(np.uint64)
    events["stdv"] = 1.0

    scalings = np.zeros(n_reads, dtype=SCALINGS_DTYPE)
    for i in range(n_reads):
        ep, en = int(event_ptr[i]), int(n_events[i])
        s, sh = mom_scalings(events["mean"][ep:ep + en], mt["level_mean"][rank[kptr[i]:kptr[i + 1]]])
        scalings[i]["scale"] = s
        scalings[i]["shift"] = sh

    return ReadBatch(seq, seq_ptr, L.astype(np.int32), events, event_ptr, n_events, scalings,
                     np.ones(n_reads, dtype=np.uint8), k,
                     meta=dict(model=model, n_reads=n_reads, mean_events=mean_events, sigma=sigma, epk=epk,
                               seed=seed))


def make_config(name: str, seed: int = 42, n_reads: int | None = None) -> ReadBatch:
    p = dict(CONFIGS[name])
    if n_reads is not None:
        p["n_reads"] = n_reads
    b = make_batch(seed=seed, **p)
    b.meta["config"] = name
    return b


def make_config_shard(name: str, rank: int, world: int, seed: int = 42, reads_per_gpu: int | None = None) -> ReadBatch:
    """Rank `rank`'s shard of a global batch of world*reads_per_gpu reads, partitioned read-wise (CONFIGS[name]["n_reads"]
    is the per-GPU count: cfg5 at world == 8 is the 32768-read batch of BASELINE configs[4]).

    Every rank draws the same global list of target lengths (one cheap log-normal draw), the list is split
    longest-first by estimated band count E*(1+1/epk), and the rank materialises only its own reads. With world == 1
    this is exactly make_config(name, seed)."""
    p = dict(CONFIGS[name])
    per = reads_per_gpu if reads_per_gpu is not None else p["n_reads"]
    if world == 1:
        return make_config(name, seed=seed, n_reads=per)
    rng = np.random.default_rng(seed)
    e_star = draw_lengths(per * world, p["mean_events"], p["sigma"], rng)
    shards = lpt_shards(e_star * (1.0 + 1.0 / p["epk"]), world)
    p.pop("n_reads")
    b = make_batch(seed=seed * 1000003 + rank + 1, e_star=e_star[shards[rank]], **p)
    b.meta.update(config=name, rank=rank, world=world, global_reads=per * world)
    return b


def make_signals(n_reads: int, mean_events: float, sigma: float, seed: int, samples_per_event: float = 5.0,
                 noise: float = 1.6, min_events: int = 40):
    """Seeded synthetic raw signals for event detection (reference getevents, src/events.c): per read a piecewise
    constant current (levels ~ N(90, 13) pA, dwell 1 + Geometric samples with the given mean) plus Gaussian noise,
    quantised to int16 ADC counts with a MinION-like calibration. Returns a dict with
    raw (float32 ADC counts, flat), raw_ptr, n_samples, offset, range, digitisation, and pa (the same samples in pA,
    converted exactly as event_single does, src/f5c.c:692-696)."""
    rng = np.random.default_rng(seed)
    e_star = np.maximum(draw_lengths(n_reads, mean_events, sigma, rng), min_events).astype(np.int64)
    total_e = int(e_star.sum())
    dwell = rng.geometric(min(1.0, 1.0 / max(1.0, samples_per_event - 1.0)), total_e).astype(np.int64) + 1
    level = rng.normal(90.0, 13.0, total_e)
    read_of_event = np.repeat(np.arange(n_reads, dtype=np.int64), e_star)
    n_samples = np.bincount(read_of_event, weights=dwell, minlength=n_reads).astype(np.int64)
    pa_true = np.repeat(level, dwell) + noise * rng.standard_normal(int(dwell.sum()))
    digitisation = np.full(n_reads, 8192.0, dtype=np.float32)
    rng_pa = rng.uniform(1380.0, 1480.0, n_reads).astype(np.float32)
    offset = np.round(rng.uniform(2.0, 20.0, n_reads)).astype(np.float32)
    read_of_sample = np.repeat(np.arange(n_reads, dtype=np.int64), n_samples)
    adc = np.round(pa_true * (digitisation[read_of_sample] / rng_pa[read_of_sample]) - offset[read_of_sample])
    raw = np.clip(adc, -32768, 32767).astype(np.int16).astype(np.float32)
    raw_ptr = np.zeros(n_reads, dtype=np.int64)
    if n_reads > 1:
        np.cumsum(n_samples[:-1], out=raw_ptr[1:])
    raw_unit = (rng_pa / digitisation).astype(np.float32)
    pa = ((raw + offset[read_of_sample]).astype(np.float32) * raw_unit[read_of_sample]).astype(np.float32)
    return dict(raw=raw, raw_ptr=raw_ptr, n_samples=n_samples.astype(np.int32), offset=offset, range=rng_pa,
                digitisation=digitisation, pa=pa)


def make_signal_batch(model: str, n_reads: int, mean_kmers: float, sigma: float, seed: int, dwell_mean: float = 9.0,
                      noise: float = 1.5):
    """Seeded raw signals WITH the sequences they come from, for the device-resident chain raw signal -> events ->
    scalings -> alignment -> recalibration: per read a random ACGT sequence; every k-mer holds the current at its
    model level (scaled / shifted per read) for 2 + Geometric samples, plus Gaussian sample noise; quantised to int16
    ADC counts like make_signals. Returns (signals dict as make_signals, seq uint8 flat NUL-separated, seq_ptr,
    read_len, kmer_size)."""
    k, mt = load_model(model)
    rng = np.random.default_rng(seed)
    K = np.maximum(draw_lengths(n_reads, mean_kmers, sigma, rng).astype(np.int64), 40)
    L = K + k - 1
    seq_ptr = np.zeros(n_reads, dtype=np.int64)
    if n_reads > 1:
        np.cumsum(L[:-1] + 1, out=seq_ptr[1:])
    seq = np.zeros(int(L.sum()) + n_reads, dtype=np.uint8)
    bases = rng.integers(0, 4, int(L.sum()))
    read_of_base = np.repeat(np.arange(n_reads, dtype=np.int64), L)
    pos = np.arange(int(L.sum()), dtype=np.int64) - np.repeat(np.cumsum(L) - L, L)
    seq[seq_ptr[read_of_base] + pos] = np.frombuffer(b"ACGT", dtype=np.uint8)[bases]
    kptr = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(K, out=kptr[1:])
    read_of_kmer = np.repeat(np.arange(n_reads, dtype=np.int64), K)
    kpos = np.arange(int(K.sum()), dtype=np.int64) - kptr[read_of_kmer]
    bstart = (np.cumsum(L) - L)[read_of_kmer] + kpos
    rank = np.zeros(int(K.sum()), dtype=np.int64)
    for j in range(k):
        rank = rank * 4 + bases[bstart + j]
    shift_r = rng.normal(0.0, 8.0, n_reads)
    scale_r = rng.normal(1.0, 0.04, n_reads)
    level = scale_r[read_of_kmer] * mt["level_mean"].astype(np.float64)[rank] + shift_r[read_of_kmer]
    dwell = rng.geometric(min(1.0, 1.0 / max(1.0, dwell_mean - 2.0)), int(K.sum())).astype(np.int64) + 2
    n_samples = np.bincount(read_of_kmer, weights=dwell, minlength=n_reads).astype(np.int64)
    pa_true = np.repeat(level, dwell) + noise * rng.standard_normal(int(dwell.sum()))
    digitisation = np.full(n_reads, 8192.0, dtype=np.float32)
    rng_pa = rng.uniform(1380.0, 1480.0, n_reads).astype(np.float32)
    offset = np.round(rng.uniform(2.0, 20.0, n_reads)).astype(np.float32)
    read_of_sample = np.repeat(np.arange(n_reads, dtype=np.int64), n_samples)
    adc = np.round(pa_true * (digitisation[read_of_sample] / rng_pa[read_of_sample]) - offset[read_of_sample])
    raw = np.clip(adc, -32768, 32767).astype(np.int16).astype(np.float32)
    raw_ptr = np.zeros(n_reads, dtype=np.int64)
    if n_reads > 1:
        np.cumsum(n_samples[:-1], out=raw_ptr[1:])
    raw_unit = (rng_pa / digitisation).astype(np.float32)
    pa = ((raw + offset[read_of_sample]).astype(np.float32) * raw_unit[read_of_sample]).astype(np.float32)
    sig = dict(raw=raw, raw_ptr=raw_ptr, n_samples=n_samples.astype(np.int32), offset=offset, range=rng_pa,
               digitisation=digitisation, pa=pa)
    return sig, seq, seq_ptr, L.astype(np.int32), k
