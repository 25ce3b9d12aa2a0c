"""Multi-GPU host logic: reads are independent, so a batch is partitioned read-wise (one process per GPU, no
collective on the data path) and only the RESULTS are gathered to rank 0 (north_star: "NCCL over NVLink only for the
final result gather"). The reference has no counterpart (multi-GPU = run several processes by hand, docs/f5c.1:271).

The same functions run over NCCL (CUDA tensors, bench.py) and over gloo (CPU tensors, tests/test_sharding_gloo.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .batch import PAIR_DTYPE


def compact_pairs(pairs: np.ndarray, pair_ptr: np.ndarray, n_pairs: np.ndarray) -> np.ndarray:
    """Concatenate the per-read pair lists (capacity layout -> dense)."""
    total = int(n_pairs.astype(np.int64).sum())
    out = np.empty(total, dtype=PAIR_DTYPE)
    pos = 0
    for i in range(n_pairs.shape[0]):
        n = int(n_pairs[i])
        if n:
            p = int(pair_ptr[i])
            out[pos:pos + n] = pairs[p:p + n]
            pos += n
    return out


def gather_results(n_pairs: np.ndarray, dense_pairs: np.ndarray, rank: int, world: int, device: str = "cpu"):
    """Gather every rank's (n_pairs, dense pair list) to rank 0.

    Returns on rank 0 a list of (n_pairs, pairs) per rank (rank order); None elsewhere. Two collectives: an
    all_gather of the shard sizes (so every rank can pad to a common length) and a gather of the padded payloads.
    """
    dev = torch.device(device)
    counts = torch.from_numpy(np.ascontiguousarray(n_pairs.astype(np.int32))).to(dev)
    flat = torch.from_numpy(np.ascontiguousarray(dense_pairs).view(np.int32).reshape(-1, 2)).to(dev)
    sizes = torch.tensor([counts.numel(), flat.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [s.cpu().numpy() for s in all_sizes]
    max_reads = int(max(s[0] for s in all_sizes))
    max_pairs = int(max(s[1] for s in all_sizes))
    cpad = torch.zeros(max_reads, dtype=torch.int32, device=dev)
    cpad[:counts.numel()] = counts
    ppad = torch.zeros((max(max_pairs, 1), 2), dtype=torch.int32, device=dev)
    ppad[:flat.shape[0]] = flat
    cl = [torch.empty_like(cpad) for _ in range(world)] if rank == 0 else None
    pl = [torch.empty_like(ppad) for _ in range(world)] if rank == 0 else None
    dist.gather(cpad, cl, dst=0)
    dist.gather(ppad, pl, dst=0)
    if rank != 0:
        return None
    out = []
    for r in range(world):
        nr, npz = int(all_sizes[r][0]), int(all_sizes[r][1])
        c = cl[r][:nr].cpu().numpy()
        p = pl[r][:npz].cpu().numpy().reshape(-1).view(PAIR_DTYPE)
        out.append((c, p))
    return out


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can view library-owned device memory without a copy."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def device_views(ctx):
    """torch views (no copy) of the context's device-resident results: pairs int32 [cap, 2], n_pairs int32 [n]."""
    dp, dn, cap, n = ctx.device_results()
    dev = torch.device("cuda", ctx.device)
    pairs = torch.as_tensor(_DevArray(dp, (max(cap, 1), 2), "<i4"), device=dev)[:cap]
    counts = torch.as_tensor(_DevArray(dn, (max(n, 1),), "<i4"), device=dev)[:n]
    return pairs, counts


class ResultExchange:
    """The result exchange of the multi-GPU path (north_star: "NCCL over NVLink only for the final result gather").

    Set up once per shard shape: every rank's read count and pair capacity are exchanged, rank 0 allocates one
    receive buffer per peer and every rank one send buffer — nothing is allocated per step. A step then (1) packs the
    rank's pair lists back to back on the device (abea_compact_results: 8 bytes per PAIR cross NVLink, not 8 bytes per
    capacity slot), (2) gathers the exact pair totals and the per-read counts (padded to the largest shard) to rank 0,
    and (3) moves every shard with one point-to-point transfer of exactly its size. On CUDA the tensors are device
    memory and the backend is NCCL; with CPU tensors the same code runs over gloo against the CPU emulation build of
    the library (tests/test_sharding_gloo.py)."""

    def __init__(self, rank: int, world: int, n_reads: int, pair_capacity: int, device):
        self.rank, self.world = rank, world
        self.dev = torch.device(device)
        self.n_reads, self.cap = int(n_reads), int(pair_capacity)
        meta = torch.tensor([self.n_reads, self.cap], dtype=torch.int64, device=self.dev)
        allm = [torch.zeros(2, dtype=torch.int64, device=self.dev) for _ in range(world)]
        dist.all_gather(allm, meta)
        self.meta = [tuple(int(v) for v in m.cpu().tolist()) for m in allm]
        self.max_reads = max(m[0] for m in self.meta)
        self.send = torch.empty((max(self.cap, 1), 2), dtype=torch.int32, device=self.dev)
        self.counts = torch.zeros(self.max_reads + 2, dtype=torch.int32, device=self.dev)   # [total lo, total hi, counts...]
        self.recv = self.recv_counts = None
        if rank == 0:
            self.recv = [self.send if r == 0 else torch.empty((max(self.meta[r][1], 1), 2), dtype=torch.int32, device=self.dev)
                         for r in range(world)]
            self.recv_counts = [torch.zeros(self.max_reads + 2, dtype=torch.int32, device=self.dev) for _ in range(world)]

    def _counts_view(self, ctx):
        dp, dn, cap, n = ctx.device_results()
        if self.dev.type == "cuda":
            return torch.as_tensor(_DevArray(dn, (max(n, 1),), "<i4"), device=self.dev)[:n]
        import ctypes
        return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_int32 * max(n, 1)).from_address(dn))[:n])

    def gather(self, ctx):
        """Returns on rank 0 a list of (n_pairs int32 [n_reads_r], dense pairs int32 [total_r, 2]) per rank; None elsewhere."""
        total = ctx.compact_results(self.send.data_ptr(), self.send.shape[0])
        self.counts[0] = total & 0x7fffffff
        self.counts[1] = total >> 31
        self.counts[2:2 + self.n_reads] = self._counts_view(ctx)
        dist.gather(self.counts, self.recv_counts, dst=0)
        if self.rank != 0:
            if total:   # one grouped point-to-point operation per shard (ncclGroupStart / ncclSend / ncclGroupEnd)
                for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, self.send[:total], 0)]):
                    q.wait()
            return None
        heads = torch.stack([c[:2] for c in self.recv_counts]).cpu().tolist()      # the one host sync of the exchange
        sizes = [int(h[0]) | (int(h[1]) << 31) for h in heads]
        ops = [dist.P2POp(dist.irecv, self.recv[r][:sizes[r]], r) for r in range(1, self.world) if sizes[r]]
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        return [(self.recv_counts[r][2:2 + self.meta[r][0]], self.recv[r][:sizes[r]]) for r in range(self.world)]


def gather_device_results(ctx, rank: int, world: int):
    """Round-1 form of the exchange, kept for comparison: gathers the CAPACITY layout padded to the largest shard and
    allocates its buffers per call. ResultExchange replaces it."""
    pairs, counts = device_views(ctx)
    dev = pairs.device
    sizes = torch.tensor([counts.numel(), pairs.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [s.cpu().numpy() for s in all_sizes]
    max_reads = int(max(s[0] for s in all_sizes))
    max_pairs = int(max(s[1] for s in all_sizes))
    cpad = torch.zeros(max_reads, dtype=torch.int32, device=dev)
    cpad[:counts.numel()] = counts
    ppad = torch.empty((max(max_pairs, 1), 2), dtype=torch.int32, device=dev)
    ppad[:pairs.shape[0]] = pairs
    cl = [torch.empty_like(cpad) for _ in range(world)] if rank == 0 else None
    pl = [torch.empty_like(ppad) for _ in range(world)] if rank == 0 else None
    dist.gather(cpad, cl, dst=0)
    dist.gather(ppad, pl, dst=0)
    if rank != 0:
        return None
    return [(cl[r][:int(all_sizes[r][0])], pl[r][:int(all_sizes[r][1])]) for r in range(world)]
