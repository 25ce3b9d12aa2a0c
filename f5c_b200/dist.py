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

    What moves between the GPUs is each rank's pair lists as PATH CODES (abea_device_codes: the first pair of a list and
    two bits per step, 8 bytes per 32 pairs) plus the per-read counts: 1/32 of the pairs' bytes, and the size of a
    rank's code buffer is a function of its shard's shape alone, so nothing has to be agreed per step — no compaction,
    no size exchange, no host synchronisation before the transfer. Set up once per shard shape: every rank's read count,
    capacity and capacity prefix sums go to rank 0, which allocates one receive buffer per peer and one dense pair
    buffer per shard — nothing is allocated per step. A step is one grouped point-to-point operation (ncclGroupStart /
    ncclSend | ncclRecv x 2 per peer / ncclGroupEnd) and, on rank 0, one expansion kernel per shard
    (abea_expand_codes) that leaves every shard's pair lists back to back in rank 0's HBM. On CUDA the tensors are
    device memory and the backend is NCCL; with CPU tensors the same code runs over gloo against the CPU emulation build
    of the library (tests/test_sharding_gloo.py)."""

    def __init__(self, rank: int, world: int, pair_capacity, device):
        self.rank, self.world = rank, world
        self.dev = torch.device(device)
        cap = np.ascontiguousarray(pair_capacity, dtype=np.int64)
        self.n_reads = int(cap.shape[0])
        self.cap_ptr = np.zeros(self.n_reads + 1, dtype=np.int64)
        np.cumsum(cap, out=self.cap_ptr[1:])
        self.cap = int(self.cap_ptr[-1])
        self.n_words = (self.cap >> 5) + 2 * self.n_reads + 2 if (self.n_reads and self.cap) else 0
        meta = torch.tensor([self.n_reads, self.cap, self.n_words], dtype=torch.int64, device=self.dev)
        allm = [torch.zeros(3, dtype=torch.int64, device=self.dev) for _ in range(world)]
        dist.all_gather(allm, meta)
        self.meta = [tuple(int(v) for v in m.cpu().tolist()) for m in allm]
        max_reads = max(m[0] for m in self.meta)
        cp = torch.zeros(max_reads + 1, dtype=torch.int64, device=self.dev)
        cp[:self.n_reads + 1] = torch.from_numpy(self.cap_ptr).to(self.dev)
        cps = [torch.zeros_like(cp) for _ in range(world)] if rank == 0 else None
        dist.gather(cp, cps, dst=0)
        self.peer_cap_ptr = self.recv_codes = self.recv_counts = self.dense = self.totals = None
        if rank == 0:
            self.peer_cap_ptr = [cps[r][:self.meta[r][0] + 1].cpu().numpy().copy() for r in range(world)]
            self.recv_codes = [None if r == 0 else torch.zeros((max(self.meta[r][2], 1), 2), dtype=torch.int32, device=self.dev)
                               for r in range(world)]
            self.recv_counts = [None if r == 0 else torch.zeros(max(self.meta[r][0], 1), dtype=torch.int32, device=self.dev)
                                for r in range(world)]
            self.dense = [torch.empty((max(self.meta[r][1], 1), 2), dtype=torch.int32, device=self.dev) for r in range(world)]
            self.totals = torch.zeros(world, dtype=torch.int64, device=self.dev)

    def _view(self, ptr: int, shape, n_elems_guard: int):
        """A tensor over library-owned memory (device memory on CUDA, host memory under the CPU emulation build)."""
        if n_elems_guard <= 0 or not ptr:
            return torch.zeros(shape, dtype=torch.int32, device=self.dev)[:0]
        if self.dev.type == "cuda":
            return torch.as_tensor(_DevArray(ptr, shape, "<i4"), device=self.dev)
        import ctypes
        n = int(np.prod(shape))
        return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_int32 * n).from_address(ptr)).reshape(shape))

    def gather(self, ctx):
        """Returns on rank 0 a list of (n_pairs int32 [n_reads_r], dense pairs int32 [total_r, 2]) per rank; None elsewhere."""
        _dp, dn, _cap, n = ctx.device_results()
        dc, n_words = ctx.device_codes()
        assert n == self.n_reads and n_words == self.n_words, "the batch does not have the shape this exchange was set up for"
        counts = self._view(dn, (max(n, 1),), n)
        codes = self._view(dc, (max(n_words, 1), 2), n_words)
        if self.rank != 0:
            if n_words:   # one grouped point-to-point operation per shard
                for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, codes, 0), dist.P2POp(dist.isend, counts, 0)]):
                    q.wait()
            return None
        ops = []
        for r in range(1, self.world):
            if self.meta[r][2]:
                ops += [dist.P2POp(dist.irecv, self.recv_codes[r], r), dist.P2POp(dist.irecv, self.recv_counts[r], r)]
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        if self.dev.type == "cuda":
            torch.cuda.current_stream(self.dev).synchronize()   # the library's kernels run on the context's own stream
        out_counts = []
        last = max((r for r in range(self.world) if self.meta[r][2]), default=-1)
        for r in range(self.world):
            c_r = counts if r == 0 else self.recv_counts[r][:self.meta[r][0]]
            out_counts.append(c_r)
            if not self.meta[r][2]:
                continue
            k_r = codes if r == 0 else self.recv_codes[r]
            ctx.expand_codes(k_r.data_ptr(), c_r.data_ptr(), self.peer_cap_ptr[r], self.dense[r].data_ptr(), self.meta[r][1],
                             total_ptr=self.totals[r:r + 1].data_ptr(), sync=(r == last))
        sizes = self.totals.cpu().tolist()       # the one host read of the exchange (after the last expansion has finished)
        return [(out_counts[r], self.dense[r][:int(sizes[r]) if self.meta[r][2] else 0]) for r in range(self.world)]


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
