"""Flat (CSR-style) ragged read batch — the Python-side mirror of ``abea_batch_t`` (include/abea_types.h).

This is the layout the reference's ``align_cuda`` flattens ``db_t`` into before its H2D copies
(reference src/f5c.cu:744-800: ``read``/``read_ptr``/``read_len``, ``event_table``/``event_ptr``/``n_events``,
``scalings``). Names follow the reference: reads, events, k-mers, scalings, aligned pairs.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

# reference: event_t, src/f5c.h:129-136 (24 bytes)
EVENT_DTYPE = np.dtype(
    {"names": ["start", "length", "mean", "stdv"],
     "formats": ["<u8", "<f4", "<f4", "<f4"],
     "offsets": [0, 8, 12, 16], "itemsize": 24})
# reference: scalings_t, src/f5c.h:158-172 (16 bytes)
SCALINGS_DTYPE = np.dtype([("scale", "<f4"), ("shift", "<f4"), ("var", "<f4"), ("log_var", "<f4")])
# reference: AlignedPair, src/f5c.h:181-184 (8 bytes)
PAIR_DTYPE = np.dtype([("ref_pos", "<i4"), ("read_pos", "<i4")])
# reference: model_t with CACHED_LOG, src/f5c.h:147-155 (12 bytes)
MODEL_DTYPE = np.dtype([("level_mean", "<f4"), ("level_stdv", "<f4"), ("level_log_stdv", "<f4")])

# reference: index_pair_t, src/f5c.h:187-190 (8 bytes)
INDEX_PAIR_DTYPE = np.dtype([("start", "<i4"), ("stop", "<i4")])
# abea_scaling_result_t (include/abea_types.h): what scaling_single (src/f5c.c:736-807) leaves behind per read
SCALING_RESULT_DTYPE = np.dtype([("scalings", SCALINGS_DTYPE), ("var_d", "<f8"), ("events_per_base", "<f8"),
                                 ("n_event_alignment", "<i4"), ("num_m_state", "<i4"), ("flags", "<u4"),
                                 ("calibrated", "<i4")])
assert SCALING_RESULT_DTYPE.itemsize == 48
FAILED_CALIBRATION, FAILED_ALIGNMENT, FAILED_QUALITY_CHK = 0x1, 0x2, 0x4   # src/f5c.h:66-68
MIN_NUM_EVENTS_TO_RESCALE = 200    # src/f5c.c:1185

ALN_BANDWIDTH = 100                # src/f5c.h:34
AVG_EVENTS_PER_KMER_MAX = 15.0     # src/f5cmisc.h:18


class CBatch(ctypes.Structure):
    """ctypes image of abea_batch_t."""
    _fields_ = [("n_reads", ctypes.c_int32),
                ("seq", ctypes.c_void_p),
                ("seq_ptr", ctypes.c_void_p),
                ("read_len", ctypes.c_void_p),
                ("events", ctypes.c_void_p),
                ("event_ptr", ctypes.c_void_p),
                ("n_events", ctypes.c_void_p),
                ("scalings", ctypes.c_void_p),
                ("good", ctypes.c_void_p),
                ("event_means", ctypes.c_void_p)]


@dataclass
class ReadBatch:
    """A batch of reads ready for ABEA.

    seq       uint8  [sum(read_len)+n]  sequences, each NUL-terminated (seq_ptr advances by read_len+1)
    seq_ptr   int64  [n]
    read_len  int32  [n]
    events    EVENT_DTYPE [sum(n_events)]
    event_ptr int64  [n]
    n_events  int32  [n]
    scalings  SCALINGS_DTYPE [n]
    good      uint8  [n]   (db->sig[i]->nsample > 0)
    kmer_size int
    """
    seq: np.ndarray
    seq_ptr: np.ndarray
    read_len: np.ndarray
    events: np.ndarray
    event_ptr: np.ndarray
    n_events: np.ndarray
    scalings: np.ndarray
    good: np.ndarray
    kmer_size: int
    meta: dict = field(default_factory=dict)

    @property
    def n_reads(self) -> int:
        return int(self.read_len.shape[0])

    @property
    def n_kmers(self) -> np.ndarray:
        return self.read_len.astype(np.int64) - self.kmer_size + 1

    @property
    def n_bands(self) -> np.ndarray:
        """NB = (E+1)+(K+1), reference src/align.c:219-221."""
        return self.n_events.astype(np.int64) + self.n_kmers + 2

    def pair_capacity(self) -> np.ndarray:
        """Per-read output capacity E+L pairs, as event_single allocates (reference src/f5c.c:724-726)."""
        return self.n_events.astype(np.int64) + self.read_len.astype(np.int64)

    def pair_ptr(self) -> np.ndarray:
        cap = self.pair_capacity()
        out = np.zeros(self.n_reads, dtype=np.int64)
        if self.n_reads > 1:
            np.cumsum(cap[:-1], out=out[1:])
        return out

    def map_ptr(self) -> np.ndarray:
        """Flat layout of the per-read base_to_event_map (reference src/f5c.c:746): read i owns max(K_i, 0) entries."""
        nk = np.maximum(self.n_kmers, 0)
        out = np.zeros(self.n_reads + 1, dtype=np.int64)
        np.cumsum(nk, out=out[1:])
        return out

    def eligible(self) -> np.ndarray:
        """align_single's filter (reference src/f5c.c:811-814): good read and events/base < 15 (float)."""
        ratio = self.n_events.astype(np.float32) / self.read_len.astype(np.float32)
        return (self.good != 0) & (ratio < np.float32(AVG_EVENTS_PER_KMER_MAX))

    def events_aligned(self) -> int:
        """Metric numerator: sum of E over reads passing the eligibility filter."""
        return int(self.n_events[self.eligible()].astype(np.int64).sum())

    def algorithmic_bytes(self, n_pairs: np.ndarray | None = None) -> int:
        """B(read) = 24E + (L+1) + 100*NB + 40*P + 36; model table counted once.

        P defaults to E (the survey's approximation) when the pair counts are not supplied.
        """
        el = self.eligible()
        E = self.n_events.astype(np.int64)[el]
        L = self.read_len.astype(np.int64)[el]
        NB = self.n_bands[el]
        P = E if n_pairs is None else np.asarray(n_pairs, dtype=np.int64)[el]
        per_read = 24 * E + (L + 1) + 100 * NB + 40 * P + 36
        return int(per_read.sum()) + 12 * (4 ** self.kmer_size)

    def event_means(self) -> np.ndarray:
        """The events' means alone as a flat float32 array (abea_batch_t.event_means): all ABEA reads of an event."""
        return np.ascontiguousarray(self.events["mean"], dtype=np.float32)

    def as_c(self, means: np.ndarray | None = None) -> CBatch:
        """ctypes image of the batch. means: a flat float32 array of event means (same indexing as `events`) to hand
        over instead of the 24-byte event table (abea_batch_t.event_means; `events` is then NULL)."""
        for name in ("seq", "seq_ptr", "read_len", "events", "event_ptr", "n_events", "scalings", "good"):
            a = getattr(self, name)
            assert a.flags["C_CONTIGUOUS"], name
        if means is not None:
            assert means.dtype == np.float32 and means.flags["C_CONTIGUOUS"] and means.shape[0] == self.events.shape[0]
        return CBatch(self.n_reads, self.seq.ctypes.data, self.seq_ptr.ctypes.data, self.read_len.ctypes.data,
                      None if means is not None else self.events.ctypes.data, self.event_ptr.ctypes.data,
                      self.n_events.ctypes.data, self.scalings.ctypes.data, self.good.ctypes.data,
                      means.ctypes.data if means is not None else None)

    def read_seq(self, i: int) -> bytes:
        p = int(self.seq_ptr[i])
        return self.seq[p:p + int(self.read_len[i])].tobytes()

    def read_events(self, i: int) -> np.ndarray:
        p = int(self.event_ptr[i])
        return self.events[p:p + int(self.n_events[i])]

    def subset(self, idx) -> "ReadBatch":
        """Gather a sub-batch (used for read-wise sharding across GPUs and for bounded CPU samples)."""
        idx = np.asarray(idx, dtype=np.int64)
        return ReadBatch.from_reads([self.read_seq(int(i)) for i in idx],
                                    [self.read_events(int(i)) for i in idx],
                                    self.scalings[idx].copy(), self.kmer_size,
                                    good=self.good[idx].copy(), meta=dict(self.meta))

    @staticmethod
    def from_reads(seqs, event_tables, scalings, kmer_size, good=None, meta=None) -> "ReadBatch":
        n = len(seqs)
        read_len = np.array([len(s) for s in seqs], dtype=np.int32)
        seq_ptr = np.zeros(n, dtype=np.int64)
        if n > 1:
            np.cumsum(read_len[:-1].astype(np.int64) + 1, out=seq_ptr[1:])
        seq = np.zeros(int(read_len.astype(np.int64).sum()) + n, dtype=np.uint8)
        for i, s in enumerate(seqs):
            p = int(seq_ptr[i])
            seq[p:p + len(s)] = np.frombuffer(bytes(s), dtype=np.uint8)
        n_events = np.array([len(e) for e in event_tables], dtype=np.int32)
        event_ptr = np.zeros(n, dtype=np.int64)
        if n > 1:
            np.cumsum(n_events[:-1].astype(np.int64), out=event_ptr[1:])
        events = np.zeros(int(n_events.astype(np.int64).sum()), dtype=EVENT_DTYPE)
        for i, e in enumerate(event_tables):
            p = int(event_ptr[i])
            events[p:p + len(e)] = e
        sc = np.zeros(n, dtype=SCALINGS_DTYPE)
        sc[:] = scalings
        g = np.ones(n, dtype=np.uint8) if good is None else np.ascontiguousarray(good, dtype=np.uint8)
        return ReadBatch(seq, seq_ptr, read_len, events, event_ptr, n_events, sc, g, int(kmer_size),
                         meta or {})
