"""ctypes binding of libabea_b200.so (include/abea_b200.h) and the host-side mirror of the reference call surface.

``AbeaContext`` plays the role of ``core_t``'s CUDA half (reference ``init_cuda``/``free_cuda``, src/f5c.cu:23-234):
it owns the device, the uploaded pore model and all device memory. ``align_db(ctx, batch)`` is the batch call the
reference spells ``align_db(core, db)`` -> ``align_cuda(core, db)`` (src/f5c.c:833-845, src/f5c.cu:647): it fills
``n_event_align_pairs`` and ``event_align_pairs`` for every read of the batch.

There is no CPU path: if the CUDA library is missing or no GPU is visible this module raises, loudly.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

from .batch import (INDEX_PAIR_DTYPE, MIN_NUM_EVENTS_TO_RESCALE, MODEL_DTYPE, PAIR_DTYPE, SCALING_RESULT_DTYPE,
                    SCALINGS_DTYPE, CBatch, ReadBatch)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libabea_b200.so")


class AbeaError(RuntimeError):
    pass


class CSignals(ctypes.Structure):
    """ctypes image of abea_signals_t."""
    _fields_ = [("n_reads", ctypes.c_int32), ("raw", ctypes.c_void_p), ("raw_ptr", ctypes.c_void_p),
                ("n_samples", ctypes.c_void_p), ("offset", ctypes.c_void_p), ("range", ctypes.c_void_p),
                ("digitisation", ctypes.c_void_p), ("raw_i16", ctypes.c_void_p)]


class CBlow5(ctypes.Structure):
    """ctypes image of abea_blow5_t."""
    _fields_ = [("n_reads", ctypes.c_int32), ("bytes", ctypes.c_void_p), ("rec_ptr", ctypes.c_void_p),
                ("rec_len", ctypes.c_void_p), ("record_method", ctypes.c_int32), ("signal_method", ctypes.c_int32)]


class CRagged(ctypes.Structure):
    """ctypes image of abea_ragged_t (one pointer per read, as db_t holds a batch)."""
    _fields_ = [("n_reads", ctypes.c_int32), ("seq", ctypes.c_void_p), ("read_len", ctypes.c_void_p),
                ("events", ctypes.c_void_p), ("n_events", ctypes.c_void_p), ("scalings", ctypes.c_void_p),
                ("good", ctypes.c_void_p), ("pairs", ctypes.c_void_p), ("n_pairs", ctypes.c_void_p)]


class Timing(ctypes.Structure):
    """abea_timing_t"""
    _fields_ = [("pack_ms", ctypes.c_double), ("h2d_ms", ctypes.c_double), ("kmer_ms", ctypes.c_double),
                ("fill_ms", ctypes.c_double), ("trace_ms", ctypes.c_double), ("kernel_ms", ctypes.c_double),
                ("d2h_ms", ctypes.c_double), ("unpack_ms", ctypes.c_double), ("h2d_bytes", ctypes.c_int64),
                ("d2h_bytes", ctypes.c_int64), ("kernel_launches", ctypes.c_int32),
                ("n_scheduled", ctypes.c_int32), ("n_wide", ctypes.c_int32), ("streamed", ctypes.c_int32),
                ("n_bands", ctypes.c_int64), ("n_events", ctypes.c_int64), ("load_ms", ctypes.c_double),
                ("mom_ms", ctypes.c_double), ("scaling_ms", ctypes.c_double), ("events_ms", ctypes.c_double),
                ("n_samples", ctypes.c_int64), ("ragged_ms", ctypes.c_double), ("blow5_ms", ctypes.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def _bind(path: str):
    if not os.path.exists(path):
        raise AbeaError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). f5c_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.abea_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    lib.abea_destroy.argtypes = [vp]
    lib.abea_destroy.restype = None
    lib.abea_last_error.argtypes = [vp]
    lib.abea_last_error.restype = ctypes.c_char_p
    lib.abea_set_model.argtypes = [vp, vp, ctypes.c_uint32]
    lib.abea_model_fill_log_stdv.argtypes = [vp, i64]
    lib.abea_model_fill_log_stdv.restype = None
    lib.abea_align_batch.argtypes = [vp, ctypes.POINTER(CBatch), vp, vp, vp, ctypes.POINTER(Timing)]
    lib.abea_align_ragged.argtypes = [vp, ctypes.POINTER(CRagged), ctypes.c_int, ctypes.POINTER(Timing)]
    lib.abea_scheduler_model.argtypes = [vp, vp]
    lib.abea_compact_results.argtypes = [vp, vp, i64, ctypes.POINTER(i64)]
    lib.abea_upload_batch.argtypes = [vp, ctypes.POINTER(CBatch), ctypes.POINTER(Timing)]
    lib.abea_run.argtypes = [vp, ctypes.POINTER(Timing)]
    lib.abea_download.argtypes = [vp, vp, vp, vp, ctypes.POINTER(Timing)]
    lib.abea_read_starts.argtypes = [vp, vp]
    lib.abea_read_respec.argtypes = [vp, vp]
    lib.abea_read_stats.argtypes = [vp, vp, vp, vp, vp]
    lib.abea_read_cycles.argtypes = [vp, vp, vp, vp]
    lib.abea_device_results.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i64),
                                        ctypes.POINTER(i32)]
    lib.abea_getevents.argtypes = [vp, ctypes.POINTER(CSignals), ctypes.c_int, vp, ctypes.POINTER(Timing)]
    lib.abea_getevents_download.argtypes = [vp, vp, vp]
    lib.abea_getevents_blow5.argtypes = [vp, ctypes.POINTER(CBlow5), ctypes.c_int, vp, vp, ctypes.POINTER(Timing)]
    lib.abea_raw_download.argtypes = [vp, vp, vp]
    lib.abea_write_pairs.argtypes = [ctypes.c_char_p, ctypes.c_int, i32, vp, vp, vp, vp, vp]
    lib.abea_write_resquiggle.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                          i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.abea_estimate_scalings.argtypes = [vp, ctypes.c_int, vp, ctypes.POINTER(Timing)]
    lib.abea_scaling_stage.argtypes = [vp, i32, ctypes.POINTER(Timing)]
    lib.abea_scaling_download.argtypes = [vp, vp, vp, vp]
    lib.abea_scaling_device_results.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i64)]
    lib.abea_host_alloc.argtypes = [ctypes.c_size_t]
    lib.abea_host_alloc.restype = vp
    lib.abea_host_free.argtypes = [vp]
    lib.abea_host_free.restype = None
    lib.abea_device_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p]
    lib.abea_host_threads.argtypes = [vp, ctypes.c_int]
    lib.abea_device_codes.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i64)]
    lib.abea_expand_codes.argtypes = [vp, vp, vp, vp, i32, vp, i64, vp, ctypes.POINTER(i64)]
    lib.abea_version.restype = ctypes.c_char_p
    return lib


_LIBS = {}


def load_library(path: str | None = None):
    path = path or os.environ.get("ABEA_LIB") or LIB_PATH   # ABEA_LIB: an experimental build of the same library (tools/)
    if path not in _LIBS:
        _LIBS[path] = _bind(path)
    return _LIBS[path]


@dataclass
class Alignment:
    """Output of align_db: the reference's db->event_align_pairs / db->n_event_align_pairs, flat."""
    pairs: np.ndarray      # PAIR_DTYPE, capacity layout (pair_ptr = prefix sum of E+L)
    pair_ptr: np.ndarray   # int64 [n]
    n_pairs: np.ndarray    # int32 [n]
    timing: dict

    def read_pairs(self, i: int) -> np.ndarray:
        p = int(self.pair_ptr[i])
        return self.pairs[p:p + int(self.n_pairs[i])]


@dataclass
class Scaling:
    """Output of scaling_db: per read what scaling_single (reference src/f5c.c:736-807) leaves in db_t —
    scalings[i] (recalibrated), events_per_base[i], read_stat_flag bits, n_event_alignment, base_to_event_map[i]."""
    results: np.ndarray    # SCALING_RESULT_DTYPE [n]
    maps: np.ndarray       # INDEX_PAIR_DTYPE, read i at map_ptr[i] .. map_ptr[i+1]
    map_ptr: np.ndarray    # int64 [n+1]
    timing: dict

    def read_map(self, i: int) -> np.ndarray:
        return self.maps[int(self.map_ptr[i]):int(self.map_ptr[i + 1])]


class AbeaContext:
    """One GPU's ABEA state (reference: init_cuda(core) ... free_cuda(core))."""

    def __init__(self, device: int = 0, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        self._h = ctypes.c_void_p()
        rc = self.lib.abea_create(ctypes.byref(self._h), device)
        if rc != 0:
            raise AbeaError(f"abea_create(device={device}) failed with {rc}: no usable CUDA device; "
                            "f5c_b200 has no CPU fallback")
        self.device = device
        self.kmer_size = None
        self._pinned = []

    # -- lifecycle ---------------------------------------------------------------------------------------
    def close(self):
        if self._h:
            for p in self._pinned:
                self.lib.abea_host_free(p)
            self._pinned = []
            self.lib.abea_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise AbeaError(f"{what} failed ({rc}): {self.lib.abea_last_error(self._h).decode()}")

    # -- model ---------------------------------------------------------------------------------------------
    def set_model(self, model: np.ndarray, kmer_size: int):
        """Upload a MODEL_DTYPE table of 4^k entries; level_log_stdv is (re)computed by the C library."""
        m = np.ascontiguousarray(model.copy())
        assert m.dtype == MODEL_DTYPE and m.shape[0] == 4 ** kmer_size
        self.lib.abea_model_fill_log_stdv(m.ctypes.data, m.shape[0])
        self._check(self.lib.abea_set_model(self._h, m.ctypes.data, kmer_size), "abea_set_model")
        self.kmer_size = kmer_size
        self.model = m
        return m

    def device_info(self):
        n = ctypes.c_int()
        name = ctypes.create_string_buffer(256)
        self.lib.abea_device_info(self._h, ctypes.byref(n), name)
        return n.value, name.value.decode()

    def host_threads(self, threads: int = -1) -> int:
        """Threads align_batch expands path codes with (abea_host_threads); a negative value only queries."""
        return int(self.lib.abea_host_threads(self._h, int(threads)))

    # -- pinned host memory ----------------------------------------------------------------------------
    def pinned_empty(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = self.lib.abea_host_alloc(max(n, 1))
        if not p:
            raise AbeaError("abea_host_alloc failed")
        self._pinned.append(p)
        buf = (ctypes.c_char * max(n, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def pin_batch(self, b: ReadBatch) -> ReadBatch:
        """Copy a batch's flat arrays into pinned host memory (so H2D runs at PCIe rate)."""
        def pin(a):
            out = self.pinned_empty(a.shape, a.dtype)
            out[...] = a
            return out
        return ReadBatch(pin(b.seq), pin(b.seq_ptr), pin(b.read_len), pin(b.events), pin(b.event_ptr),
                         pin(b.n_events), pin(b.scalings), pin(b.good), b.kmer_size, dict(b.meta))

    # -- the path ------------------------------------------------------------------------------------------
    def alloc_output(self, batch: ReadBatch, pinned: bool = False):
        cap = int(batch.pair_capacity().sum())
        if pinned:
            pairs = self.pinned_empty((cap,), PAIR_DTYPE)
            n_pairs = self.pinned_empty((batch.n_reads,), np.int32)
        else:
            pairs = np.empty(cap, dtype=PAIR_DTYPE)
            n_pairs = np.empty(batch.n_reads, dtype=np.int32)
        return pairs, batch.pair_ptr(), n_pairs

    def align_batch(self, batch: ReadBatch, out=None, means: np.ndarray | None = None) -> Alignment:
        """abea_align_batch: host buffers in, host buffers out (the e2e path). means: flat float32 event means handed
        over instead of the event table (abea_batch_t.event_means): 4 bytes per event cross PCIe instead of 24."""
        assert batch.kmer_size == self.kmer_size, "batch k-mer size does not match the uploaded model"
        pairs, pair_ptr, n_pairs = out if out is not None else self.alloc_output(batch)
        t = Timing()
        cb = batch.as_c(means)
        self._check(self.lib.abea_align_batch(self._h, ctypes.byref(cb), pairs.ctypes.data, pair_ptr.ctypes.data,
                                              n_pairs.ctypes.data, ctypes.byref(t)), "abea_align_batch")
        return Alignment(pairs, pair_ptr, n_pairs, t.as_dict())

    def pin_array(self, a: np.ndarray) -> np.ndarray:
        out = self.pinned_empty(a.shape, a.dtype)
        out[...] = a
        return out

    def ragged_view(self, batch: ReadBatch):
        """The batch as db_t holds it: per-read pointers into separately allocated sequences, event tables and pair
        buffers (each read's arrays are copied out of the flat batch so that nothing is contiguous across reads).
        Returns (CRagged, keep-alive list, per-read pair arrays, n_pairs)."""
        n = batch.n_reads
        seqs = [np.frombuffer(batch.read_seq(i) + b"\0", dtype=np.uint8).copy() for i in range(n)]
        evs = [np.ascontiguousarray(batch.read_events(i)).copy() for i in range(n)]
        cap = batch.pair_capacity()
        outs = [np.zeros(int(cap[i]) + 1, dtype=PAIR_DTYPE) if batch.good[i] else None for i in range(n)]
        P = ctypes.c_void_p * max(n, 1)
        seq_p = P(*[s.ctypes.data for s in seqs])
        ev_p = P(*[e.ctypes.data if len(e) else None for e in evs])
        out_p = P(*[o.ctypes.data if o is not None else None for o in outs])
        n_pairs = np.full(n, -7, dtype=np.int32)
        read_len = np.ascontiguousarray(batch.read_len, dtype=np.int32)
        n_events = np.ascontiguousarray(batch.n_events, dtype=np.int32)
        sc = np.ascontiguousarray(batch.scalings)
        good = np.ascontiguousarray(batch.good, dtype=np.uint8)
        rg = CRagged(n, ctypes.addressof(seq_p), read_len.ctypes.data, ctypes.addressof(ev_p), n_events.ctypes.data,
                     sc.ctypes.data, good.ctypes.data, ctypes.addressof(out_p), n_pairs.ctypes.data)
        keep = [seqs, evs, seq_p, ev_p, out_p, read_len, n_events, sc, good]
        return rg, keep, outs, n_pairs

    def align_ragged(self, batch: ReadBatch, threads: int = 4, view=None) -> Alignment:
        """abea_align_ragged on a db_t-like view of the batch; the result is gathered back into the flat capacity
        layout so that it compares with align_batch."""
        rg, keep, outs, n_pairs = view if view is not None else self.ragged_view(batch)
        t = Timing()
        self._check(self.lib.abea_align_ragged(self._h, ctypes.byref(rg), int(threads), ctypes.byref(t)), "abea_align_ragged")
        pp = batch.pair_ptr()
        pairs = np.zeros(int(batch.pair_capacity().sum()), dtype=PAIR_DTYPE)
        for i in range(batch.n_reads):
            if n_pairs[i] > 0:
                pairs[int(pp[i]):int(pp[i]) + int(n_pairs[i])] = outs[i][:int(n_pairs[i])]
        return Alignment(pairs, pp, n_pairs.copy(), t.as_dict())

    def scheduler_model(self) -> dict:
        m = np.zeros(4, dtype=np.float64)
        self._check(self.lib.abea_scheduler_model(self._h, m.ctypes.data), "abea_scheduler_model")
        return dict(cyc_wide=float(m[0]), cyc_narrow=float(m[1]), cyc_long=float(m[2]), cyc_trace=float(m[3]))

    def upload(self, batch: ReadBatch, with_scalings: bool = True, device_events: bool = False,
               means: np.ndarray | None = None) -> dict:
        """abea_upload_batch. with_scalings=False leaves abea_batch_t.scalings NULL: estimate_scalings() must follow.
        device_events=True leaves abea_batch_t.events NULL: the event tables of the last getevents() are aligned where
        they lie on the device (batch.n_events must repeat its counts)."""
        assert batch.kmer_size == self.kmer_size
        t = Timing()
        cb = batch.as_c(means)
        if not with_scalings:
            cb.scalings = None
        if device_events:
            cb.events = None
            cb.event_means = None
        self._check(self.lib.abea_upload_batch(self._h, ctypes.byref(cb), ctypes.byref(t)), "abea_upload_batch")
        return t.as_dict()

    def run(self) -> dict:
        t = Timing()
        self._check(self.lib.abea_run(self._h, ctypes.byref(t)), "abea_run")
        return t.as_dict()

    def download(self, batch: ReadBatch, out=None) -> Alignment:
        pairs, pair_ptr, n_pairs = out if out is not None else self.alloc_output(batch)
        t = Timing()
        self._check(self.lib.abea_download(self._h, pairs.ctypes.data, pair_ptr.ctypes.data, n_pairs.ctypes.data,
                                           ctypes.byref(t)), "abea_download")
        return Alignment(pairs, pair_ptr, n_pairs, t.as_dict())

    # -- event detection ---------------------------------------------------------------------------------
    def getevents(self, raw, raw_ptr, n_samples, calibration=None, rna: bool = False, download: bool = True):
        """abea_getevents + abea_getevents_download: the reference's getevents (src/events.c:562-582) per read.

        raw: float32 samples (ADC counts, or pA when calibration is None); calibration: (offset, range, digitisation)
        float32 arrays per read. Returns (events EVENT_DTYPE, event_ptr int64 [n], n_events int32 [n], timing)."""
        from .batch import EVENT_DTYPE
        raw_ptr = np.ascontiguousarray(raw_ptr, dtype=np.int64)
        n_samples = np.ascontiguousarray(n_samples, dtype=np.int32)
        n = int(n_samples.shape[0])
        if np.asarray(raw).dtype == np.int16:   # the int16 front door: ADC counts as the file holds them
            raw = np.ascontiguousarray(raw)
            cs = CSignals(n, None, raw_ptr.ctypes.data, n_samples.ctypes.data, None, None, None, raw.ctypes.data)
        else:
            raw = np.ascontiguousarray(raw, dtype=np.float32)
            cs = CSignals(n, raw.ctypes.data, raw_ptr.ctypes.data, n_samples.ctypes.data, None, None, None, None)
        keep = None
        if calibration is not None:
            keep = [np.ascontiguousarray(a, dtype=np.float32) for a in calibration]
            cs.offset, cs.range, cs.digitisation = (a.ctypes.data for a in keep)
        n_events = np.zeros(n, dtype=np.int32)
        t = Timing()
        self._check(self.lib.abea_getevents(self._h, ctypes.byref(cs), int(bool(rna)), n_events.ctypes.data,
                                            ctypes.byref(t)), "abea_getevents")
        if not download:   # the tables stay on the device for upload(..., device_events=True)
            return None, None, n_events, t.as_dict()
        cnt = np.maximum(n_events, 0).astype(np.int64)
        event_ptr = np.zeros(n, dtype=np.int64)
        if n > 1:
            np.cumsum(cnt[:-1], out=event_ptr[1:])
        events = np.zeros(int(cnt.sum()), dtype=EVENT_DTYPE)
        self._check(self.lib.abea_getevents_download(self._h, events.ctypes.data if len(events) else None,
                                                     event_ptr.ctypes.data), "abea_getevents_download")
        return events, event_ptr, n_events, t.as_dict()

    def getevents_blow5(self, payload: np.ndarray, rec_ptr, rec_len, record_method: int, signal_method: int,
                        rna: bool = False):
        """abea_getevents_blow5: BLOW5 records (their stored bytes, back to back in `payload`) -> event tables on the
        device. Returns (n_events int32 [n], n_samples int32 [n], timing); follow with getevents_download-style calls
        or upload(device_events=True)."""
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        rec_ptr = np.ascontiguousarray(rec_ptr, dtype=np.int64)
        rec_len = np.ascontiguousarray(rec_len, dtype=np.int32)
        n = int(rec_len.shape[0])
        cb = CBlow5(n, payload.ctypes.data, rec_ptr.ctypes.data, rec_len.ctypes.data, int(record_method), int(signal_method))
        n_events = np.zeros(n, dtype=np.int32)
        n_samples = np.zeros(n, dtype=np.int32)
        t = Timing()
        self._check(self.lib.abea_getevents_blow5(self._h, ctypes.byref(cb), int(bool(rna)), n_events.ctypes.data,
                                                  n_samples.ctypes.data, ctypes.byref(t)), "abea_getevents_blow5")
        return n_events, n_samples, t.as_dict()

    def events_download(self, n_events: np.ndarray):
        """abea_getevents_download into a freshly laid out flat table: (events, event_ptr)."""
        from .batch import EVENT_DTYPE
        cnt = np.maximum(n_events, 0).astype(np.int64)
        event_ptr = np.zeros(len(cnt), dtype=np.int64)
        if len(cnt) > 1:
            np.cumsum(cnt[:-1], out=event_ptr[1:])
        events = np.zeros(int(cnt.sum()), dtype=EVENT_DTYPE)
        self._check(self.lib.abea_getevents_download(self._h, events.ctypes.data if len(events) else None,
                                                     event_ptr.ctypes.data), "abea_getevents_download")
        return events, event_ptr

    def raw_download(self, n_samples: np.ndarray):
        """abea_raw_download: the float samples the last getevents worked on, (raw float32 flat, raw_ptr)."""
        cnt = np.maximum(n_samples, 0).astype(np.int64)
        raw_ptr = np.zeros(len(cnt), dtype=np.int64)
        if len(cnt) > 1:
            np.cumsum(cnt[:-1], out=raw_ptr[1:])
        raw = np.zeros(int(cnt.sum()), dtype=np.float32)
        self._check(self.lib.abea_raw_download(self._h, raw.ctypes.data if len(raw) else None, raw_ptr.ctypes.data),
                    "abea_raw_download")
        return raw, raw_ptr

    # -- the stages either side of the alignment ----------------------------------------------------------
    def estimate_scalings(self, n_reads: int, reverse_events: bool = False):
        """abea_estimate_scalings on the resident batch: (SCALINGS_DTYPE [n], timing). Reference
        estimate_scalings_using_mom (src/align.c:58-106) + the RNA event reversal (src/f5c.c:713-721)."""
        out = np.zeros(n_reads, dtype=SCALINGS_DTYPE)
        t = Timing()
        self._check(self.lib.abea_estimate_scalings(self._h, int(bool(reverse_events)), out.ctypes.data,
                                                    ctypes.byref(t)), "abea_estimate_scalings")
        return out, t.as_dict()

    def scaling_stage(self, min_num_events_to_rescale: int = MIN_NUM_EVENTS_TO_RESCALE) -> dict:
        t = Timing()
        self._check(self.lib.abea_scaling_stage(self._h, int(min_num_events_to_rescale), ctypes.byref(t)),
                    "abea_scaling_stage")
        return t.as_dict()

    def scaling_download(self, batch: ReadBatch, timing: dict | None = None) -> Scaling:
        mp = batch.map_ptr()
        res = np.zeros(batch.n_reads, dtype=SCALING_RESULT_DTYPE)
        maps = np.full(2 * int(mp[-1]), -1, dtype=np.int32).view(INDEX_PAIR_DTYPE)
        self._check(self.lib.abea_scaling_download(self._h, res.ctypes.data, maps.ctypes.data, mp.ctypes.data),
                    "abea_scaling_download")
        return Scaling(res, maps, mp, timing or {})

    def read_cycles(self, n_reads: int) -> dict:
        fc = np.zeros(n_reads, dtype=np.int64)
        tc = np.zeros(n_reads, dtype=np.int64)
        wd = np.zeros(n_reads, dtype=np.int32)
        self._check(self.lib.abea_read_cycles(self._h, fc.ctypes.data, tc.ctypes.data, wd.ctypes.data), "abea_read_cycles")
        return dict(fill_cycles=fc, trace_cycles=tc, wide=wd)

    def read_respec(self, n_reads: int) -> np.ndarray:
        rs = np.zeros(n_reads, dtype=np.int32)
        self._check(self.lib.abea_read_respec(self._h, rs.ctypes.data), "abea_read_respec")
        return rs

    def read_starts(self, n_reads: int) -> np.ndarray:
        st = np.zeros(n_reads, dtype=np.int32)
        self._check(self.lib.abea_read_starts(self._h, st.ctypes.data), "abea_read_starts")
        return st

    def device_results(self):
        """(pairs_ptr, n_pairs_ptr, total_pair_capacity, n_reads) of the last run, as raw device addresses."""
        dp, dn = ctypes.c_void_p(), ctypes.c_void_p()
        cap, n = ctypes.c_int64(), ctypes.c_int32()
        self._check(self.lib.abea_device_results(self._h, ctypes.byref(dp), ctypes.byref(dn), ctypes.byref(cap),
                                                 ctypes.byref(n)), "abea_device_results")
        return dp.value, dn.value, cap.value, n.value

    def compact_results(self, dst_ptr: int, dst_capacity: int) -> int:
        """abea_compact_results: the last run's pair lists packed back to back into device memory at dst_ptr
        (capacity in pairs); returns the number of pairs written."""
        total = ctypes.c_int64()
        self._check(self.lib.abea_compact_results(self._h, dst_ptr, int(dst_capacity), ctypes.byref(total)),
                    "abea_compact_results")
        return int(total.value)

    def device_codes(self):
        """(codes_ptr, n_words) of the last run's pair lists as path codes (abea_device_codes), a raw device address."""
        dp, n = ctypes.c_void_p(), ctypes.c_int64()
        self._check(self.lib.abea_device_codes(self._h, ctypes.byref(dp), ctypes.byref(n)), "abea_device_codes")
        return dp.value, n.value

    def expand_codes(self, codes_ptr: int, n_pairs_ptr: int, cap_ptr: np.ndarray, dst_ptr: int, dst_capacity: int,
                     total_ptr: int = 0, sync: bool = False) -> int:
        """abea_expand_codes: path codes + counts (device addresses) -> dense pair lists at dst_ptr (device). cap_ptr is
        the int64 capacity prefix sum of the reads the codes describe (host). Returns the total when sync, else -1."""
        cp = np.ascontiguousarray(cap_ptr, dtype=np.int64)
        total = ctypes.c_int64(-1)
        self._check(self.lib.abea_expand_codes(self._h, codes_ptr, n_pairs_ptr, cp.ctypes.data, int(cp.shape[0] - 1), dst_ptr,
                                               int(dst_capacity), total_ptr or None, ctypes.byref(total) if sync else None),
                    "abea_expand_codes")
        return int(total.value)

    def read_stats(self, n_reads: int) -> dict:
        se = np.zeros(n_reads, dtype=np.float64)
        na = np.zeros(n_reads, dtype=np.int32)
        ee = np.zeros(n_reads, dtype=np.int32)
        mg = np.zeros(n_reads, dtype=np.int32)
        self._check(self.lib.abea_read_stats(self._h, se.ctypes.data, na.ctypes.data, ee.ctypes.data, mg.ctypes.data),
                    "abea_read_stats")
        return dict(sum_emission=se, n_aligned=na, end_event=ee, max_gap=mg)


def align_db(ctx: AbeaContext, batch: ReadBatch) -> Alignment:
    """The reference's align_db(core, db) for the GPU build (src/f5c.c:833-845): ABEA for a data batch."""
    return ctx.align_batch(batch)


def scaling_db(ctx: AbeaContext, batch: ReadBatch,
               min_num_events_to_rescale: int = MIN_NUM_EVENTS_TO_RESCALE) -> Scaling:
    """The reference's scaling_db(core, db) (scaling_single per read, src/f5c.c:736-807) on the pair lists the last
    align left on the device: postalign + recalibrate_model + the read flags."""
    t = ctx.scaling_stage(min_num_events_to_rescale)
    return ctx.scaling_download(batch, t)


def write_pairs(path: str, names, aln: Alignment, flags: np.ndarray | None = None, append: bool = False, lib_path: str | None = None):
    """abea_write_pairs: the reference's --print-banded-aln dump (src/f5c.c:989-1006) of an Alignment."""
    lib = load_library(lib_path)
    n = len(names)
    arr = (ctypes.c_char_p * max(n, 1))(*[s.encode() if isinstance(s, str) else s for s in names])
    pairs = np.ascontiguousarray(aln.pairs)
    pp = np.ascontiguousarray(aln.pair_ptr, dtype=np.int64)
    npairs = np.ascontiguousarray(aln.n_pairs, dtype=np.int32)
    fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint32)
    rc = lib.abea_write_pairs(path.encode(), int(append), n, ctypes.cast(arr, ctypes.c_void_p), npairs.ctypes.data,
                              pairs.ctypes.data, pp.ctypes.data, None if fl is None else fl.ctypes.data)
    if rc != 0:
        raise AbeaError(f"abea_write_pairs failed ({rc})")


def write_resquiggle(path: str, names, read_len, n_samples, events: np.ndarray, event_ptr, results: np.ndarray, maps: np.ndarray,
                     map_ptr, kmer_size: int, fmt: str = "tsv", rna: bool = False, header: bool = True, append: bool = False,
                     lib_path: str | None = None):
    """abea_write_resquiggle: f5c resquiggle's TSV / PAF output (reference src/resquiggle.c:322-447) from the event tables
    (start, length) and what the scaling stage left per read (flags, scalings, k-mer -> event-range map)."""
    lib = load_library(lib_path)
    n = len(names)
    arr = (ctypes.c_char_p * max(n, 1))(*[s_.encode() if isinstance(s_, str) else s_ for s_ in names])
    rl = np.ascontiguousarray(read_len, dtype=np.int32)
    ns = np.ascontiguousarray(n_samples, dtype=np.int64)
    ev = np.ascontiguousarray(events)
    ep = np.ascontiguousarray(event_ptr, dtype=np.int64)
    res = np.ascontiguousarray(results)
    mp_ = np.ascontiguousarray(maps)
    mptr = np.ascontiguousarray(map_ptr, dtype=np.int64)
    rc = lib.abea_write_resquiggle(path.encode(), int(append), 1 if fmt == "paf" else 0, int(header), int(rna), int(kmer_size), n,
                                   ctypes.cast(arr, ctypes.c_void_p), rl.ctypes.data, ns.ctypes.data, ev.ctypes.data,
                                   ep.ctypes.data, res.ctypes.data, mp_.ctypes.data, mptr.ctypes.data)
    if rc != 0:
        raise AbeaError(f"abea_write_resquiggle failed ({rc})")

