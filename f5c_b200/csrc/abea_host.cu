/* abea_host.cu — host side of libabea_b200.so: context, device memory, batch packer, launches, unpacker.
 *
 * Replaces the GPU dispatch of the reference (src/f5c.cu: init_cuda :23-202, free_cuda :204-234, align_cuda
 * :647-1061) behind the C ABI in include/abea_b200.h. What is deliberately NOT reproduced: the CPU side-pool for
 * long / over-segmented reads (src/f5c.cu:243-452) and the load/memory advisors (:457-644) — every eligible read is
 * aligned on the GPU, scheduled longest-first onto persistent warps.
 *
 * There is no CPU implementation of the alignment in this library: without a CUDA device every entry point fails
 * with ABEA_ERR_NODEVICE.
 */
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sched.h>

#ifndef ABEA_SIMT_EMU
#include <cuda_runtime.h>
#define ABEA_LAUNCH(kern, grid, block, stream, ...) kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define ABEA_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#else /* tests/simt: CPU lock-step emulation of the same sources (test infrastructure) */
#define ABEA_LAUNCH(kern, grid, block, stream, ...) SIMT_LAUNCH(kern, grid, block, __VA_ARGS__)
#define ABEA_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) SIMT_LAUNCH_SMEM(kern, grid, block, smem, __VA_ARGS__)
#endif

#include "../../include/abea_b200.h"
#include "abea_kernels.cuh"
#include "scaling_kernels.cuh"
#include "events_kernels.cuh"
#include "blow5_kernels.cuh"

#define ABEA_VERSION_STR "abea-b200 0.1 (sm_100a)"

namespace {

double now_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

enum { EV_H2D0, EV_H2D1, EV_K0, EV_K1, EV_K2, EV_K3, EV_D2H0, EV_D2H1, EV_S0, EV_S1, EV_COUNT };

} // namespace

struct abea_devbuf {
    void* p = nullptr;
    size_t cap = 0;
};
struct abea_hostbuf { /* pinned */
    void* p = nullptr;
    size_t cap = 0;
};
typedef abea_devbuf DevBuf;
typedef abea_hostbuf HostBuf;

/* Host worker threads of abea_align_ragged: the packer / unpacker copies between the caller's ragged per-read arrays
 * and the pinned staging run here, overlapped with the kernels (the reference does these copies on the one thread
 * that calls align_cuda, before and after its kernels: src/f5c.cu:744-800, 1005-1030). */
struct abea_pool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv, cv_done;
    std::function<void(int)> job;
    int gen = 0, pending = 0;
    bool quit = false;
    void ensure(int n) {
        while ((int)th.size() < n) {
            const int id = (int)th.size();
            th.emplace_back([this, id]() {
                int seen = 0;
                for (;;) {
                    std::function<void(int)> f;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [&]() { return quit || (gen != seen && id < active); });
                        if (quit) return;
                        seen = gen;
                        f = job;
                    }
                    f(id);
                    {
                        std::lock_guard<std::mutex> lk(m);
                        if (--pending == 0) cv_done.notify_all();
                    }
                }
            });
        }
    }
    int active = 0;
    void kick(int n, std::function<void(int)> f) { /* returns at once; wait() joins */
        ensure(n);
        {
            std::lock_guard<std::mutex> lk(m);
            job = std::move(f);
            active = n;
            pending = n;
            gen++;
        }
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&]() { return pending == 0; });
    }
    ~abea_pool() {
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
        }
        cv.notify_all();
        for (std::thread& t : th) t.join();
    }
};

struct abea_ctx {
    int device = 0;
    int sm_count = 0;
    char dev_name[256] = {0};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[EV_COUNT] = {};
    char err[1024] = {0};

    /* model */
    DevBuf d_model;
    uint32_t kmer_size = 0;
    bool have_model = false;

    /* resident batch */
    DevBuf d_seq, d_events, d_means, d_reads, d_kparams, d_trace, d_pairs, d_results, d_queue, d_flags, d_npairs;
    DevBuf d_codes;                   /* path codes of the last run's pair lists (abea_code_t), made on demand by abea_device_codes */
    bool codes_current = false;       /* d_codes describes the last run */
    DevBuf d_xcap, d_xoff;            /* abea_expand_codes: a peer's capacity prefix sums, dense offsets */
    std::vector<int64_t> cap_ptr;     /* canonical pair_ptr of the caller's batch: prefix sum of E+L over ALL reads */
    HostBuf h_results, h_pairs, h_reads, h_items; /* pinned staging */
    bool prepared = false;            /* abea_prepare_kernel of the resident batch was already launched by the upload */
    int prep_launches = 0;
    std::vector<abea_read_t> reads;   /* scheduled reads, longest first */
    int32_t n_batch_reads = 0;        /* reads in the caller's batch */
    int64_t total_kmers = 0, total_trace_words = 0, total_pair_cap = 0, total_bands = 0, total_events = 0;
    bool uploaded = false, ran = false;
    abea_consts_t cst;
    abea_timing_t last = {};
    int32_t n_wide = 0;               /* the first n_wide scheduled reads (the longest) are filled by the wide kernel */
    cudaStream_t wide_stream = nullptr; /* the wide fill runs beside the narrow one */
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int wide_mode = 1;                /* ABEA_WIDE=0 disables the wide kernel */
    double wide_alpha = 0.8;          /* ABEA_WIDE_ALPHA scales the wide/narrow threshold (sweeps in profiles/) */
    int wide_cap = 0;                 /* ABEA_WIDE_CAP: most reads (= SMs) given to the wide kernel; default sm_count/4 */
    double wide_min_bands = 1024.0;   /* ABEA_WIDE_MIN_BANDS: reads shorter than this are never wide */
    int fill_ctas_per_sm = 1;  /* persistent narrow grid = sm_count * this (ABEA_FILL_CTAS_PER_SM) */
    int fill_warps_per_cta = 12; /* 4 primary + 8 secondary warps (ABEA_FILL_WARPS_PER_CTA, multiple of 4, <= 16; 12 measured best) */
    double long_alpha = 0.8;   /* ABEA_LONG_ALPHA: a read is "long" (runs alone on its sub-partition) above this share of the batch time */
    int sm_reserve = 0;        /* ABEA_SM_RESERVE: SMs left out of the narrow grid in addition to the wide CTAs' (when there are wide CTAs) */
    int sched_policy = 1;      /* ABEA_SCHED: 0 two-ended queue (secondary warps shortest-first), 1 longest-first for every warp */

    /* streaming (abea_align_batch with pinned host buffers): events pulled over PCIe by abea_load_kernel in the order
     * the fill asks for them, pair lists written to the caller's mapped buffer by the traceback */
    int stream_mode = 7;       /* ABEA_STREAM: bit 0 events streamed in; pair lists out: bit 2 as path codes expanded by host
                                * threads (wins when there are host threads), bit 1 written whole into the caller's mapped
                                * buffer; 0: copy engine both ways */
    int host_threads = 8;      /* ABEA_HOST_THREADS: threads of abea_align_batch that expand path codes (0: no codes) */
    int load_ctas = 64;        /* ABEA_LOAD_CTAS */
    int load_ctas_rag = 296;   /* ABEA_LOAD_CTAS_RAG: loader CTAs when the host packs while the loader runs (2 per SM) */
    int64_t load_piece = 0;    /* ABEA_LOAD_PIECE_KB: smallest piece of a read the loader delivers (0: 2048 events) */
    int64_t load_piece_cur = ABEA_LOAD_PIECE_BYTES; /* the value in use for the batch being streamed */
    double load_crit = 1.0;    /* ABEA_LOAD_CRIT: need times of the makespan-setting reads (wide, long) are scaled by this */
    DevBuf d_ready, d_items, d_capptr, d_dense_off;
    std::vector<abea_load_item_t> items;
    std::vector<int32_t> finish_order; /* scheduled reads by the time the replayed schedule expects them to finish */
    int32_t n_items = 0;        /* entries of `items` in use */
    int32_t n_first_items = 0;  /* of those, the first wave's first pieces: shipped before the rest of the list exists */
    bool pending_order = false; /* the rest of the list is still to be built and shipped (finish_load_order) */
    bool rag_mode = false;
    const void* ev_alias_cur = nullptr;
    std::atomic<int> rag_items_final{0};
    std::vector<double> sched_start, sched_finish; /* scratch of build_load_order, kept across batches */
    struct sched_slot_t { double t; int kind; };
    std::vector<sched_slot_t> sched_heap;
    struct sched_need_t { int32_t read, piece, bucket; };
    std::vector<sched_need_t> sched_need;
    std::vector<int32_t> sched_count;
    /* abea_align_ragged */
    abea_pool pool;
    HostBuf h_rseq, h_rmeans, h_rpairs, h_rnp, h_hostready, h_rmeta;
    HostBuf h_codes, h_fnp;    /* path codes and per-read counts written by the traceback (mapped) */
    std::atomic<int> rag_items_ready{0};
    /* Cycles per band of the three forms a read is filled in, cycles per traceback step and the band time of a fully
     * loaded sub-partition: the scheduler's model of the kernels. Starting values measured on B200 at 1.965 GHz
     * (profiles/); re-derived from the per-read clock64 counts of the batches that run (calibrate()). */
    double cyc_wide = 400.0, cyc_narrow = 856.0, cyc_long = 626.0, cyc_trace = 100.0; /* measured on B200 with this revision of the kernels (profiles/read_cycles_final_r02_cfg5.txt) */
    int calib_mode = 1;        /* ABEA_CALIBRATE=0 keeps the starting values */
    size_t wide_excl_bytes = (size_t)160 * 1024; /* dynamic shared memory a wide CTA asks for to have its SM to itself */
    int tb_mode = 1;           /* ABEA_TB: 1 segment-parallel traceback (a walk per lane), 0 the serial walk */
    int tb_margin = 64;        /* ABEA_TB_MARGIN: bands a speculative walk starts above its segment */
    int calib_runs = 0;
    cudaStream_t load_stream = nullptr;
    cudaEvent_t ev_meta = nullptr, ev_loaded = nullptr, ev_load0 = nullptr;
    bool streaming = false;    /* the resident batch is being streamed in: its fill must wait on d_ready */
    bool results_on_device = false; /* d_pairs / d_npairs hold the final lists of the last run */
    int64_t event_bytes = 0;   /* size in bytes of the batch's event array as the caller holds it (means: 4 B, AoS: 24 B per event) */
    bool load_aos = false;     /* the streamed source is the reference's AoS event table (else a flat array of means) */
    bool means_from_evcap = false; /* d_means was extracted from the tables abea_getevents left on the device */
    bool means_reversed = false;   /* abea_estimate_scalings(reverse_events) has turned the resident means 3'->5' */
    int64_t evcap_total = 0;   /* entries of d_evcap (capacity layout of the last abea_getevents) */

    /* the stages either side of ABEA (scaling_kernels.cuh): descriptors of ALL reads in the caller's order */
    std::vector<abea_sread_t> sreads;
    std::vector<abea_scalings_t> in_scalings; /* the caller's scalings (empty: to be estimated on the device) */
    DevBuf d_sreads, d_scalings, d_maps, d_sres;
    int64_t total_map = 0;           /* entries of d_maps: sum of max(K, 0) over the batch */
    bool sreads_on_device = false;   /* d_sreads holds the resident batch's descriptors */
    bool scalings_on_device = false; /* d_scalings holds the batch's scalings (uploaded or estimated) */
    bool need_scalings = false;      /* the batch came without scalings: abea_estimate_scalings must run before abea_run */
    bool scaled = false;             /* abea_scaling_stage has run on the last abea_run's results */

    /* event detection (events_kernels.cuh) */
    std::vector<abea_sig_t> sigs;
    DevBuf d_raw, d_sum, d_sumsq, d_ts1, d_ts2, d_peaks, d_evcap, d_sigs, d_sigorder, d_nev, d_evptr, d_evout;
    DevBuf d_chunks, d_spec, d_fix, d_spec_cnt, d_fix_cnt, d_sync, d_spec_end;
    DevBuf d_raw16, d_b5in, d_b5out, d_b5recs, d_b5len, d_b5status, d_b5hdr, d_b5rawoff, d_b5ex; /* BLOW5 decode (blow5_kernels.cuh) */
    HostBuf h_b5;
    int evt_chunk = 1024;            /* ABEA_EVT_CHUNK: samples per chunk of the speculative peak detector (multiple of 4) */
    std::vector<int32_t> nev;        /* event counts of the last abea_getevents */
    bool events_ready = false;
};

namespace {

int fail(abea_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU(call)                                                                                            \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return fail(c, ABEA_ERR_CUDA, "Cuda error: %s (%s) at %s:%d", cudaGetErrorString(e_), #call,    \
                        __FILE__, __LINE__);                                                                \
    } while (0)

int dev_reserve(abea_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256; /* head-room so that similar batches do not reallocate */
    CU(cudaMalloc(&b.p, want));
    /* once per growth, not per batch: the traceback stages whole 32-band chunks of trace lines, so it copies (and
     * never looks at) the padding words of a line and the lines past a read's last band */
    CU(cudaMemsetAsync(b.p, 0, want, c->stream));
    /* growth is rare; the buffer may next be written from another stream of the context (the loader's) */
    CU(cudaStreamSynchronize(c->stream));
    b.cap = want;
    return 0;
}

int host_reserve(abea_ctx* c, HostBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CU(cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CU(cudaMallocHost(&b.p, want));
    b.cap = want;
    return 0;
}

float ev_ms(abea_ctx* c, int a, int b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) return 0.f;
    return ms;
}

/* the device-side alias of a pinned (cudaHostAlloc / cudaHostRegister) host pointer, or NULL */
void* mapped_alias(const void* host) {
    if (!host) return nullptr;
#ifdef ABEA_SIMT_EMU
    return (void*)host;
#else
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type != cudaMemoryTypeHost) return nullptr;
    return at.devicePointer;
#endif
}


/* abea_prepare_kernel over the scheduled reads; check_events = 0 when the loader validates the event means */
void launch_prepare(abea_ctx* c, int64_t check_events) {
    const int32_t n = (int32_t)c->reads.size();
    const int threads = 256;
    const int64_t work = std::max(c->total_kmers, check_events);
    int blocks = (int)std::min<int64_t>((work + threads - 1) / threads, (int64_t)c->sm_count * 16);
    if (blocks < 1) blocks = 1;
    ABEA_LAUNCH(abea_prepare_kernel, blocks, threads, c->stream,
        (const abea_read_t*)c->d_reads.p, n, (const uint8_t*)c->d_seq.p, (const abea_model_t*)c->d_model.p,
        c->kmer_size, (const float*)c->d_means.p, (float4*)c->d_kparams.p, (uint32_t*)c->d_flags.p,
        c->total_kmers, check_events);
}

/* The scheduler's thresholds, from its model of the kernels (abea_ctx::cyc_*).
 * cyc_tput: band time of a sub-partition that runs its full share of narrow warps (warps per CTA / 4 of them). */
double cyc_tput(const abea_ctx* c) { return c->cyc_narrow / std::max(1.0, (double)c->fill_warps_per_cta / 4.0); }
double batch_cycles(const abea_ctx* c) { return (double)c->total_bands * cyc_tput(c) * 1.08 / ((double)c->sm_count * 4.0); }
/* a narrow read is "long" when, sharing its sub-partition, it would take more than long_alpha of the time the whole
 * batch needs at full throughput: it then runs alone on its sub-partition */
int32_t long_threshold(const abea_ctx* c) {
    /* The two thresholds below were tuned by sweeps (profiles/sweep_*): on the target config a read is long above
     * 15.4 k bands and wide above 23.0 k. Their constants (0.904, 1.40) are what reproduces those optima from the
     * MEASURED cycle counts, so that the thresholds do not move when the model is re-derived from a batch's own counts
     * (round 2 found them tuned against stale start values: after the first calibration 18 reads went wide instead of
     * 6 and the step took 8.47 instead of 8.12 ms). */
    const double shared = 0.5 * (c->cyc_narrow + c->cyc_long) * 0.904; /* two warps per sub-partition */
    return (int32_t)std::min(2.0e9, std::max(1.0, c->long_alpha * batch_cycles(c) / shared));
}
/* a read goes to the wide kernel when it would outlast the whole batch even as a lone narrow warp */
double wide_threshold(const abea_ctx* c, int64_t total_bands) {
    const double cyc_batch = (double)total_bands * cyc_tput(c) * 1.08 / ((double)c->sm_count * 4.0);
    return std::max(c->wide_min_bands, c->wide_alpha * 1.40 * cyc_batch / c->cyc_long);
}

/* A pair list from its path codes (abea_code_t, abea_kernels.cuh): the first pair, then one step per bit. Measured: a
 * host core expands ~0.8 G pairs/s, bound by writing the list to memory — a 256-entry table of four-step running sums
 * with 64-bit packed pairs and non-temporal stores were both slower than this loop (profiles/path_codes_r02.txt). */
void decode_codes(const abea_code_t* w, int32_t n, abea_pair_t* dst) {
    if (n <= 0) return;
    int32_t k = (int32_t)w[0].a, e = (int32_t)w[0].b;
    dst[0].ref_pos = k;
    dst[0].read_pos = e;
    int32_t j = 1;
    for (const abea_code_t* q = w + 1; j < n; q++) {
        uint32_t mk = q->a, me = q->b;
        const int32_t lim = std::min<int32_t>(32, n - j);
        for (int32_t t = 0; t < lim; t++) {
            k += (int32_t)(mk & 1u);
            e += (int32_t)(me & 1u);
            mk >>= 1;
            me >>= 1;
            dst[j].ref_pos = k;
            dst[j].read_pos = e;
            j++;
        }
    }
}
int64_t code_words(int64_t total_pair_cap, int32_t n_reads) { return (total_pair_cap >> 5) + 2 * (int64_t)n_reads + 2; }
int64_t code_bytes_used(int32_t np) { return np > 0 ? (int64_t)(1 + (np - 1 + 31) / 32) * (int64_t)sizeof(abea_code_t) : 0; }

/* see upload_impl */
void build_load_order(abea_ctx* c, bool want_finish_order) {
    const int64_t n = (int64_t)c->reads.size();
    const int nw = c->n_wide;
    const int wpc = c->fill_warps_per_cta;
    const int wide_sms = nw > 0 ? std::min(c->sm_count, nw + c->sm_reserve) : 0;
    int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((int64_t)(c->sm_count - wide_sms) * c->fill_ctas_per_sm,
                                                            (n - nw + wpc - 1) / wpc));
    if (nw > 0 && blocks > 1) blocks &= ~(int64_t)1; /* as in run_impl */
    const int64_t n_pri = 4 * blocks, n_sec = (int64_t)(wpc - 4) * blocks;
    /* cycles per band (wide CTA; narrow warp sharing its sub-partition; narrow warp alone on it = a "long" read) and
     * per traceback step. The longest reads set the makespan, so their pieces are asked for early rather than late. */
    const double CYC_WIDE = c->cyc_wide, CYC_NARROW = c->cyc_narrow, CYC_LONG = c->cyc_long, CYC_TRACE = c->cyc_trace;
    const double long_thr = (double)long_threshold(c); /* as in run_impl */
    auto rate = [&](int64_t r) {
        if (r < nw) return CYC_WIDE;
        const double nb = (double)c->reads[r].n_events + c->reads[r].n_kmers + 2;
        return nb > long_thr ? CYC_LONG : CYC_NARROW;
    };
    std::vector<double>& start = c->sched_start;
    std::vector<double>& finish = c->sched_finish;
    start.assign((size_t)n, 0.0);
    finish.assign((size_t)n, 0.0);
    /* (time a warp becomes free, kind: 0 wide, 1 primary, 2 secondary) in a binary min-heap that is updated in place:
     * the slot at the root takes the next read and sinks (this runs on the critical path of every streamed batch) */
    typedef abea_ctx::sched_slot_t slot_t;
    std::vector<slot_t>& heap = c->sched_heap;
    heap.clear();
    /* at t = 0 the three kinds ask in this order of urgency; the tiny offsets only order the first requests: the
     * reads that set the makespan first, and primary and secondary warps served alternately after that */
    for (int i = 0; i < std::min<int64_t>(nw, c->sm_count); i++) heap.push_back(slot_t{0.0, 0});
    for (int64_t i = 0; i < n_pri; i++) heap.push_back(slot_t{1e-3 * (double)i / (double)n_pri, 1});
    for (int64_t i = 0; i < n_sec; i++) heap.push_back(slot_t{1.5e-3 * (double)i / (double)std::max<int64_t>(1, n_sec), 2});
    std::make_heap(heap.begin(), heap.end(), [](const slot_t& x, const slot_t& y) { return x.t > y.t; });
    auto sift_down = [&]() {
        const size_t hn = heap.size();
        size_t i = 0;
        const slot_t v = heap[0];
        for (;;) {
            size_t l = 2 * i + 1;
            if (l >= hn) break;
            if (l + 1 < hn && heap[l + 1].t < heap[l].t) l++;
            if (!(heap[l].t < v.t)) break;
            heap[i] = heap[l];
            i = l;
        }
        heap[i] = v;
    };
    int64_t w = 0, h = nw, t = n - 1;
    double t_max = 0.0;
    while ((w < nw || h <= t) && !heap.empty()) {
        slot_t& sl = heap[0];
        int64_t r = -1;
        if (sl.kind == 0) {
            if (w < nw) r = w++;
        } else if (h <= t) {
            r = (sl.kind == 1 || c->sched_policy == 1) ? h++ : t--;
        }
        if (r < 0) { /* nothing left for this kind of slot: it leaves the heap */
            heap[0] = heap.back();
            heap.pop_back();
            if (!heap.empty()) sift_down();
            continue;
        }
        const abea_read_t& rd = c->reads[r];
        const double nb = (double)rd.n_events + rd.n_kmers + 2;
        start[r] = sl.t;
        sl.t += nb * rate(r) + (double)rd.n_events * CYC_TRACE;
        finish[r] = sl.t;
        if (sl.t > t_max) t_max = sl.t;
        sift_down();
    }
    /* every (read, piece) with the time the fill first touches it: event e when the read is about e/E of the way through
     * its bands. Ordered by that time with a counting sort over 8192 time buckets (the order inside a bucket — 1/8192 of
     * the batch — does not matter to a heuristic; a comparison sort of the ~10^4 items was most of this function). */
    typedef abea_ctx::sched_need_t need_t;
    std::vector<need_t>& need = c->sched_need;
    need.clear();
    const int64_t esz = c->load_aos ? (int64_t)sizeof(abea_event_t) : (int64_t)sizeof(float);
    const int NBUCKET = 8192;
    const double to_bucket = t_max > 0.0 ? (double)(NBUCKET - 1) / t_max : 0.0;
    std::vector<int32_t>& count = c->sched_count;
    count.assign((size_t)NBUCKET + 1, 0);
    for (int64_t r = 0; r < n; r++) {
        const abea_read_t& rd = c->reads[r];
        const abea_load_geom_t g = abea_load_geom(rd.ev_off, rd.n_events, c->event_bytes, c->load_piece_cur, esz);
        const double nb = (double)rd.n_events + rd.n_kmers + 2;
        const double rt = rate(r);
        const double fill = nb * rt * (rt < CYC_NARROW ? c->load_crit : 1.0);
        const double per_byte = fill / (double)(g.b - g.a); /* time per byte of the read's own range */
        for (int32_t q = (r < c->n_first_items) ? 1 : 0; q < g.n_pieces; q++) {
            const int64_t off = g.lo + (int64_t)q * g.piece - g.a; /* first byte of the piece, relative to the read's first */
            const double tq = start[r] + (off > 0 ? (double)off * per_byte : 0.0);
            int bk = (int)(tq * to_bucket);
            bk = bk < 0 ? 0 : (bk >= NBUCKET ? NBUCKET - 1 : bk);
            need.push_back(need_t{(int32_t)r, q, bk});
            count[(size_t)bk + 1]++;
        }
    }
    for (int i = 0; i < NBUCKET; i++) count[(size_t)i + 1] += count[(size_t)i];
    /* the first pieces of the first wave are already in the list (and on their way): the rest follows them */
    const size_t base = (size_t)c->n_first_items;
    if (c->items.size() < base + need.size()) c->items.resize(base + need.size());
    for (const need_t& x : need) c->items[base + (size_t)count[(size_t)x.bucket]++] = abea_load_item_t{x.read, x.piece};
    c->n_items = (int32_t)(base + need.size());
    if (want_finish_order) {
        c->finish_order.resize((size_t)n);
        for (int64_t r = 0; r < n; r++) c->finish_order[(size_t)r] = (int32_t)r;
        std::stable_sort(c->finish_order.begin(), c->finish_order.end(), [&](int32_t x, int32_t y) { return finish[x] < finish[y]; });
    }
}

/* Re-derive the scheduler's model from what the batch that has just run measured: every read reports the SM cycles of
 * its fill and of its traceback (abea_result_t, in-kernel clock64), so the medians per form replace the constants a
 * particular clock, SM count or kernel revision was tuned on. Only resident runs are used (a streamed read's cycles
 * include its waits for PCIe); values are clamped to [1/2, 2] x the starting ones, and a class updates only when it has
 * enough reads to have a meaningful median. */
int calibrate(abea_ctx* c, int32_t long_thr) {
    const size_t n = c->reads.size();
    if (n < 256) return 0;
    std::vector<abea_result_t> res(n);
    CU(cudaMemcpyAsync(res.data(), c->d_results.p, n * sizeof(abea_result_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<double> wide, lng, shared, trace;
    for (size_t j = 0; j < n; j++) {
        const double nb = (double)c->reads[j].n_events + c->reads[j].n_kmers + 2;
        if (nb < 512 || res[j].fill_cycles <= 0) continue;
        const double per = (double)res[j].fill_cycles / nb;
        if (res[j].wide) wide.push_back(per);
        else if (nb > (double)long_thr) lng.push_back(per);
        else shared.push_back(per);
        if (res[j].n_aligned > 256 && res[j].trace_cycles > 0) trace.push_back((double)res[j].trace_cycles / res[j].n_aligned);
    }
    auto median = [](std::vector<double>& v) {
        std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end());
        return v[v.size() / 2];
    };
    auto clampd = [](double x, double ref) { return std::min(2.0 * ref, std::max(0.5 * ref, x)); };
    if (wide.size() >= 2) c->cyc_wide = clampd(median(wide), 400.0);
    if (lng.size() >= 4) c->cyc_long = clampd(median(lng), 626.0);
    if (shared.size() >= 64) c->cyc_narrow = clampd(median(shared), 856.0);
    if (trace.size() >= 64) c->cyc_trace = std::min(700.0, std::max(2.0, median(trace))); /* 350 serial, a few tens segment-parallel */
    return 0;
}

/* items [first, first + count) of the loader's work list on the loader's stream; `slot`: its counter in d_queue */
int launch_loader(abea_ctx* c, int32_t first, int32_t count, int slot) {
    if (count <= 0) return ABEA_OK;
    /* A loader CTA has one item in flight. When the items are gated by the host packers (ragged front door) an item's
     * latency includes the poll of its flag over PCIe and the host's memory system is busy with the packers: 64 CTAs
     * then bound the batch (align_cuda 12.6 ms; 12.1 with 148, 11.6 with 296; profiles/dropin_experiments_r02.txt),
     * while with everything already in pinned memory 64 are the fastest (10.47 / 10.60 / 10.71 ms). */
    const int ctas = c->rag_mode ? c->load_ctas_rag : c->load_ctas;
    const int blocks = (int)std::min<int64_t>((int64_t)count, (int64_t)ctas);
    const uint32_t* host_ready = c->rag_mode ? (const uint32_t*)mapped_alias(c->h_hostready.p) + first : nullptr;
    if (c->load_aos)
        ABEA_LAUNCH(abea_load_kernel<true>, blocks, ABEA_LOAD_THREADS, c->load_stream, (const abea_read_t*)c->d_reads.p,
                    (const abea_load_item_t*)c->d_items.p + first, count, (const uint4*)c->ev_alias_cur,
                    (float*)c->d_means.p, c->event_bytes, (uint32_t*)c->d_flags.p, (uint32_t*)c->d_ready.p,
                    (int32_t*)c->d_queue.p + slot, c->load_piece_cur, (const volatile uint32_t*)host_ready,
                    (uint32_t*)c->d_queue.p + 15);
    else
        ABEA_LAUNCH(abea_load_kernel<false>, blocks, ABEA_LOAD_THREADS, c->load_stream, (const abea_read_t*)c->d_reads.p,
                    (const abea_load_item_t*)c->d_items.p + first, count, (const uint4*)c->ev_alias_cur,
                    (float*)c->d_means.p, c->event_bytes, (uint32_t*)c->d_flags.p, (uint32_t*)c->d_ready.p,
                    (int32_t*)c->d_queue.p + slot, c->load_piece_cur, (const volatile uint32_t*)host_ready,
                    (uint32_t*)c->d_queue.p + 15);
    CU(cudaGetLastError());
    return ABEA_OK;
}

/* the rest of the loader's work list, built while the GPU is already filling the first wave (see upload_impl) */
int finish_load_order(abea_ctx* c) {
    const double t0 = now_ms();
    build_load_order(c, c->rag_mode);
    const int32_t rest = c->n_items - c->n_first_items;
    if (rest > 0) {
        memcpy((abea_load_item_t*)c->h_items.p + c->n_first_items, c->items.data() + c->n_first_items, (size_t)rest * sizeof(abea_load_item_t));
        CU(cudaMemcpyAsync((abea_load_item_t*)c->d_items.p + c->n_first_items, (abea_load_item_t*)c->h_items.p + c->n_first_items,
                           (size_t)rest * sizeof(abea_load_item_t), cudaMemcpyHostToDevice, c->load_stream));
    }
    if (c->rag_mode) {
        c->rag_items_ready.store(c->n_items, std::memory_order_release);
        c->rag_items_final.store(1, std::memory_order_release);
    }
    const int rc = launch_loader(c, c->n_first_items, rest, 13);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev_loaded, c->load_stream));
    c->pending_order = false;
    if (getenv("ABEA_TIME_PACK")) fprintf(stderr, "[abea pack] rest of the load order %.3f ms (%d items), behind the fill launches\n", now_ms() - t0, rest);
    return ABEA_OK;
}

} // namespace

extern "C" {

const char* abea_version(void) { return ABEA_VERSION_STR; }

const char* abea_last_error(const abea_ctx_t* ctx) { return ctx ? ctx->err : "null context"; }

void abea_model_fill_log_stdv(abea_model_t* model, int64_t n) {
    /* the reference is C++: log(float) is the float overload == logf (src/model.c:93,179) */
    for (int64_t i = 0; i < n; i++) model[i].level_log_stdv = logf(model[i].level_stdv);
}

void* abea_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

void abea_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int abea_create(abea_ctx_t** out, int device) {
    if (!out) return ABEA_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        fprintf(stderr, "[abea_create::ERROR] no usable CUDA device (requested %d of %d); this library has no CPU path\n",
                device, n);
        return ABEA_ERR_NODEVICE;
    }
    abea_ctx* c = new abea_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete c;
        return ABEA_ERR_CUDA;
    }
    c->sm_count = prop.multiProcessorCount;
    snprintf(c->dev_name, sizeof(c->dev_name), "%s", prop.name);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return ABEA_ERR_CUDA;
    }
    for (int i = 0; i < EV_COUNT; i++) {
        if (cudaEventCreate(&c->ev[i]) != cudaSuccess) {
            delete c;
            return ABEA_ERR_CUDA;
        }
    }
    /* transition constants shared by all reads, host double (reference src/align.c:212-216) */
    if (const char* e = getenv("ABEA_WIDE")) c->wide_mode = atoi(e);
    if (const char* e = getenv("ABEA_WIDE_ALPHA")) c->wide_alpha = atof(e);
    c->wide_cap = std::max(1, c->sm_count / 4);
    if (const char* e = getenv("ABEA_WIDE_CAP")) c->wide_cap = std::max(0, atoi(e));
    if (const char* e = getenv("ABEA_WIDE_MIN_BANDS")) c->wide_min_bands = atof(e);
    if (cudaStreamCreateWithFlags(&c->wide_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev_fork) != cudaSuccess || cudaEventCreate(&c->ev_join) != cudaSuccess) {
        delete c;
        return ABEA_ERR_CUDA;
    }
    if (const char* e = getenv("ABEA_FILL_CTAS_PER_SM")) c->fill_ctas_per_sm = std::max(1, atoi(e));
    if (const char* e = getenv("ABEA_FILL_WARPS_PER_CTA")) c->fill_warps_per_cta = std::min(ABEA_NARROW_WARPS_MAX, std::max(4, atoi(e) / 4 * 4));
    if (const char* e = getenv("ABEA_LONG_ALPHA")) c->long_alpha = atof(e);
    if (const char* e = getenv("ABEA_CALIBRATE")) c->calib_mode = atoi(e);
    if (const char* e = getenv("ABEA_TB")) c->tb_mode = atoi(e) & 3;
    if (const char* e = getenv("ABEA_TB_MARGIN")) c->tb_margin = std::max(0, atoi(e));
    if (const char* e = getenv("ABEA_SCHED")) c->sched_policy = atoi(e) ? 1 : 0;
    if (const char* e = getenv("ABEA_SM_RESERVE")) c->sm_reserve = std::max(0, atoi(e));
    if (const char* e = getenv("ABEA_STREAM")) c->stream_mode = atoi(e);
    {
        cpu_set_t set;
        CPU_ZERO(&set);
        int cpus = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : 1;
        c->host_threads = std::max(1, std::min(8, cpus));
    }
    if (const char* e = getenv("ABEA_HOST_THREADS")) c->host_threads = std::min(64, std::max(0, atoi(e)));
    c->load_ctas_rag = 2 * c->sm_count;
    if (const char* e = getenv("ABEA_LOAD_CTAS")) c->load_ctas = c->load_ctas_rag = std::max(1, atoi(e));
    if (const char* e = getenv("ABEA_LOAD_CTAS_RAG")) c->load_ctas_rag = std::max(1, atoi(e));
    if (const char* e = getenv("ABEA_LOAD_PIECE_KB")) c->load_piece = (int64_t)std::max(1, atoi(e)) * 1024;
    if (const char* e = getenv("ABEA_LOAD_CRIT")) c->load_crit = atof(e);
    if (const char* e = getenv("ABEA_EVT_CHUNK")) c->evt_chunk = std::max(8, atoi(e) / 4 * 4);
    if (cudaStreamCreateWithFlags(&c->load_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev_meta) != cudaSuccess || cudaEventCreate(&c->ev_loaded) != cudaSuccess ||
        cudaEventCreate(&c->ev_load0) != cudaSuccess) {
        delete c;
        return ABEA_ERR_CUDA;
    }
    /* Every kernel asks for the SAME shared-memory carve-out. An SM cannot change its L1/shared split while CTAs are
     * resident, so kernels with different preferences keep each other off an SM until the other's CTAs have left —
     * measured twice: a loader CTA with a small carve-out held the persistent fill CTAs back until the loader had
     * finished (round 1), and loader CTAs at 35 % held the wide CTAs (then still at 100 %) back, so that the longest
     * reads started when the stream had ended (align_cuda 16.1 ms instead of 13.4; profiles/dropin_breakdown_r02.txt).
     * ABEA_CARVEOUT: percent of the maximum shared memory. The narrow CTA needs 64 KB; what is left of the SM's 256 KB
     * is L1, which the traceback's scattered reads and the cp.async staging live in: measured 10.04 ms (100 %) ->
     * 9.75 (60 %) -> 9.71 (35 %, i.e. the 100 KB configuration) on the target config, traceback steps of median reads
     * 190 -> 110 cycles (profiles/read_cycles_partd_r02_*). A wide CTA claims its SM by asking for so much dynamic
     * shared memory that neither a narrow CTA nor a second wide one fits beside it (wide_excl_bytes). */
    {
        cudaError_t e = cudaSuccess;
        int carve = 35;
        if (const char* e = getenv("ABEA_CARVEOUT")) carve = std::min(100, std::max(30, atoi(e)));
        c->wide_excl_bytes = carve < 60 ? (size_t)64 * 1024 : (size_t)160 * 1024;
        if (e == cudaSuccess) e = cudaFuncSetAttribute(abea_load_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(abea_load_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(abea_extract_means_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(abea_prepare_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        const void* fills[] = {(const void*)abea_fill_kernel<true, false>,      (const void*)abea_fill_kernel<false, false>,
                               (const void*)abea_fill_kernel<true, true>,       (const void*)abea_fill_kernel<false, true>,
                               (const void*)abea_fill_wide_kernel<true, false>, (const void*)abea_fill_wide_kernel<false, false>,
                               (const void*)abea_fill_wide_kernel<true, true>,  (const void*)abea_fill_wide_kernel<false, true>};
        for (int i = 0; i < 8; i++)
            if (e == cudaSuccess) e = cudaFuncSetAttribute(fills[i], cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e != cudaSuccess) {
            delete c;
            return ABEA_ERR_CUDA;
        }
    }
    c->cst.lp_skip = log(1e-10);
    c->cst.lp_trim = log(0.01);
    *out = c;
    return ABEA_OK;
}

void abea_destroy(abea_ctx_t* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    DevBuf* bufs[] = {&c->d_codes, &c->d_xcap, &c->d_xoff, &c->d_model, &c->d_seq, &c->d_events, &c->d_means, &c->d_reads, &c->d_kparams,
                      &c->d_trace, &c->d_pairs, &c->d_results, &c->d_queue, &c->d_flags, &c->d_npairs,
                      &c->d_ready, &c->d_items, &c->d_capptr, &c->d_dense_off, &c->d_sreads, &c->d_scalings, &c->d_maps, &c->d_sres,
                      &c->d_raw, &c->d_sum, &c->d_sumsq, &c->d_ts1, &c->d_ts2, &c->d_peaks, &c->d_evcap, &c->d_sigs, &c->d_sigorder,
                      &c->d_nev, &c->d_evptr, &c->d_evout, &c->d_chunks, &c->d_spec, &c->d_fix, &c->d_spec_cnt, &c->d_fix_cnt,
                      &c->d_sync, &c->d_spec_end, &c->d_raw16, &c->d_b5in, &c->d_b5out, &c->d_b5recs, &c->d_b5len, &c->d_b5status,
                      &c->d_b5hdr, &c->d_b5rawoff};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (c->h_results.p) cudaFreeHost(c->h_results.p);
    if (c->h_pairs.p) cudaFreeHost(c->h_pairs.p);
    if (c->h_reads.p) cudaFreeHost(c->h_reads.p);
    if (c->h_items.p) cudaFreeHost(c->h_items.p);
    for (HostBuf* hb : {&c->h_rseq, &c->h_rmeans, &c->h_rpairs, &c->h_rnp, &c->h_hostready, &c->h_rmeta, &c->h_b5, &c->h_codes, &c->h_fnp})
        if (hb->p) cudaFreeHost(hb->p);
    for (int i = 0; i < EV_COUNT; i++)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_meta) cudaEventDestroy(c->ev_meta);
    if (c->ev_loaded) cudaEventDestroy(c->ev_loaded);
    if (c->ev_load0) cudaEventDestroy(c->ev_load0);
    if (c->load_stream) cudaStreamDestroy(c->load_stream);
    if (c->wide_stream) cudaStreamDestroy(c->wide_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int abea_host_threads(abea_ctx_t* c, int threads) {
    if (!c) return ABEA_ERR_ARG;
    if (threads >= 0) c->host_threads = std::min(64, threads);
    return c->host_threads;
}

int abea_device_info(abea_ctx_t* c, int* sm_count, char* name) {
    if (!c) return ABEA_ERR_ARG;
    if (sm_count) *sm_count = c->sm_count;
    if (name) snprintf(name, 256, "%s", c->dev_name);
    return ABEA_OK;
}

int abea_set_model(abea_ctx_t* c, const abea_model_t* model, uint32_t kmer_size) {
    if (!c || !model || kmer_size < 1 || kmer_size > ABEA_MAX_KMER_SIZE) return fail(c, ABEA_ERR_ARG, "bad model");
    CU(cudaSetDevice(c->device));
    size_t n = (size_t)1 << (2 * kmer_size);
    if (dev_reserve(c, c->d_model, n * sizeof(abea_model_t))) return ABEA_ERR_CUDA;
    CU(cudaMemcpyAsync(c->d_model.p, model, n * sizeof(abea_model_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->kmer_size = kmer_size;
    c->have_model = true;
    return ABEA_OK;
}

/* ev_alias: device-side alias of the caller's pinned event array — batch->event_means when it is given, else
 * batch->events — (the batch is streamed in by abea_load_kernel while the fill runs), or NULL (the events go through
 * the copy engine before anything starts). rag: the call comes from abea_align_ragged, whose worker threads fill the
 * pinned array piece by piece in the order of the loader's work list while the loader runs (host_ready of abea_load_kernel). */
static int upload_impl(abea_ctx_t* c, const abea_batch_t* b, const void* ev_alias, abea_timing_t* timing,
                       bool rag = false) {
    if (!c || !b || b->n_reads < 0) return fail(c, ABEA_ERR_ARG, "bad batch");
    if (!c->have_model) return fail(c, ABEA_ERR_NOMODEL, "abea_set_model has not been called");
    CU(cudaSetDevice(c->device));
    double t0 = now_ms();
    c->uploaded = false;
    c->ran = false;
    c->n_batch_reads = b->n_reads;
    c->reads.clear();
    c->reads.reserve(b->n_reads);

    /* batch->events == NULL: the event tables are the ones the last abea_getevents left on the device (same reads,
     * same order); they are used where they lie (capacity layout), nothing is copied */
    const bool have_means = (b->event_means != nullptr);
    const bool dev_events = (b->events == nullptr) && !have_means && b->n_reads > 0;
    if (dev_events) {
        if (!c->events_ready || (int32_t)c->nev.size() != b->n_reads)
            return fail(c, ABEA_ERR_STATE, "batch without events, but no matching abea_getevents result on the device");
        for (int32_t i = 0; i < b->n_reads; i++)
            if (b->n_events[i] != std::max(c->nev[i], 0))
                return fail(c, ABEA_ERR_ARG, "n_events[%d] = %d does not match abea_getevents (%d)", i, b->n_events[i], c->nev[i]);
        ev_alias = nullptr;
    }
    /* total sizes of the caller's flat arrays */
    int64_t seq_bytes = 0, n_ev_total = 0;
    for (int32_t i = 0; i < b->n_reads; i++) {
        int64_t se = b->seq_ptr[i] + b->read_len[i] + 1;
        if (se > seq_bytes) seq_bytes = se;
        int64_t ee = dev_events ? 0 : b->event_ptr[i] + b->n_events[i];
        if (ee > n_ev_total) n_ev_total = ee;
    }

    /* the bases do not depend on the schedule: their copy is enqueued first and crosses PCIe while this thread builds the
     * descriptors and the schedule (0.4 ms per 4096 reads) */
    CU(cudaEventRecord(c->ev[EV_H2D0], c->stream));
    if (dev_reserve(c, c->d_seq, (size_t)seq_bytes + 16)) return ABEA_ERR_CUDA;
    if (seq_bytes) CU(cudaMemcpyAsync(c->d_seq.p, b->seq, (size_t)seq_bytes, cudaMemcpyHostToDevice, c->stream));

    /* canonical output layout: read i owns pairs [cap_ptr[i], cap_ptr[i] + E_i + L_i) (reference src/f5c.c:724-726) */
    c->cap_ptr.assign((size_t)b->n_reads + 1, 0);
    for (int32_t i = 0; i < b->n_reads; i++)
        c->cap_ptr[i + 1] = c->cap_ptr[i] + (int64_t)b->n_events[i] + (int64_t)b->read_len[i];

    /* eligibility: align_single's filter (reference src/f5c.c:811-830). Reads with no events or fewer bases than
     * k are undefined in the reference; they get 0 pairs here. */
    const int32_t k = (int32_t)c->kmer_size;
    const double exp_lp_skip = exp(c->cst.lp_skip); /* the same double for every read (src/align.c:215) */
    for (int32_t i = 0; i < b->n_reads; i++) {
        const int32_t E = b->n_events[i], L = b->read_len[i];
        const bool good = b->good ? (b->good[i] != 0) : true;
        if (!good || E < 1 || L < k) continue;
        if ((int64_t)E + (int64_t)L + 2 >= (int64_t)0x7fffffff) continue; /* band counters are 32-bit */
        if (!((float)(size_t)E / (float)L < ABEA_AVG_EVENTS_PER_KMER_MAX)) continue;
        abea_read_t r;
        memset(&r, 0, sizeof(r));
        r.seq_off = b->seq_ptr[i];
        r.ev_off = dev_events ? c->sigs[i].cap_off : b->event_ptr[i];
        r.n_events = E;
        r.n_kmers = L - k + 1;
        r.pair_cap = E + L;
        r.orig_index = i;
        r.scale = b->scalings ? b->scalings[i].scale : 0.f;
        r.shift = b->scalings ? b->scalings[i].shift : 0.f;
        /* per-read transition penalties in host double (reference src/align.c:207-215) */
        double events_per_kmer = (double)(size_t)E / (size_t)r.n_kmers;
        double p_stay = 1 - (1 / (events_per_kmer + 1));
        r.lp_stay = log(p_stay);
        r.lp_step = log(1.0 - exp_lp_skip - exp(r.lp_stay));
        c->reads.push_back(r);
    }
    /* longest-first schedule: a read is a serial chain of NB = E+K+2 bands. Sorted through 12-byte keys, not by
     * moving the 96-byte descriptors around (this is on the critical path of abea_align_batch) */
    {
        struct key_t { int64_t nb; int32_t idx; };
        std::vector<key_t> keys(c->reads.size());
        for (size_t j = 0; j < keys.size(); j++)
            keys[j] = key_t{(int64_t)c->reads[j].n_events + c->reads[j].n_kmers, (int32_t)j};
        std::stable_sort(keys.begin(), keys.end(), [](const key_t& x, const key_t& y) { return x.nb > y.nb; });
        std::vector<abea_read_t> sorted(c->reads.size());
        for (size_t j = 0; j < keys.size(); j++) sorted[j] = c->reads[(size_t)keys[j].idx];
        c->reads.swap(sorted);
    }
    int64_t kp = 0, tw = 0, pc = 0, nb = 0, ne = 0;
    for (abea_read_t& r : c->reads) {
        const int64_t NB = (int64_t)r.n_events + r.n_kmers + 2;
        r.kp_off = kp;
        r.evs_off = ne;
        r.trace_off = tw;
        r.pair_off = c->cap_ptr[r.orig_index];
        kp += r.n_kmers;
        tw += ((NB + 3) / 4) * ABEA_TRACE_GROUP_WORDS;
        pc += r.pair_cap;
        nb += NB;
        ne += r.n_events;
    }
    /* The longest reads go to the wide kernel (one CTA of 4 warps per read). Measured on B200 (profiles/): a narrow
     * warp alone on its SM sub-partition needs ~650 cycles per band, a wide CTA alone on its SM ~400 (cyc_long /
     * cyc_wide of the scheduler's model, re-measured per batch), and when SMs are shared the wide form has no
     * advantage. So a read is made wide only when it would outlast the whole batch even as a lone narrow warp — it is
     * going to run (nearly) alone at the end anyway, and then the wide form finishes it sooner. The target config
     * (log-normal sigma 0.5) sends 6-8 reads wide; cfg3 (sigma 1.0) and cfg4 fill the cap of SMs / 4. */
    c->n_wide = 0;
    if (c->wide_mode && !c->reads.empty()) {
        const double thr = wide_threshold(c, nb);
        /* at most one wide CTA per SM: co-resident wide CTAs lose their advantage */
        /* a batch with fewer reads than SMs may run entirely wide; a larger one gives at most wide_cap SMs away */
        const int32_t cap = ((int32_t)c->reads.size() <= c->sm_count) ? c->sm_count : c->wide_cap;
        while (c->n_wide < (int32_t)c->reads.size() && c->n_wide < cap &&
               (double)c->reads[c->n_wide].n_events + c->reads[c->n_wide].n_kmers + 2 > thr)
            c->n_wide++;
        /* The two SMs of a TPC are not independent: a wide CTA whose partner SM runs a narrow CTA needs 550-620 cycles
         * per band instead of 396 (measured, profiles/sweep_tpc_pairing_r01.txt; a wide or an idle partner costs
         * nothing). The hardware fills TPCs pairwise, so an EVEN number of wide CTAs beside an even narrow grid keeps
         * every wide CTA next to another wide CTA: an odd count takes the next-longest read along. */
        if ((c->n_wide & 1) && c->n_wide < (int32_t)c->reads.size() && (int32_t)c->reads.size() > c->sm_count) c->n_wide++;
    }
    c->total_kmers = kp;
    c->total_trace_words = tw;
    c->total_pair_cap = c->cap_ptr[b->n_reads];
    (void)pc;
    c->total_bands = nb;
    c->total_events = ne;
    const size_t n_sched = c->reads.size();
    /* descriptors for the stages either side of the alignment: every read, caller's order */
    c->sreads.assign((size_t)b->n_reads, abea_sread_t());
    c->total_map = 0;
    for (int32_t i = 0; i < b->n_reads; i++) {
        abea_sread_t& sr = c->sreads[i];
        const int32_t E = b->n_events[i], L = b->read_len[i];
        sr.seq_off = b->seq_ptr[i];
        sr.ev_off = dev_events ? c->sigs[i].cap_off : b->event_ptr[i];
        sr.map_off = c->total_map;
        sr.pair_off = c->cap_ptr[i];
        sr.n_events = E;
        sr.read_len = L;
        sr.sched = -1;
        sr.usable = ((b->good ? b->good[i] != 0 : true) && E >= 1 && L >= k) ? 1 : 0;
        c->total_map += std::max(L - k + 1, 0);
    }
    for (size_t j = 0; j < n_sched; j++) c->sreads[c->reads[j].orig_index].sched = (int32_t)j;
    if (b->scalings) c->in_scalings.assign(b->scalings, b->scalings + b->n_reads);
    else c->in_scalings.clear();
    c->need_scalings = (b->scalings == nullptr);
    c->sreads_on_device = false;
    c->scalings_on_device = false;
    c->scaled = false;
    double t1 = now_ms();

    /* the alignment reads event means only (reference src/align.cu:415): they live in d_means, indexed like the source
     * (the caller's event_ptr space, or the capacity layout of abea_getevents). An AoS table that goes through the copy
     * engine is staged in d_events and its means are extracted on the device. */
    const int64_t n_means = dev_events ? c->evcap_total : n_ev_total;
    if (c->need_scalings || n_sched == 0) ev_alias = nullptr; /* the estimate needs a read's events before its fill */
    const bool stage_aos = !dev_events && !have_means && !ev_alias;
    if (stage_aos && dev_reserve(c, c->d_events, (size_t)n_ev_total * sizeof(abea_event_t) + 16)) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_means, (size_t)n_means * sizeof(float) + 64)) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_reads, (n_sched + 1) * sizeof(abea_read_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_kparams, (size_t)(kp + 1) * sizeof(float4))) return ABEA_ERR_CUDA;
    /* + one traceback chunk of slack: the prefetcher copies whole 16-group chunks */
    if (dev_reserve(c, c->d_trace, (size_t)(tw + (ABEA_TB_CHUNK_GROUPS + 1) * ABEA_TRACE_GROUP_WORDS) * sizeof(uint32_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_pairs, (size_t)(c->total_pair_cap + 1) * sizeof(abea_pair_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_npairs, ((size_t)b->n_reads + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_results, (n_sched + 1) * sizeof(abea_result_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_queue, 64)) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_flags, (n_sched + 1) * sizeof(uint32_t))) return ABEA_ERR_CUDA;

    c->load_aos = !have_means;
    c->event_bytes = n_ev_total * (int64_t)(have_means ? sizeof(float) : sizeof(abea_event_t));
    c->load_piece_cur = c->load_piece > 0 ? c->load_piece : (have_means ? ABEA_LOAD_PIECE_BYTES_MEANS : ABEA_LOAD_PIECE_BYTES);
    c->means_from_evcap = dev_events;
    c->means_reversed = false;
    c->streaming = false;
    c->results_on_device = false;
    double t2 = t1;
    c->prepared = false;
    if (n_sched) { /* descriptors through pinned staging: a pageable source would make the copy synchronous */
        if (host_reserve(c, c->h_reads, n_sched * sizeof(abea_read_t))) return ABEA_ERR_CUDA;
        memcpy(c->h_reads.p, c->reads.data(), n_sched * sizeof(abea_read_t));
        CU(cudaMemcpyAsync(c->d_reads.p, c->h_reads.p, n_sched * sizeof(abea_read_t), cudaMemcpyHostToDevice, c->stream));
    }
    if (ev_alias && n_sched) {
        /* Everything the k-mer parameter kernel needs goes first, and the kernel with it: the host then works out
         * the loader's order while the GPU is busy with that. */
        if (dev_reserve(c, c->d_ready, ABEA_READY_WORDS * (n_sched + 1) * sizeof(uint32_t))) return ABEA_ERR_CUDA;
        CU(cudaMemsetAsync(c->d_queue.p, 0, 64, c->stream));
        CU(cudaMemsetAsync(c->d_flags.p, 0x01, n_sched * sizeof(uint32_t), c->stream));
        CU(cudaMemsetAsync(c->d_ready.p, 0, ABEA_READY_WORDS * n_sched * sizeof(uint32_t), c->stream));
        CU(cudaEventRecord(c->ev_meta, c->stream));
        CU(cudaEventRecord(c->ev[EV_K0], c->stream));
        launch_prepare(c, 0);
        CU(cudaEventRecord(c->ev[EV_K1], c->stream));
        c->prepared = true;
        /* Work list of the loader: every (read, piece) in the order the fill is expected to NEED it. The fill starts
         * a read when its first piece has landed and then chases the loader piece by piece, so what matters is when
         * each piece is first touched. That is predicted by replaying the schedule with the scheduler's model of the
         * kernels (build_load_order) — half a millisecond of host work. The FIRST pieces of the reads the first wave
         * of warps starts with need no prediction: they go out at once, with a loader launch of their own, and the
         * rest of the list is built and shipped after the fill kernels have been launched (finish_load_order). */
        const double t_enq = now_ms();
        {
            const int wpc = c->fill_warps_per_cta;
            const int nw = c->n_wide;
            const int wide_sms = nw > 0 ? std::min(c->sm_count, nw + c->sm_reserve) : 0;
            int64_t blocks = std::max<int64_t>(1, std::min<int64_t>((int64_t)(c->sm_count - wide_sms) * c->fill_ctas_per_sm,
                                                                    ((int64_t)n_sched - nw + wpc - 1) / wpc));
            c->n_first_items = (int32_t)std::min<int64_t>((int64_t)n_sched, (int64_t)nw + blocks * wpc);
        }
        const size_t max_items = (size_t)n_sched * 2 + (size_t)(c->event_bytes / c->load_piece_cur) + 8;
        if (c->items.size() < max_items) c->items.resize(max_items);
        for (int32_t r = 0; r < c->n_first_items; r++) c->items[(size_t)r] = abea_load_item_t{r, 0};
        c->n_items = c->n_first_items;
        if (dev_reserve(c, c->d_items, max_items * sizeof(abea_load_item_t))) return ABEA_ERR_CUDA;
        if (host_reserve(c, c->h_items, max_items * sizeof(abea_load_item_t))) return ABEA_ERR_CUDA;
        memcpy(c->h_items.p, c->items.data(), (size_t)c->n_first_items * sizeof(abea_load_item_t));
        const uint32_t* host_ready = nullptr;
        c->rag_mode = rag;
        if (rag) { /* one flag per work item, set by the packer threads; they start as soon as a part of the list exists */
            if (host_reserve(c, c->h_hostready, (max_items + 1) * sizeof(uint32_t))) return ABEA_ERR_CUDA;
            memset(c->h_hostready.p, 0, (max_items + 1) * sizeof(uint32_t));
            host_ready = (const uint32_t*)mapped_alias(c->h_hostready.p);
            if (!host_ready) return fail(c, ABEA_ERR_CUDA, "pinned staging is not mapped into the device address space");
            c->rag_items_final.store(0, std::memory_order_relaxed);
            c->rag_items_ready.store(c->n_first_items, std::memory_order_release);
        }
        t2 = now_ms();
        if (getenv("ABEA_TIME_PACK"))
            fprintf(stderr, "[abea pack] descriptors+sort %.3f ms, reserve+enqueue+prepare %.3f ms, first wave list %.3f ms (%d items)\n",
                    t1 - t0, t_enq - t1, t2 - t_enq, c->n_first_items);
        /* the loader's stream waits for the descriptors and the cleared counters, not for the k-mer kernel */
        CU(cudaStreamWaitEvent(c->load_stream, c->ev_meta, 0));
        CU(cudaMemcpyAsync(c->d_items.p, c->h_items.p, (size_t)c->n_first_items * sizeof(abea_load_item_t),
                           cudaMemcpyHostToDevice, c->load_stream));
        CU(cudaEventRecord(c->ev_load0, c->load_stream));
        c->ev_alias_cur = ev_alias;
        int rc_l = launch_loader(c, 0, c->n_first_items, 12);
        if (rc_l) return rc_l;
        c->pending_order = true;
#ifdef ABEA_SIMT_EMU /* the emulator runs a kernel to completion at its launch: the whole list must be in before the fill */
        {
            const int rc_f = finish_load_order(c);
            if (rc_f) return rc_f;
        }
#endif
        c->streaming = true;
    } else if (have_means) {
        if (n_ev_total)
            CU(cudaMemcpyAsync(c->d_means.p, b->event_means, (size_t)n_ev_total * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    } else {
        const abea_event_t* src_tab = dev_events ? (const abea_event_t*)c->d_evcap.p : (const abea_event_t*)c->d_events.p;
        if (!dev_events && n_ev_total)
            CU(cudaMemcpyAsync(c->d_events.p, b->events, (size_t)n_ev_total * sizeof(abea_event_t),
                               cudaMemcpyHostToDevice, c->stream));
        if (n_means > 0) {
            const int blocks = (int)std::min<int64_t>((n_means + 255) / 256, (int64_t)c->sm_count * 16);
            ABEA_LAUNCH(abea_extract_means_kernel, blocks, 256, c->stream, src_tab, (float*)c->d_means.p, n_means);
        }
    }
    CU(cudaEventRecord(c->ev[EV_H2D1], c->stream));
    if (!c->streaming) CU(cudaStreamSynchronize(c->stream));
    c->uploaded = true;

    c->last = abea_timing_t();
    c->last.pack_ms = (c->streaming ? t2 : t1) - t0;
    c->last.h2d_ms = c->streaming ? 0.f : ev_ms(c, EV_H2D0, EV_H2D1); /* streamed: filled in after the run */
    c->last.h2d_bytes = seq_bytes + (dev_events ? 0 : c->event_bytes) + (int64_t)(n_sched * sizeof(abea_read_t));
    c->last.n_scheduled = (int32_t)n_sched;
    c->last.n_wide = c->n_wide;
    c->last.streamed = c->streaming ? 1 : 0;
    c->last.n_bands = nb;
    c->last.n_events = ne;
    if (timing) *timing = c->last;
    return ABEA_OK;
}

int abea_upload_batch(abea_ctx_t* c, const abea_batch_t* b, abea_timing_t* timing) {
    return upload_impl(c, b, nullptr, timing);
}

/* fin_pairs / fin_np: device-side aliases of the caller's pinned output buffers (canonical layout), or NULL: the
 * lists stay in d_pairs / d_npairs for abea_download. */
static int run_impl(abea_ctx_t* c, abea_pair_t* fin_pairs, int32_t* fin_np, abea_timing_t* timing,
                    abea_code_t* fin_codes = nullptr) {
    if (!c) return ABEA_ERR_ARG;
    if (!c->uploaded) return fail(c, ABEA_ERR_STATE, "abea_run before abea_upload_batch");
    if (c->need_scalings)
        return fail(c, ABEA_ERR_STATE, "the batch has no scalings: call abea_estimate_scalings before abea_run");
    CU(cudaSetDevice(c->device));
    c->scaled = false;
    const int32_t n = (int32_t)c->reads.size();
    int launches = 0;
    const bool streaming = c->streaming; /* set by upload_impl: queue, flags and ready counters are already armed */
    abea_stream_t io;
    io.ready = streaming ? (const uint32_t*)c->d_ready.p : nullptr;
    io.pairs_final = fin_pairs;
    io.n_pairs_final = fin_np;
    io.codes_final = fin_codes;
    c->codes_current = false;
    io.n_pairs_dev = (int32_t*)c->d_npairs.p;
    io.stalled = (uint32_t*)c->d_queue.p + 15; /* zeroed with the queue */
    io.tb_mode = c->tb_mode;
    io.tb_margin = c->tb_margin;
    if (!streaming) CU(cudaMemsetAsync(c->d_queue.p, 0, 64, c->stream));
    CU(cudaMemsetAsync(c->d_npairs.p, 0, ((size_t)c->n_batch_reads + 1) * sizeof(int32_t), c->stream));
    if (!c->prepared) CU(cudaEventRecord(c->ev[EV_K0], c->stream));
    if (n > 0) {
        int32_t* queue = (int32_t*)c->d_queue.p;
        if (!c->prepared) {
            /* every read starts as "fast"; abea_prepare_kernel clears the flag of reads with out-of-range inputs
             * (a streamed batch: launched by the upload; the loader checks the event means as they pass through it) */
            CU(cudaMemsetAsync(c->d_flags.p, 0x01, (size_t)n * sizeof(uint32_t), c->stream));
            launch_prepare(c, c->total_events);
            CU(cudaEventRecord(c->ev[EV_K1], c->stream));
        }
        launches++;
        {
            /* The longest n_wide reads: one CTA of 4 warps each (wide kernel) on a second stream, beside the persistent
             * narrow CTAs (one CTA of 12 warps per remaining SM, each warp pulls reads longest-first). The FAST
             * instantiations take the reads whose inputs passed validation, the EXACT ones the rest (normally none). */
            const int32_t nw = c->n_wide;
            if (nw > 0) {
                CU(cudaEventRecord(c->ev_fork, c->stream));
                CU(cudaStreamWaitEvent(c->wide_stream, c->ev_fork, 0));
                int wblocks = std::min(c->sm_count, (int)nw);
                /* A wide CTA is only faster than a lone narrow warp when it has its SM to itself (measured: 400
                 * cycles/band alone, 800 beside a narrow CTA). It therefore asks for so much dynamic shared memory
                 * that no narrow CTA fits on the same SM. Tiny batches (every read wide) do not need the exclusion. */
                const size_t excl = (n > nw) ? c->wide_excl_bytes : 0;
                /* the STREAM instantiations chase the loader (abea_wait_landed_events); the others assume a resident batch */
                auto wide_fast = streaming ? abea_fill_wide_kernel<true, true> : abea_fill_wide_kernel<true, false>;
                auto wide_exact = streaming ? abea_fill_wide_kernel<false, true> : abea_fill_wide_kernel<false, false>;
                CU(cudaFuncSetAttribute((const void*)wide_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)excl));
                CU(cudaFuncSetAttribute((const void*)wide_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)excl));
                ABEA_LAUNCH_SMEM(wide_fast, wblocks, 128, excl, c->wide_stream,
                    (const abea_read_t*)c->d_reads.p, nw, (const float*)c->d_means.p, (const float4*)c->d_kparams.p,
                    (const uint32_t*)c->d_flags.p, (uint32_t*)c->d_trace.p, (abea_pair_t*)c->d_pairs.p,
                    (abea_result_t*)c->d_results.p, io, c->cst, queue + 6);
                ABEA_LAUNCH_SMEM(wide_exact, wblocks, 128, excl, c->wide_stream,
                    (const abea_read_t*)c->d_reads.p, nw, (const float*)c->d_means.p, (const float4*)c->d_kparams.p,
                    (const uint32_t*)c->d_flags.p, (uint32_t*)c->d_trace.p, (abea_pair_t*)c->d_pairs.p,
                    (abea_result_t*)c->d_results.p, io, c->cst, queue + 7);
                CU(cudaEventRecord(c->ev_join, c->wide_stream));
                launches += 2;
            }
            if (n > nw) {
                const int wpc = c->fill_warps_per_cta;
                /* SMs taken by SM-exclusive wide CTAs are left out of the persistent narrow grid, so that both kernels
                 * are resident from the start whatever order the hardware dispatches them in */
                const int wide_sms = (nw > 0) ? std::min(c->sm_count, (int)nw + c->sm_reserve) : 0;
                int blocks = std::min((c->sm_count - wide_sms) * c->fill_ctas_per_sm, (n - nw + wpc - 1) / wpc);
                if (nw > 0 && blocks > 1) blocks &= ~1; /* whole TPCs (see upload_impl) */
                if (blocks < 1) blocks = 1;
                const size_t smem = (size_t)wpc * (sizeof(abea_fill_smem_t) +
                                                   sizeof(uint32_t) * ABEA_TB_RING_GROUPS * ABEA_TRACE_GROUP_WORDS);
                auto fill_fast = streaming ? abea_fill_kernel<true, true> : abea_fill_kernel<true, false>;
                auto fill_exact = streaming ? abea_fill_kernel<false, true> : abea_fill_kernel<false, false>;
                CU(cudaFuncSetAttribute((const void*)fill_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CU(cudaFuncSetAttribute((const void*)fill_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                /* a narrow read is "long" when, sharing its sub-partition (~670 cycles/band at two warps), it would take more than
                 * long_alpha of the time the whole batch needs at full throughput */
                const int32_t long_thr = long_threshold(c);
                ABEA_LAUNCH_SMEM(fill_fast, blocks, 32 * wpc, smem, c->stream,
                    (const abea_read_t*)c->d_reads.p, n, (const float*)c->d_means.p, (const float4*)c->d_kparams.p,
                    (const uint32_t*)c->d_flags.p, (uint32_t*)c->d_trace.p, (abea_pair_t*)c->d_pairs.p,
                    (abea_result_t*)c->d_results.p, io, c->cst, queue, nw, long_thr, c->sched_policy);
                ABEA_LAUNCH_SMEM(fill_exact, blocks, 32 * wpc, smem, c->stream,
                    (const abea_read_t*)c->d_reads.p, n, (const float*)c->d_means.p, (const float4*)c->d_kparams.p,
                    (const uint32_t*)c->d_flags.p, (uint32_t*)c->d_trace.p, (abea_pair_t*)c->d_pairs.p,
                    (abea_result_t*)c->d_results.p, io, c->cst, queue + 8, nw, long_thr, c->sched_policy);
                launches += 2;
            }
            if (nw > 0) CU(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        }
        CU(cudaEventRecord(c->ev[EV_K2], c->stream)); /* traceback + QC are fused into the fill kernels */
    } else {
        if (!c->prepared) CU(cudaEventRecord(c->ev[EV_K1], c->stream));
        CU(cudaEventRecord(c->ev[EV_K2], c->stream));
    }
    if (streaming) {
        if (c->pending_order) { /* the fill kernels are launched and working on the first wave: now the rest of the list */
            const int rc = finish_load_order(c);
            if (rc) return rc;
        }
        CU(cudaStreamWaitEvent(c->stream, c->ev_loaded, 0));
        launches += 2;
    }
    CU(cudaEventRecord(c->ev[EV_K3], c->stream));
    uint32_t stalled = 0;
    if (streaming)
        CU(cudaMemcpyAsync(&stalled, (uint32_t*)c->d_queue.p + 15, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    if (stalled) {
        c->streaming = false;
        c->prepared = false;
        c->uploaded = false;
        return fail(c, ABEA_ERR_CUDA, "streaming stalled: events did not arrive from the caller's pinned buffer");
    }
    c->ran = true;
    c->prepared = false;
    c->streaming = false; /* everything has landed: a further abea_run works on the resident copy */
    if (c->calib_mode && !streaming && (c->calib_runs++ & 7) == 0) { /* the first resident run and every eighth after it */
        const int rc = calibrate(c, long_threshold(c));
        if (rc) return rc;
    }
    c->results_on_device = true; /* a streamed-out run keeps the device copy too */
    if (streaming) {
        c->last.h2d_ms = ev_ms(c, EV_H2D0, EV_H2D1);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_load0, c->ev_loaded) == cudaSuccess) c->last.load_ms = ms;
    }
    if (fin_pairs) c->last.streamed |= 2;
    if (fin_codes) c->last.streamed |= 4;
    c->last.kmer_ms = ev_ms(c, EV_K0, EV_K1);
    c->last.fill_ms = ev_ms(c, EV_K1, EV_K2);
    c->last.trace_ms = ev_ms(c, EV_K2, EV_K3);
    c->last.kernel_ms = ev_ms(c, EV_K0, EV_K3);
    c->last.kernel_launches = launches;
    if (timing) *timing = c->last;
    return ABEA_OK;
}

int abea_run(abea_ctx_t* c, abea_timing_t* timing) { return run_impl(c, nullptr, nullptr, timing); }

int abea_download(abea_ctx_t* c, abea_pair_t* pairs, const int64_t* pair_ptr, int32_t* n_pairs,
                  abea_timing_t* timing) {
    if (!c || !n_pairs || (!pairs && c->total_pair_cap > 0) || !pair_ptr) return fail(c, ABEA_ERR_ARG, "bad output");
    if (!c->ran) return fail(c, ABEA_ERR_STATE, "abea_download before abea_run");
    if (!c->results_on_device) return fail(c, ABEA_ERR_STATE, "no results on the device yet");
    CU(cudaSetDevice(c->device));
    const int32_t nb = c->n_batch_reads;
    /* The device holds the pairs in the canonical capacity layout (read i at prefix-sum(E+L)); when the caller's
     * pair_ptr is that layout — it is what align_cuda's packer and ReadBatch.pair_ptr() produce — both results are
     * copied straight into the caller's buffers. Any other layout goes through pinned staging and a scatter. */
    bool canonical = true;
    for (int32_t i = 0; i < nb && canonical; i++) canonical = (pair_ptr[i] == c->cap_ptr[i]);
    double unpack_ms = 0.0;
    int64_t d2h = 0;
    CU(cudaEventRecord(c->ev[EV_D2H0], c->stream));
    if (nb > 0) {
        CU(cudaMemcpyAsync(n_pairs, c->d_npairs.p, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        d2h += (int64_t)nb * (int64_t)sizeof(int32_t);
    }
    if (canonical) {
        if (c->total_pair_cap > 0) {
            CU(cudaMemcpyAsync(pairs, c->d_pairs.p, (size_t)c->total_pair_cap * sizeof(abea_pair_t),
                               cudaMemcpyDeviceToHost, c->stream));
            d2h += c->total_pair_cap * (int64_t)sizeof(abea_pair_t);
        }
        CU(cudaEventRecord(c->ev[EV_D2H1], c->stream));
        CU(cudaStreamSynchronize(c->stream));
    } else {
        if (host_reserve(c, c->h_pairs, (size_t)(c->total_pair_cap + 1) * sizeof(abea_pair_t))) return ABEA_ERR_CUDA;
        if (c->total_pair_cap > 0) {
            CU(cudaMemcpyAsync(c->h_pairs.p, c->d_pairs.p, (size_t)c->total_pair_cap * sizeof(abea_pair_t),
                               cudaMemcpyDeviceToHost, c->stream));
            d2h += c->total_pair_cap * (int64_t)sizeof(abea_pair_t);
        }
        CU(cudaEventRecord(c->ev[EV_D2H1], c->stream));
        CU(cudaStreamSynchronize(c->stream));
        double t0 = now_ms();
        const abea_pair_t* hp = (const abea_pair_t*)c->h_pairs.p;
        for (int32_t i = 0; i < nb; i++)
            if (n_pairs[i] > 0)
                memcpy(pairs + pair_ptr[i], hp + c->cap_ptr[i], (size_t)n_pairs[i] * sizeof(abea_pair_t));
        unpack_ms = now_ms() - t0;
    }
    c->last.d2h_ms = ev_ms(c, EV_D2H0, EV_D2H1);
    c->last.unpack_ms = unpack_ms;
    c->last.d2h_bytes = d2h;
    if (timing) *timing = c->last;
    return ABEA_OK;
}

int abea_read_stats(abea_ctx_t* c, double* sum_emission, int32_t* n_aligned, int32_t* end_event, int32_t* max_gap) {
    if (!c) return ABEA_ERR_ARG;
    if (!c->ran) return fail(c, ABEA_ERR_STATE, "abea_read_stats before abea_run");
    CU(cudaSetDevice(c->device));
    const size_t n = c->reads.size();
    std::vector<abea_result_t> res(n);
    if (n) CU(cudaMemcpy(res.data(), c->d_results.p, n * sizeof(abea_result_t), cudaMemcpyDeviceToHost));
    for (int32_t i = 0; i < c->n_batch_reads; i++) {
        if (sum_emission) sum_emission[i] = 0;
        if (n_aligned) n_aligned[i] = 0;
        if (end_event) end_event[i] = 0;
        if (max_gap) max_gap[i] = 0;
    }
    for (size_t j = 0; j < n; j++) {
        int32_t i = c->reads[j].orig_index;
        if (sum_emission) sum_emission[i] = res[j].sum_emission;
        if (n_aligned) n_aligned[i] = res[j].n_aligned;
        if (end_event) end_event[i] = res[j].end_event;
        if (max_gap) max_gap[i] = res[j].max_gap;
    }
    return ABEA_OK;
}

int abea_read_cycles(abea_ctx_t* c, int64_t* fill_cycles, int64_t* trace_cycles, int32_t* wide) {
    if (!c) return ABEA_ERR_ARG;
    if (!c->ran) return fail(c, ABEA_ERR_STATE, "abea_read_cycles before abea_run");
    CU(cudaSetDevice(c->device));
    const size_t n = c->reads.size();
    std::vector<abea_result_t> res(n);
    if (n) CU(cudaMemcpy(res.data(), c->d_results.p, n * sizeof(abea_result_t), cudaMemcpyDeviceToHost));
    for (int32_t i = 0; i < c->n_batch_reads; i++) {
        if (fill_cycles) fill_cycles[i] = 0;
        if (trace_cycles) trace_cycles[i] = 0;
        if (wide) wide[i] = 0;
    }
    for (size_t j = 0; j < n; j++) {
        int32_t i = c->reads[j].orig_index;
        if (fill_cycles) fill_cycles[i] = res[j].fill_cycles;
        if (trace_cycles) trace_cycles[i] = res[j].trace_cycles;
        if (wide) wide[i] = res[j].wide;
    }
    return ABEA_OK;
}

int abea_read_respec(abea_ctx_t* c, int32_t* respec) {
    if (!c || !respec) return ABEA_ERR_ARG;
    if (!c->ran) return fail(c, ABEA_ERR_STATE, "abea_read_respec before abea_run");
    CU(cudaSetDevice(c->device));
    const size_t n = c->reads.size();
    std::vector<abea_result_t> res(n);
    if (n) CU(cudaMemcpy(res.data(), c->d_results.p, n * sizeof(abea_result_t), cudaMemcpyDeviceToHost));
    for (int32_t i = 0; i < c->n_batch_reads; i++) respec[i] = 0;
    for (size_t j = 0; j < n; j++) respec[c->reads[j].orig_index] = res[j].respec;
    return ABEA_OK;
}

int abea_read_starts(abea_ctx_t* c, int32_t* start_us) {
    if (!c || !start_us) return ABEA_ERR_ARG;
    if (!c->ran) return fail(c, ABEA_ERR_STATE, "abea_read_starts before abea_run");
    CU(cudaSetDevice(c->device));
    const size_t n = c->reads.size();
    std::vector<abea_result_t> res(n);
    if (n) CU(cudaMemcpy(res.data(), c->d_results.p, n * sizeof(abea_result_t), cudaMemcpyDeviceToHost));
    for (int32_t i = 0; i < c->n_batch_reads; i++) start_us[i] = -1;
    for (size_t j = 0; j < n; j++) start_us[c->reads[j].orig_index] = res[j].start_us;
    return ABEA_OK;
}

int abea_device_results(abea_ctx_t* c, const abea_pair_t** d_pairs, const int32_t** d_n_pairs,
                        int64_t* total_pairs_capacity, int32_t* n_reads) {
    if (!c) return ABEA_ERR_ARG;
    if (!c->ran) return fail(c, ABEA_ERR_STATE, "abea_device_results before abea_run");
    if (!c->results_on_device) return fail(c, ABEA_ERR_STATE, "no results on the device yet");
    if (d_pairs) *d_pairs = (const abea_pair_t*)c->d_pairs.p;
    if (d_n_pairs) *d_n_pairs = (const int32_t*)c->d_npairs.p;
    if (total_pairs_capacity) *total_pairs_capacity = c->total_pair_cap;
    if (n_reads) *n_reads = c->n_batch_reads;
    return ABEA_OK;
}

int abea_compact_results(abea_ctx_t* c, abea_pair_t* d_dst, int64_t dst_capacity, int64_t* total_pairs) {
    if (!c || !total_pairs) return ABEA_ERR_ARG;
    if (!c->ran || !c->results_on_device) return fail(c, ABEA_ERR_STATE, "abea_compact_results before abea_run");
    CU(cudaSetDevice(c->device));
    const int32_t n = c->n_batch_reads;
    *total_pairs = 0;
    if (n == 0) return ABEA_OK;
    if (dev_reserve(c, c->d_capptr, ((size_t)n + 1) * sizeof(int64_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_dense_off, ((size_t)n + 1) * sizeof(int64_t))) return ABEA_ERR_CUDA;
    if (host_reserve(c, c->h_items, ((size_t)n + 2) * sizeof(int64_t))) return ABEA_ERR_CUDA; /* pinned scratch */
    memcpy(c->h_items.p, c->cap_ptr.data(), ((size_t)n + 1) * sizeof(int64_t));
    CU(cudaMemcpyAsync(c->d_capptr.p, c->h_items.p, ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    ABEA_LAUNCH(abea_pair_offsets_kernel, 1, ABEA_SCAN_THREADS, c->stream, (const int32_t*)c->d_npairs.p, n,
                (int64_t*)c->d_dense_off.p);
    int64_t* h_total = (int64_t*)c->h_items.p + (n + 1);
    CU(cudaMemcpyAsync(h_total, (const int64_t*)c->d_dense_off.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *total_pairs = *h_total;
    if (*h_total > dst_capacity) return fail(c, ABEA_ERR_ARG, "dense buffer holds %lld pairs, %lld needed", (long long)dst_capacity, (long long)*h_total);
    if (*h_total > 0) {
        if (!d_dst) return fail(c, ABEA_ERR_ARG, "no destination");
        const int blocks = std::max(1, std::min((n + 7) / 8, c->sm_count * 8));
        ABEA_LAUNCH(abea_compact_pairs_kernel, blocks, 256, c->stream, (const abea_pair_t*)c->d_pairs.p,
                    (const int64_t*)c->d_capptr.p, (const int32_t*)c->d_npairs.p, (const int64_t*)c->d_dense_off.p, n, d_dst);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
    }
    return ABEA_OK;
}

int abea_device_codes(abea_ctx_t* c, const abea_code_t** d_codes, int64_t* n_words) {
    if (!c) return ABEA_ERR_ARG;
    if (!c->ran || !c->results_on_device) return fail(c, ABEA_ERR_STATE, "abea_device_codes before abea_run");
    CU(cudaSetDevice(c->device));
    const int32_t n = c->n_batch_reads;
    const int64_t words = (n > 0 && c->total_pair_cap > 0) ? code_words(c->total_pair_cap, n) : 0;
    if (words > 0 && !c->codes_current) { /* one warp per read over the lists in d_pairs; once per run */
        if (dev_reserve(c, c->d_codes, (size_t)words * sizeof(abea_code_t))) return ABEA_ERR_CUDA;
        if (dev_reserve(c, c->d_capptr, ((size_t)n + 1) * sizeof(int64_t))) return ABEA_ERR_CUDA;
        if (host_reserve(c, c->h_items, ((size_t)n + 2) * sizeof(int64_t))) return ABEA_ERR_CUDA; /* pinned scratch */
        memcpy(c->h_items.p, c->cap_ptr.data(), ((size_t)n + 1) * sizeof(int64_t));
        CU(cudaMemcpyAsync(c->d_capptr.p, c->h_items.p, ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        const int blocks = std::max(1, std::min((n + 7) / 8, c->sm_count * 8));
        ABEA_LAUNCH(abea_pairs_to_codes_kernel, blocks, 256, c->stream, (const abea_pair_t*)c->d_pairs.p,
                    (const int64_t*)c->d_capptr.p, (const int32_t*)c->d_npairs.p, n, (abea_code_t*)c->d_codes.p);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
        c->codes_current = true;
    }
    if (d_codes) *d_codes = (const abea_code_t*)c->d_codes.p;
    if (n_words) *n_words = words;
    return ABEA_OK;
}

int abea_expand_codes(abea_ctx_t* c, const abea_code_t* d_codes, const int32_t* d_n_pairs, const int64_t* cap_ptr,
                      int32_t n_reads, abea_pair_t* d_dst, int64_t dst_capacity, int64_t* d_total, int64_t* total_pairs) {
    if (!c || n_reads < 0 || (n_reads > 0 && (!d_codes || !d_n_pairs || !cap_ptr))) return fail(c, ABEA_ERR_ARG, "bad codes");
    CU(cudaSetDevice(c->device));
    if (total_pairs) *total_pairs = 0;
    if (n_reads == 0) return ABEA_OK;
    if (cap_ptr[n_reads] > dst_capacity) /* a list never has more pairs than capacity slots */
        return fail(c, ABEA_ERR_ARG, "dense buffer holds %lld pairs, up to %lld needed", (long long)dst_capacity, (long long)cap_ptr[n_reads]);
    if (!d_dst) return fail(c, ABEA_ERR_ARG, "no destination");
    if (dev_reserve(c, c->d_xcap, ((size_t)n_reads + 1) * sizeof(int64_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_xoff, ((size_t)n_reads + 1) * sizeof(int64_t))) return ABEA_ERR_CUDA;
    /* pageable source: the copy is staged by the runtime before the call returns (the caller may reuse cap_ptr) */
    CU(cudaMemcpyAsync(c->d_xcap.p, cap_ptr, ((size_t)n_reads + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    ABEA_LAUNCH(abea_pair_offsets_kernel, 1, ABEA_SCAN_THREADS, c->stream, d_n_pairs, n_reads, (int64_t*)c->d_xoff.p);
    const int blocks = std::max(1, std::min((n_reads + 7) / 8, c->sm_count * 8));
    ABEA_LAUNCH(abea_expand_codes_kernel, blocks, 256, c->stream, d_codes, (const int64_t*)c->d_xcap.p, d_n_pairs,
                (const int64_t*)c->d_xoff.p, n_reads, d_dst);
    CU(cudaGetLastError());
    if (d_total) CU(cudaMemcpyAsync(d_total, (const int64_t*)c->d_xoff.p + n_reads, sizeof(int64_t), cudaMemcpyDeviceToDevice, c->stream));
    if (total_pairs) {
        if (host_reserve(c, c->h_items, 64)) return ABEA_ERR_CUDA;
        CU(cudaMemcpyAsync(c->h_items.p, (const int64_t*)c->d_xoff.p + n_reads, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        *total_pairs = *(const int64_t*)c->h_items.p;
    }
    return ABEA_OK;
}

/* ---- event detection ----------------------------------------------------------------------------------------- */

/* raw_kind: 0 the samples are s->raw (float, host); 1 they are s->raw_i16 (int16 ADC counts, host: half the bytes over
 * PCIe, widened on the device as src/f5cio.c:461 does on the host); 2 they are already in d_raw at s->raw_ptr (put
 * there by the BLOW5 decoder) */
static int getevents_impl(abea_ctx_t* c, const abea_signals_t* s, int raw_kind, int rna, int32_t* n_events_out, abea_timing_t* timing) {
    if (!c || !s || s->n_reads < 0 || !n_events_out) return fail(c, ABEA_ERR_ARG, "bad signals");
    if (s->n_reads > 0 && (!s->raw_ptr || !s->n_samples)) return fail(c, ABEA_ERR_ARG, "bad signals");
    if (s->n_reads > 0 && ((raw_kind == 0 && !s->raw) || (raw_kind == 1 && !s->raw_i16))) return fail(c, ABEA_ERR_ARG, "bad signals");
    if (s->offset && (!s->range || !s->digitisation)) return fail(c, ABEA_ERR_ARG, "offset without range / digitisation");
    CU(cudaSetDevice(c->device));
    c->events_ready = false;
    if (c->means_from_evcap) { /* a resident batch whose means came out of the previous event tables dies with them */
        c->uploaded = false;
        c->ran = false;
        c->means_from_evcap = false;
    }
    const int32_t n = s->n_reads;
    c->sigs.assign((size_t)n, abea_sig_t());
    int64_t raw_total = 0, sum_total = 0, cap_total = 0, ts_total = 0;
    const int32_t cl = c->evt_chunk, capc = cl / 2 + 8;
    std::vector<abea_chunk_t> chunks;
    for (int32_t i = 0; i < n; i++) {
        abea_sig_t& g = c->sigs[i];
        const int32_t ns = s->n_samples[i] > 0 ? s->n_samples[i] : 0;
        g.raw_off = s->raw_ptr[i];
        g.sum_off = sum_total;
        g.ts_off = ts_total;
        g.cap_off = cap_total;
        g.chunk_off = (int64_t)chunks.size();
        if (ns >= 100)
            for (int32_t q = 0; q < (ns + cl - 1) / cl; q++) chunks.push_back(abea_chunk_t{i, q});
        g.n_samples = ns;
        g.cap = ns / 2 + 2;
        g.offset = s->offset ? s->offset[i] : 0.f;
        g.raw_unit = s->offset ? s->range[i] / s->digitisation[i] : 0.f; /* float division, src/f5c.c:693 */
        raw_total = std::max(raw_total, g.raw_off + ns);
        sum_total += (int64_t)ns + 1;
        ts_total += ((int64_t)ns + 3) & ~(int64_t)3;
        cap_total += g.cap;
    }
    /* the detector runs one thread per read: longest first, so that the 32 reads of a warp end together */
    std::vector<int32_t> order((size_t)n);
    for (int32_t i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return c->sigs[x].n_samples > c->sigs[y].n_samples; });
    if (dev_reserve(c, c->d_raw, (size_t)(raw_total + 1) * sizeof(float))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_sum, (size_t)(sum_total + 1) * sizeof(double))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_sumsq, (size_t)(sum_total + 1) * sizeof(double))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_ts1, (size_t)(ts_total + 4) * sizeof(float))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_ts2, (size_t)(ts_total + 4) * sizeof(float))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_sigorder, ((size_t)n + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    const size_t nck = chunks.size();
    if (dev_reserve(c, c->d_chunks, (nck + 1) * sizeof(abea_chunk_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_spec, (nck + 1) * (size_t)capc * sizeof(int2))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_fix, (nck + 1) * (size_t)capc * sizeof(int2))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_spec_cnt, (nck + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_fix_cnt, (nck + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_sync, (nck + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_spec_end, (nck + 1) * sizeof(evt_state_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_peaks, (size_t)(cap_total + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_evcap, (size_t)(cap_total + 1) * sizeof(abea_event_t))) return ABEA_ERR_CUDA;
    c->evcap_total = cap_total;
    if (dev_reserve(c, c->d_sigs, ((size_t)n + 1) * sizeof(abea_sig_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_nev, ((size_t)n + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    CU(cudaEventRecord(c->ev[EV_H2D0], c->stream));
    if (n > 0) {
        CU(cudaMemcpyAsync(c->d_sigs.p, c->sigs.data(), (size_t)n * sizeof(abea_sig_t), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_sigorder.p, order.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        if (nck) CU(cudaMemcpyAsync(c->d_chunks.p, chunks.data(), nck * sizeof(abea_chunk_t), cudaMemcpyHostToDevice, c->stream));
        if (raw_total > 0 && raw_kind == 0)
            CU(cudaMemcpyAsync(c->d_raw.p, s->raw, (size_t)raw_total * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        if (raw_total > 0 && raw_kind == 1) {
            if (dev_reserve(c, c->d_raw16, (size_t)(raw_total + 1) * sizeof(int16_t))) return ABEA_ERR_CUDA;
            CU(cudaMemcpyAsync(c->d_raw16.p, s->raw_i16, (size_t)raw_total * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream));
            const int blocks = (int)std::min<int64_t>((raw_total + 255) / 256, (int64_t)c->sm_count * 16);
            ABEA_LAUNCH(abea_i16_to_f32_kernel, blocks, 256, c->stream, (const int16_t*)c->d_raw16.p, (float*)c->d_raw.p, raw_total);
        }
    }
    CU(cudaEventRecord(c->ev[EV_H2D1], c->stream));
    abea_det_param_t P;
    if (rna) { P.w1 = 7; P.w2 = 14; P.thr1 = 2.5f; P.thr2 = 9.0f; P.peak_height = 1.0f; }  /* src/events.c:59-63 */
    else     { P.w1 = 3; P.w2 = 6;  P.thr1 = 1.4f; P.thr2 = 9.0f; P.peak_height = 0.2f; }  /* src/events.c:52-56 */
    CU(cudaEventRecord(c->ev[EV_S0], c->stream));
    if (n > 0) {
        ABEA_LAUNCH(abea_events_sums_kernel, (n + EVT_WARPS - 1) / EVT_WARPS, 32 * EVT_WARPS, c->stream,
                    (const abea_sig_t*)c->d_sigs.p, n, (const float*)c->d_raw.p, (double*)c->d_sum.p, (double*)c->d_sumsq.p);
        if (nck) {
            ABEA_LAUNCH(abea_events_tstat_kernel, (int)nck, 128, c->stream, (const abea_sig_t*)c->d_sigs.p,
                        (const abea_chunk_t*)c->d_chunks.p, cl, (const double*)c->d_sum.p, (const double*)c->d_sumsq.p,
                        (float*)c->d_ts1.p, (float*)c->d_ts2.p, P);
            ABEA_LAUNCH(abea_events_spec_kernel, (int)((nck + 63) / 64), 64, c->stream, (const abea_sig_t*)c->d_sigs.p,
                        (const abea_chunk_t*)c->d_chunks.p, (int32_t)nck, cl, capc, (const float*)c->d_ts1.p,
                        (const float*)c->d_ts2.p, (int2*)c->d_spec.p, (int32_t*)c->d_spec_cnt.p,
                        (evt_state_t*)c->d_spec_end.p, P);
        }
        ABEA_LAUNCH(abea_events_stitch_kernel, (n + 31) / 32, 32, c->stream, (const abea_sig_t*)c->d_sigs.p,
                    (const int32_t*)c->d_sigorder.p, n, cl, capc, (const float*)c->d_ts1.p, (const float*)c->d_ts2.p,
                    (const int2*)c->d_spec.p, (const int32_t*)c->d_spec_cnt.p, (const evt_state_t*)c->d_spec_end.p,
                    (int2*)c->d_fix.p, (int32_t*)c->d_fix_cnt.p, (int32_t*)c->d_sync.p, (int32_t*)c->d_nev.p, P);
        ABEA_LAUNCH(abea_events_create_kernel, (n + EVT_WARPS - 1) / EVT_WARPS, 32 * EVT_WARPS, c->stream,
                    (const abea_sig_t*)c->d_sigs.p, n, cl, capc, (const double*)c->d_sum.p, (const double*)c->d_sumsq.p,
                    (const int2*)c->d_spec.p, (const int32_t*)c->d_spec_cnt.p, (const int2*)c->d_fix.p,
                    (const int32_t*)c->d_fix_cnt.p, (const int32_t*)c->d_sync.p, (int32_t*)c->d_peaks.p,
                    (abea_event_t*)c->d_evcap.p, (const int32_t*)c->d_nev.p);
    }
    CU(cudaEventRecord(c->ev[EV_S1], c->stream));
    c->nev.assign((size_t)n, 0);
    if (n > 0) CU(cudaMemcpyAsync(c->nev.data(), c->d_nev.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    for (int32_t i = 0; i < n; i++) n_events_out[i] = c->nev[i];
    c->events_ready = true;
    c->last.events_ms = ev_ms(c, EV_S0, EV_S1);
    c->last.h2d_ms = ev_ms(c, EV_H2D0, EV_H2D1);
    c->last.n_samples = raw_total;
    if (timing) *timing = c->last;
    return ABEA_OK;
}

int abea_getevents(abea_ctx_t* c, const abea_signals_t* s, int rna, int32_t* n_events_out, abea_timing_t* timing) {
    if (!s) return fail(c, ABEA_ERR_ARG, "bad signals");
    return getevents_impl(c, s, (s->raw == nullptr && s->raw_i16 != nullptr) ? 1 : 0, rna, n_events_out, timing);
}

/* BLOW5 records in, event tables out: what read_slow5_single + event_single do per read on the host —
 * slow5lib's record decompression, record parsing and signal decompression (slow5_press.c:921-1010, 1118-1170;
 * slow5.c:2840-2930), the widening to float (src/f5cio.c:461), the conversion to pA and getevents (src/f5c.c:684-703)
 * — for a whole batch on the device. The file's own bytes cross PCIe; the raw signal never exists on the host. */
int abea_getevents_blow5(abea_ctx_t* c, const abea_blow5_t* f, int rna, int32_t* n_events_out, int32_t* n_samples_out,
                         abea_timing_t* timing) {
    if (!c || !f || f->n_reads < 0 || !n_events_out) return fail(c, ABEA_ERR_ARG, "bad records");
    const int32_t n = f->n_reads;
    if (n > 0 && (!f->bytes || !f->rec_ptr || !f->rec_len)) return fail(c, ABEA_ERR_ARG, "bad records");
    if (f->record_method != B5_REC_NONE && f->record_method != B5_REC_ZLIB)
        return fail(c, ABEA_ERR_ARG, "record compression %d is not supported (none and zlib are)", f->record_method);
    if (f->signal_method != B5_SIG_NONE && f->signal_method != B5_SIG_SVB_ZD && f->signal_method != B5_SIG_EX_ZD)
        return fail(c, ABEA_ERR_ARG, "signal compression %d is not supported (none, svb-zd and ex-zd are)", f->signal_method);
    CU(cudaSetDevice(c->device));
    const double t0 = now_ms();
    int64_t total_in = 0;
    for (int32_t i = 0; i < n; i++) {
        if (f->rec_len[i] < 0 || f->rec_ptr[i] < 0) return fail(c, ABEA_ERR_ARG, "bad record %d", i);
        total_in = std::max(total_in, f->rec_ptr[i] + (int64_t)f->rec_len[i]);
    }
    if (dev_reserve(c, c->d_b5in, (size_t)total_in + 16)) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_b5recs, ((size_t)n + 1) * sizeof(abea_b5rec_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_b5len, ((size_t)n + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_b5status, ((size_t)n + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_b5hdr, ((size_t)n + 1) * sizeof(abea_b5hdr_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_b5rawoff, ((size_t)n + 1) * sizeof(int64_t))) return ABEA_ERR_CUDA;
    const size_t hb = ((size_t)n + 1) * (sizeof(abea_b5rec_t) + sizeof(abea_b5hdr_t) + 2 * sizeof(int32_t) + sizeof(int64_t));
    if (host_reserve(c, c->h_b5, hb)) return ABEA_ERR_CUDA;
    abea_b5hdr_t* h_hdr = (abea_b5hdr_t*)c->h_b5.p;
    abea_b5rec_t* h_recs = (abea_b5rec_t*)(h_hdr + (n + 1));
    int64_t* h_rawoff = (int64_t*)(h_recs + (n + 1));
    int32_t* h_len = (int32_t*)(h_rawoff + (n + 1));
    int32_t* h_status = h_len + (n + 1);
    CU(cudaEventRecord(c->ev[EV_H2D0], c->stream));
    if (total_in > 0) CU(cudaMemcpyAsync(c->d_b5in.p, f->bytes, (size_t)total_in, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev[EV_H2D1], c->stream));
    CU(cudaEventRecord(c->ev[EV_D2H0], c->stream)); /* reused as "decode starts" */
    const uint8_t* d_data = (const uint8_t*)c->d_b5in.p;
    if (f->record_method == B5_REC_ZLIB && n > 0) {
        /* a record's inflated size is not stored: guess generously (a raw signal deflates to ~60 %), let the kernel
         * report an overflow, and retry with more — up to DEFLATE's own ceiling of 1032 : 1 */
        const size_t smem = (size_t)B5_INFLATE_WARPS * sizeof(b5_tables_t);
        CU(cudaFuncSetAttribute((const void*)abea_inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int64_t mult = 4;; mult *= 4) {
            int64_t off = 0;
            for (int32_t i = 0; i < n; i++) {
                h_recs[i].in_off = f->rec_ptr[i];
                h_recs[i].in_len = f->rec_len[i];
                const int64_t cap = std::min<int64_t>(((int64_t)f->rec_len[i] * std::min<int64_t>(mult, 1032) + 4096 + 15) & ~(int64_t)15, 0x7ffffff0);
                h_recs[i].out_off = off;
                h_recs[i].out_cap = (int32_t)cap;
                off += cap;
            }
            if (dev_reserve(c, c->d_b5out, (size_t)off + 16)) return ABEA_ERR_CUDA;
            CU(cudaMemcpyAsync(c->d_b5recs.p, h_recs, (size_t)n * sizeof(abea_b5rec_t), cudaMemcpyHostToDevice, c->stream));
            ABEA_LAUNCH_SMEM(abea_inflate_kernel, (n + B5_INFLATE_WARPS - 1) / B5_INFLATE_WARPS, 32 * B5_INFLATE_WARPS, smem, c->stream,
                             (const abea_b5rec_t*)c->d_b5recs.p, n, (const uint8_t*)c->d_b5in.p, (uint8_t*)c->d_b5out.p,
                             (int32_t*)c->d_b5len.p, (int32_t*)c->d_b5status.p);
            CU(cudaMemcpyAsync(h_status, c->d_b5status.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(c->stream));
            bool overflow = false;
            for (int32_t i = 0; i < n; i++) {
                if (h_status[i] == B5_ERR_DATA) return fail(c, ABEA_ERR_ARG, "record %d: malformed zlib stream", i);
                overflow |= (h_status[i] == B5_ERR_OVERFLOW);
            }
            if (!overflow) break;
            if (mult >= 1032) return fail(c, ABEA_ERR_ARG, "a record inflates beyond DEFLATE's maximum ratio");
        }
        d_data = (const uint8_t*)c->d_b5out.p;
    } else if (n > 0) {
        for (int32_t i = 0; i < n; i++) {
            h_recs[i].in_off = f->rec_ptr[i];
            h_recs[i].in_len = f->rec_len[i];
            h_recs[i].out_off = f->rec_ptr[i];
            h_recs[i].out_cap = f->rec_len[i];
            h_len[i] = f->rec_len[i];
        }
        CU(cudaMemcpyAsync(c->d_b5recs.p, h_recs, (size_t)n * sizeof(abea_b5rec_t), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_b5len.p, h_len, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    }
    std::vector<int32_t> ns((size_t)n, 0);
    std::vector<int64_t> raw_ptr((size_t)n, 0);
    std::vector<float> cal_off((size_t)n, 0.f), cal_range((size_t)n, 1.f), cal_dig((size_t)n, 1.f);
    int64_t raw_total = 0;
    if (n > 0) {
        ABEA_LAUNCH(abea_blow5_parse_kernel, (n + 127) / 128, 128, c->stream, (const abea_b5rec_t*)c->d_b5recs.p, n, d_data,
                    (const int32_t*)c->d_b5len.p, (const int32_t*)nullptr, f->signal_method, (abea_b5hdr_t*)c->d_b5hdr.p);
        CU(cudaMemcpyAsync(h_hdr, c->d_b5hdr.p, (size_t)n * sizeof(abea_b5hdr_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
        for (int32_t i = 0; i < n; i++) {
            if (h_hdr[i].status != B5_OK) return fail(c, ABEA_ERR_ARG, "record %d: malformed BLOW5 record", i);
            ns[i] = h_hdr[i].n_samples;
            raw_ptr[i] = raw_total;
            h_rawoff[i] = raw_total;
            raw_total += ns[i];
            cal_dig[i] = (float)h_hdr[i].digitisation; /* the narrowing of read_slow5_single, src/f5cio.c:455-458 */
            cal_off[i] = (float)h_hdr[i].offset;
            cal_range[i] = (float)h_hdr[i].range;
        }
        if (dev_reserve(c, c->d_raw, (size_t)(raw_total + 1) * sizeof(float))) return ABEA_ERR_CUDA;
        CU(cudaMemcpyAsync(c->d_b5rawoff.p, h_rawoff, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        uint32_t* ex_scratch = nullptr; /* ex-zd: positions and values of a record's exceptions, at most one per sample */
        if (f->signal_method == B5_SIG_EX_ZD) {
            if (dev_reserve(c, c->d_b5ex, (size_t)(2 * raw_total + 2) * sizeof(uint32_t))) return ABEA_ERR_CUDA;
            ex_scratch = (uint32_t*)c->d_b5ex.p;
        }
        ABEA_LAUNCH(abea_blow5_signal_kernel, (n + B5_SIG_WARPS - 1) / B5_SIG_WARPS, 32 * B5_SIG_WARPS, c->stream,
                    (const abea_b5rec_t*)c->d_b5recs.p, n, d_data, (const abea_b5hdr_t*)c->d_b5hdr.p,
                    (const int64_t*)c->d_b5rawoff.p, f->signal_method, (float*)c->d_raw.p, (int32_t*)c->d_b5status.p, ex_scratch);
        CU(cudaEventRecord(c->ev[EV_D2H1], c->stream));
        CU(cudaMemcpyAsync(h_status, c->d_b5status.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
        for (int32_t i = 0; i < n; i++)
            if (h_status[i] != B5_OK) return fail(c, ABEA_ERR_ARG, "record %d: malformed compressed signal", i);
    } else {
        CU(cudaEventRecord(c->ev[EV_D2H1], c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    const float decode_ms = ev_ms(c, EV_D2H0, EV_D2H1);
    const float in_ms = ev_ms(c, EV_H2D0, EV_H2D1);
    if (n_samples_out)
        for (int32_t i = 0; i < n; i++) n_samples_out[i] = ns[i];
    abea_signals_t s;
    memset(&s, 0, sizeof(s));
    s.n_reads = n;
    s.raw_ptr = raw_ptr.data();
    s.n_samples = ns.data();
    s.offset = cal_off.data();
    s.range = cal_range.data();
    s.digitisation = cal_dig.data();
    const int rc = getevents_impl(c, &s, 2, rna, n_events_out, nullptr);
    if (rc) return rc;
    c->last.blow5_ms = decode_ms;
    c->last.h2d_ms = in_ms;
    c->last.h2d_bytes = total_in;
    c->last.pack_ms = now_ms() - t0;
    if (timing) *timing = c->last;
    return ABEA_OK;
}

/* The float samples the last abea_getevents / abea_getevents_blow5 worked on (ADC counts as widened from int16, or
 * whatever the caller handed over), read i at raw[raw_ptr[i] ..]: lets a test compare the device-side decoders. */
int abea_raw_download(abea_ctx_t* c, float* raw, const int64_t* raw_ptr) {
    if (!c || !raw_ptr) return fail(c, ABEA_ERR_ARG, "bad output");
    if (!c->events_ready) return fail(c, ABEA_ERR_STATE, "abea_raw_download before abea_getevents");
    CU(cudaSetDevice(c->device));
    for (size_t i = 0; i < c->sigs.size(); i++)
        if (c->sigs[i].n_samples > 0) {
            if (!raw) return fail(c, ABEA_ERR_ARG, "bad output");
            CU(cudaMemcpy(raw + raw_ptr[i], (const float*)c->d_raw.p + c->sigs[i].raw_off, (size_t)c->sigs[i].n_samples * sizeof(float),
                          cudaMemcpyDeviceToHost));
        }
    return ABEA_OK;
}

int abea_getevents_download(abea_ctx_t* c, abea_event_t* events, const int64_t* event_ptr) {
    if (!c || !event_ptr) return fail(c, ABEA_ERR_ARG, "bad output");
    if (!c->events_ready) return fail(c, ABEA_ERR_STATE, "abea_getevents_download before abea_getevents");
    CU(cudaSetDevice(c->device));
    const int32_t n = (int32_t)c->sigs.size();
    int64_t total = 0;
    for (int32_t i = 0; i < n; i++)
        if (c->nev[i] > 0) total = std::max(total, event_ptr[i] + (int64_t)c->nev[i]);
    if (total == 0) return ABEA_OK;
    if (!events) return fail(c, ABEA_ERR_ARG, "bad output");
    if (dev_reserve(c, c->d_evptr, ((size_t)n + 1) * sizeof(int64_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_evout, (size_t)(total + 1) * sizeof(abea_event_t))) return ABEA_ERR_CUDA;
    CU(cudaMemcpyAsync(c->d_evptr.p, event_ptr, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_evout.p, 0, (size_t)total * sizeof(abea_event_t), c->stream));
    ABEA_LAUNCH(abea_events_compact_kernel, (n + 3) / 4, 128, c->stream, (const abea_sig_t*)c->d_sigs.p, n,
                (const abea_event_t*)c->d_evcap.p, (const int32_t*)c->d_nev.p, (const int64_t*)c->d_evptr.p,
                (abea_event_t*)c->d_evout.p);
    CU(cudaMemcpyAsync(events, c->d_evout.p, (size_t)total * sizeof(abea_event_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return ABEA_OK;
}

/* ---- the stages either side of ABEA ------------------------------------------------------------------------- */

static int scaling_descriptors(abea_ctx_t* c) {
    const size_t nb = (size_t)c->n_batch_reads;
    if (!c->sreads_on_device) {
        if (dev_reserve(c, c->d_sreads, (nb + 1) * sizeof(abea_sread_t))) return ABEA_ERR_CUDA;
        if (dev_reserve(c, c->d_scalings, (nb + 1) * sizeof(abea_scalings_t))) return ABEA_ERR_CUDA;
        if (nb) CU(cudaMemcpyAsync(c->d_sreads.p, c->sreads.data(), nb * sizeof(abea_sread_t), cudaMemcpyHostToDevice, c->stream));
        c->sreads_on_device = true;
    }
    if (!c->scalings_on_device && !c->in_scalings.empty()) {
        CU(cudaMemcpyAsync(c->d_scalings.p, c->in_scalings.data(), nb * sizeof(abea_scalings_t), cudaMemcpyHostToDevice, c->stream));
        c->scalings_on_device = true;
    }
    return ABEA_OK;
}

int abea_estimate_scalings(abea_ctx_t* c, int reverse_events, abea_scalings_t* scalings_out, abea_timing_t* timing) {
    if (!c) return ABEA_ERR_ARG;
    if (!c->uploaded) return fail(c, ABEA_ERR_STATE, "abea_estimate_scalings before abea_upload_batch");
    if (c->streaming) return fail(c, ABEA_ERR_STATE, "abea_estimate_scalings needs a resident batch (abea_upload_batch)");
    CU(cudaSetDevice(c->device));
    const int32_t nb = c->n_batch_reads;
    int rc = scaling_descriptors(c);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[EV_S0], c->stream));
    if (nb > 0) {
        if (!c->scalings_on_device) CU(cudaMemsetAsync(c->d_scalings.p, 0, (size_t)nb * sizeof(abea_scalings_t), c->stream));
        ABEA_LAUNCH(abea_mom_kernel, (nb + SCL_WARPS - 1) / SCL_WARPS, 32 * SCL_WARPS, c->stream,
                    (const abea_sread_t*)c->d_sreads.p, nb, (const uint8_t*)c->d_seq.p, (float*)c->d_means.p,
                    (const abea_model_t*)c->d_model.p, c->kmer_size, (abea_scalings_t*)c->d_scalings.p,
                    (abea_read_t*)c->d_reads.p, (int32_t)((reverse_events && !c->means_reversed) ? 1 : 0));
    }
    CU(cudaEventRecord(c->ev[EV_S1], c->stream));
    if (scalings_out && nb > 0)
        CU(cudaMemcpyAsync(scalings_out, c->d_scalings.p, (size_t)nb * sizeof(abea_scalings_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    c->scalings_on_device = true;
    c->need_scalings = false;
    if (reverse_events) c->means_reversed = true; /* a second request leaves the resident means 3'->5' as they are */
    c->prepared = false; /* the k-mer parameter cache depends on scale / shift */
    c->last.mom_ms = ev_ms(c, EV_S0, EV_S1);
    if (timing) *timing = c->last;
    return ABEA_OK;
}

int abea_scaling_stage(abea_ctx_t* c, int32_t min_num_events_to_rescale, abea_timing_t* timing) {
    if (!c) return ABEA_ERR_ARG;
    if (!c->ran || !c->results_on_device) return fail(c, ABEA_ERR_STATE, "abea_scaling_stage before abea_run");
    CU(cudaSetDevice(c->device));
    const int32_t nb = c->n_batch_reads;
    int rc = scaling_descriptors(c);
    if (rc) return rc;
    if (!c->scalings_on_device) return fail(c, ABEA_ERR_STATE, "no scalings on the device");
    if (dev_reserve(c, c->d_maps, (size_t)(c->total_map + 1) * sizeof(abea_index_pair_t))) return ABEA_ERR_CUDA;
    if (dev_reserve(c, c->d_sres, ((size_t)nb + 1) * sizeof(abea_scaling_result_t))) return ABEA_ERR_CUDA;
    CU(cudaEventRecord(c->ev[EV_S0], c->stream));
    if (nb > 0)
        ABEA_LAUNCH(abea_scaling_kernel, (nb + SCL_WARPS - 1) / SCL_WARPS, 32 * SCL_WARPS, c->stream,
                    (const abea_sread_t*)c->d_sreads.p, nb, (const uint8_t*)c->d_seq.p,
                    (const float*)c->d_means.p, (const abea_model_t*)c->d_model.p, c->kmer_size,
                    (const abea_pair_t*)c->d_pairs.p, (const int32_t*)c->d_npairs.p,
                    (const abea_scalings_t*)c->d_scalings.p, (abea_index_pair_t*)c->d_maps.p,
                    (abea_scaling_result_t*)c->d_sres.p, min_num_events_to_rescale);
    CU(cudaEventRecord(c->ev[EV_S1], c->stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    c->scaled = true;
    c->last.scaling_ms = ev_ms(c, EV_S0, EV_S1);
    if (timing) *timing = c->last;
    return ABEA_OK;
}

int abea_scaling_download(abea_ctx_t* c, abea_scaling_result_t* results, abea_index_pair_t* maps, const int64_t* map_ptr) {
    if (!c || !results) return fail(c, ABEA_ERR_ARG, "bad output");
    if (!c->scaled) return fail(c, ABEA_ERR_STATE, "abea_scaling_download before abea_scaling_stage");
    if (maps && !map_ptr) return fail(c, ABEA_ERR_ARG, "maps without map_ptr");
    CU(cudaSetDevice(c->device));
    const int32_t nb = c->n_batch_reads;
    if (nb > 0) CU(cudaMemcpy(results, c->d_sres.p, (size_t)nb * sizeof(abea_scaling_result_t), cudaMemcpyDeviceToHost));
    for (int32_t i = 0; i < nb; i++) /* CACHED_LOG: scallings->log_var = log(var), glibc double log (src/align.c:757) */
        if (results[i].calibrated) results[i].scalings.log_var = (float)log(results[i].var_d);
    if (maps && c->total_map > 0) {
        bool canonical = true;
        for (int32_t i = 0; i < nb && canonical; i++) canonical = (map_ptr[i] == c->sreads[i].map_off);
        if (canonical) {
            CU(cudaMemcpy(maps, c->d_maps.p, (size_t)c->total_map * sizeof(abea_index_pair_t), cudaMemcpyDeviceToHost));
        } else {
            for (int32_t i = 0; i < nb; i++) {
                const int64_t K = (int64_t)c->sreads[i].read_len - (int64_t)c->kmer_size + 1;
                if (K > 0 && results[i].n_event_alignment > 0)
                    CU(cudaMemcpy(maps + map_ptr[i], (const abea_index_pair_t*)c->d_maps.p + c->sreads[i].map_off,
                                  (size_t)K * sizeof(abea_index_pair_t), cudaMemcpyDeviceToHost));
            }
        }
    }
    return ABEA_OK;
}

int abea_scaling_device_results(abea_ctx_t* c, const abea_scaling_result_t** d_results, const abea_index_pair_t** d_maps,
                                int64_t* total_map_entries) {
    if (!c) return ABEA_ERR_ARG;
    if (!c->scaled) return fail(c, ABEA_ERR_STATE, "abea_scaling_device_results before abea_scaling_stage");
    if (d_results) *d_results = (const abea_scaling_result_t*)c->d_sres.p;
    if (d_maps) *d_maps = (const abea_index_pair_t*)c->d_maps.p;
    if (total_map_entries) *total_map_entries = c->total_map;
    return ABEA_OK;
}

int abea_align_batch(abea_ctx_t* c, const abea_batch_t* batch, abea_pair_t* pairs, const int64_t* pair_ptr,
                     int32_t* n_pairs, abea_timing_t* timing) {
    if (!c || !batch) return fail(c, ABEA_ERR_ARG, "bad batch");
    if (!n_pairs || !pair_ptr) return fail(c, ABEA_ERR_ARG, "bad output");
    /* Pinned (mapped) caller buffers are streamed: events in over PCIe while the fill runs, pair lists out as each
     * read finishes. Anything else is staged through the copy engine (abea_upload_batch / abea_download). */
    const void* ev_alias = nullptr;
    const void* ev_host = batch->event_means ? (const void*)batch->event_means : (const void*)batch->events;
    if ((c->stream_mode & 1) && batch->n_reads > 0 && ev_host && ((uintptr_t)ev_host & 15) == 0) ev_alias = mapped_alias(ev_host);
    int rc = upload_impl(c, batch, ev_alias, nullptr);
    if (rc) return rc;
    abea_pair_t* fin_pairs = nullptr;
    int32_t* fin_np = nullptr;
    abea_code_t* fin_codes = nullptr;
    const int32_t n = batch->n_reads;
    /* Pair lists out. (a) As path codes: the traceback writes 8 bytes per 32 pairs into the context's pinned staging and
     * publishes each read's count behind them; host threads expand the lists into the caller's buffer — pinned or not,
     * any pair_ptr layout — while the kernels are still running. (b) Whole, into the caller's mapped buffer. (c) Through
     * the copy engine after the kernels (abea_download). */
    if ((c->stream_mode & 4) && c->host_threads > 0 && n > 0 && c->total_pair_cap > 0 && pairs) {
        if (host_reserve(c, c->h_codes, (size_t)code_words(c->total_pair_cap, n) * sizeof(abea_code_t))) return ABEA_ERR_CUDA;
        if (host_reserve(c, c->h_fnp, ((size_t)n + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
        fin_codes = (abea_code_t*)mapped_alias(c->h_codes.p);
        fin_np = (int32_t*)mapped_alias(c->h_fnp.p);
        if (!fin_codes || !fin_np) fin_codes = nullptr, fin_np = nullptr;
    }
    if (!fin_codes && (c->stream_mode & 2) && n > 0 && c->total_pair_cap > 0) {
        bool canonical = true;
        for (int32_t i = 0; i < n && canonical; i++) canonical = (pair_ptr[i] == c->cap_ptr[i]);
        if (canonical) {
            fin_pairs = (abea_pair_t*)mapped_alias(pairs);
            fin_np = (int32_t*)mapped_alias(n_pairs);
            if (!fin_pairs || !fin_np) fin_pairs = nullptr, fin_np = nullptr;
        }
    }
    if (fin_np) memset(n_pairs, 0, (size_t)n * sizeof(int32_t)); /* reads that are not scheduled */
    if (c->need_scalings) { /* batch->scalings == NULL: method-of-moments estimate on the device first */
        rc = abea_estimate_scalings(c, 0, nullptr, nullptr);
        if (rc) return rc;
    }
    std::atomic<int> abort_flag(0);
    std::atomic<int64_t> code_bytes(0);
    if (fin_codes) {
        volatile int32_t* h_np = (volatile int32_t*)c->h_fnp.p;
        for (int32_t i = 0; i < n; i++) h_np[i] = -1; /* "not done": the traceback stores the count when the codes are out */
        const abea_code_t* h_codes = (const abea_code_t*)c->h_codes.p;
        const int T = c->host_threads;
        /* scheduled read j belongs to thread j % T (the schedule is longest-first, so the shares are even); a thread
         * keeps sweeping its reads for counts that have appeared — no prediction of the finishing order is needed */
        c->pool.kick(T, [&, h_np, h_codes, T](int tid) {
            std::vector<int32_t> mine;
            for (int32_t j = tid; j < (int32_t)c->reads.size(); j += T) mine.push_back(c->reads[(size_t)j].orig_index);
            int64_t bytes = 0;
            while (!mine.empty()) {
                bool any = false;
                for (size_t q = 0; q < mine.size();) {
                    const int32_t i = mine[q];
                    const int32_t np = h_np[i];
                    if (np < 0) { q++; continue; }
                    std::atomic_thread_fence(std::memory_order_acquire);
                    n_pairs[i] = np;
                    if (np > 0) decode_codes(h_codes + abea_code_offset(c->cap_ptr[(size_t)i], i), np, pairs + pair_ptr[i]);
                    bytes += code_bytes_used(np);
                    mine[q] = mine.back();
                    mine.pop_back();
                    any = true;
                }
                if (!any) {
                    if (abort_flag.load()) return;
                    sched_yield();
                }
            }
            code_bytes.fetch_add(bytes);
        });
    }
    const bool was_streamed = c->streaming;
    rc = run_impl(c, fin_pairs, fin_np, nullptr, fin_codes);
    if (fin_codes) {
        if (rc != ABEA_OK) abort_flag.store(1);
        c->pool.wait();
    }
    if (rc == ABEA_ERR_CUDA && was_streamed && cudaGetLastError() == cudaSuccess) {
        /* the stream from the caller's pinned buffer stalled (host contention, a tool slowing the loader down): the
         * batch is intact in the caller's buffers, so run it once more through the copy engine before giving up */
        rc = upload_impl(c, batch, nullptr, nullptr);
        if (rc == ABEA_OK && c->need_scalings) rc = abea_estimate_scalings(c, 0, nullptr, nullptr);
        if (rc == ABEA_OK) rc = run_impl(c, nullptr, nullptr, nullptr);
        fin_pairs = nullptr;
        fin_np = nullptr;
        fin_codes = nullptr;
    }
    if (rc) return rc;
    if (!fin_np) {
        rc = abea_download(c, pairs, pair_ptr, n_pairs, nullptr);
        if (rc) return rc;
    } else {
        int64_t np = 0;
        for (int32_t i = 0; i < n; i++) np += n_pairs[i];
        c->last.d2h_ms = 0.f;
        c->last.unpack_ms = 0.0;
        c->last.d2h_bytes = (fin_codes ? code_bytes.load() : np * (int64_t)sizeof(abea_pair_t)) + (int64_t)n * (int64_t)sizeof(int32_t);
    }
    if (timing) *timing = c->last;
    return ABEA_OK;
}

/* The reference's --print-banded-aln dump (src/f5c.c:989-1006), for diffing this path against an f5c run: per read
 * that did not fail the alignment,  ">name\tN_ALGN_PAIR:n\t{ref_pos,read_pos}\n"  then "{k,e}\t" per pair and "\n". */
int abea_write_pairs(const char* path, int append, int32_t n_reads, const char* const* names, const int32_t* n_pairs,
                     const abea_pair_t* pairs, const int64_t* pair_ptr, const uint32_t* read_stat_flag) {
    if (!path || n_reads < 0 || (n_reads > 0 && (!names || !n_pairs || !pair_ptr))) return ABEA_ERR_ARG;
    FILE* fp = strcmp(path, "-") == 0 ? stdout : fopen(path, append ? "a" : "w");
    if (!fp) return ABEA_ERR_ARG;
    for (int32_t i = 0; i < n_reads; i++) {
        if (read_stat_flag && (read_stat_flag[i] & ABEA_FAILED_ALIGNMENT)) continue;
        fprintf(fp, ">%s\tN_ALGN_PAIR:%d\t{ref_pos,read_pos}\n", names[i], (int)n_pairs[i]);
        const abea_pair_t* p = pairs + pair_ptr[i];
        for (int32_t j = 0; j < n_pairs[i]; j++) fprintf(fp, "{%d,%d}\t", p[j].ref_pos, p[j].read_pos);
        fprintf(fp, "\n");
    }
    if (fp != stdout) fclose(fp);
    else fflush(fp);
    return ABEA_OK;
}

/* f5c resquiggle's output_db_rsq (reference src/resquiggle.c:322-447), TSV (fmt 0) or PAF (fmt 1), from the event tables
 * and the k-mer -> event-range maps of the scaling stage. */
int abea_write_resquiggle(const char* path, int append, int fmt, int header, int rna, uint32_t kmer_size, int32_t n_reads,
                          const char* const* names, const int32_t* read_len, const int64_t* n_samples,
                          const abea_event_t* events, const int64_t* event_ptr, const abea_scaling_result_t* results,
                          const abea_index_pair_t* maps, const int64_t* map_ptr) {
    if (!path || n_reads < 0 || (fmt != 0 && fmt != 1)) return ABEA_ERR_ARG;
    if (n_reads > 0 && (!names || !read_len || !n_samples || !event_ptr || !results || !maps || !map_ptr)) return ABEA_ERR_ARG;
    FILE* fp = strcmp(path, "-") == 0 ? stdout : fopen(path, append ? "a" : "w");
    if (!fp) return ABEA_ERR_ARG;
    if (header && fmt == 0) fprintf(fp, "read_id\tkmer_idx\tstart_raw_idx\tend_raw_idx\n"); /* src/resquiggle.c:724-727 */
    int rc = ABEA_OK;
    std::vector<abea_index_pair_t> map;
    std::string ss;
    char num[32];
    for (int32_t i = 0; i < n_reads && rc == ABEA_OK; i++) {
        if (results[i].flags) continue; /* failed calibration / alignment / QC: counted, not printed */
        const int32_t n_kmers = read_len[i] - (int32_t)kmer_size + 1;
        if (n_kmers <= 0 || !events) { rc = ABEA_ERR_ARG; break; }
        const abea_event_t* ev = events + event_ptr[i];
        map.assign(maps + map_ptr[i], maps + map_ptr[i] + n_kmers);
        if (rna) { /* :345-357 */
            std::reverse(map.begin(), map.end());
            for (abea_index_pair_t& m : map) std::swap(m.start, m.stop);
        }
        int64_t sig_start = -1, sig_start2 = -1, sig_end = -1, sig_end2 = -1, read_start = -1, read_end = -1;
        int64_t ci = 0, mi = 0, d = 0, count_samples = 0;
        bool first = true;
        int matches = 0;
        ss.clear();
        for (int32_t j = 0; j < n_kmers; j++) {
            const int32_t se = map[(size_t)j].start, ee = map[(size_t)j].stop;
            if (se == -1) { /* deletion from the read */
                sig_start = sig_end = -1;
                if (!first) d++;
            } else {
                sig_start = (int64_t)ev[se].start; /* inclusive */
                if (first) {
                    sig_start2 = sig_start;
                    read_start = j;
                    ci = sig_start;
                    first = false;
                }
                sig_end2 = sig_end = (int64_t)ev[ee].start + (int)ev[ee].length; /* non-inclusive */
                read_end = j;
                if (fmt) {
                    if (d > 0) {
                        snprintf(num, sizeof num, "%dD", (int)d);
                        ss += num;
                        d = 0;
                    }
                    if (j == 0) ci = sig_start;
                    ci += (mi = sig_start - ci);
                    if (mi) {
                        snprintf(num, sizeof num, "%dI", (int)mi);
                        ss += num;
                        count_samples += mi;
                    }
                    ci += (mi = sig_end - sig_start);
                    if (mi) {
                        matches++;
                        snprintf(num, sizeof num, "%d,", (int)mi);
                        ss += num;
                        count_samples += mi;
                    }
                }
            }
            if (fmt == 0) {
                fprintf(fp, "%s\t%d\t", names[i], rna ? n_kmers - j - 1 : j);
                if (sig_start < 0) fprintf(fp, ".\t");
                else fprintf(fp, "%ld\t", (long)sig_start);
                if (sig_end < 0) fprintf(fp, ".");
                else fprintf(fp, "%ld", (long)sig_end);
                fprintf(fp, "\n");
                if (sig_start >= 0 && sig_end >= 0 && sig_end <= sig_start) rc = ABEA_ERR_ARG; /* the reference exits here */
            }
        }
        if (fmt == 1 && rc == ABEA_OK) {
            if (sig_start2 == -1 || sig_end2 == -1 || count_samples != sig_end2 - sig_start2) { rc = ABEA_ERR_ARG; break; } /* its asserts */
            fprintf(fp, "%s\t%ld\t%ld\t%ld\t+\t", names[i], (long)n_samples[i], (long)sig_start2, (long)sig_end2);
            fprintf(fp, "%s\t%d\t%ld\t%ld\t", names[i], n_kmers, (long)(rna ? n_kmers - read_start : read_start),
                    (long)(rna ? n_kmers - 1 - read_end : read_end + 1));
            fprintf(fp, "%d\t%d\t%d\t", matches, n_kmers, 255);
            fprintf(fp, "sc:f:%f\t", results[0].scalings.scale); /* db->scalings->scale: the batch's first read, :442 */
            fprintf(fp, "sh:f:%f\t", results[0].scalings.shift);
            fprintf(fp, "ss:Z:%s\n", ss.c_str());
        }
    }
    if (fp != stdout) fclose(fp);
    else fflush(fp);
    return rc;
}

/* ---- the ragged front door ------------------------------------------------------------------------------------ */

int abea_scheduler_model(abea_ctx_t* c, double* cycles4) {
    if (!c || !cycles4) return ABEA_ERR_ARG;
    cycles4[0] = c->cyc_wide;
    cycles4[1] = c->cyc_narrow;
    cycles4[2] = c->cyc_long;
    cycles4[3] = c->cyc_trace;
    return ABEA_OK;
}

int abea_align_ragged(abea_ctx_t* c, const abea_ragged_t* r, int threads, abea_timing_t* timing) {
    if (!c || !r || r->n_reads < 0) return fail(c, ABEA_ERR_ARG, "bad batch");
    if (!c->have_model) return fail(c, ABEA_ERR_NOMODEL, "abea_set_model has not been called");
    const int32_t n = r->n_reads;
    if (n > 0 && (!r->seq || !r->read_len || !r->events || !r->n_events || !r->scalings || !r->pairs || !r->n_pairs))
        return fail(c, ABEA_ERR_ARG, "bad batch");
    CU(cudaSetDevice(c->device));
    const double t0 = now_ms();
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;

    /* flat layout of the staging, caller's order */
    if (host_reserve(c, c->h_rmeta, ((size_t)n + 1) * (3 * sizeof(int64_t) + 2 * sizeof(int32_t)))) return ABEA_ERR_CUDA;
    int64_t* seq_ptr = (int64_t*)c->h_rmeta.p;
    int64_t* event_ptr = seq_ptr + (n + 1);
    int64_t* pair_ptr = event_ptr + (n + 1);
    int32_t* n_events = (int32_t*)(pair_ptr + (n + 1));
    int32_t* read_len = n_events + (n + 1);
    int64_t sp = 0, ep = 0, pp = 0;
    for (int32_t i = 0; i < n; i++) {
        seq_ptr[i] = sp; event_ptr[i] = ep; pair_ptr[i] = pp;
        read_len[i] = r->read_len[i];
        n_events[i] = r->n_events[i] > 0 ? r->n_events[i] : 0;
        sp += (int64_t)read_len[i] + 1; ep += n_events[i]; pp += (int64_t)n_events[i] + read_len[i];
    }
    if (host_reserve(c, c->h_rseq, (size_t)sp + 16)) return ABEA_ERR_CUDA;
    if (host_reserve(c, c->h_rmeans, (size_t)ep * sizeof(float) + 64)) return ABEA_ERR_CUDA;
    /* pair lists come back as path codes (stream_mode bit 2: 8 bytes per 32 pairs) or whole (bit 1); the staging for
     * whole lists is also what the copy-engine fall-backs below download into */
    const bool want_codes = (c->stream_mode & 4) != 0;
    if (want_codes && host_reserve(c, c->h_codes, (size_t)code_words(pp, n) * sizeof(abea_code_t))) return ABEA_ERR_CUDA;
    if (!want_codes && host_reserve(c, c->h_rpairs, (size_t)(pp + 1) * sizeof(abea_pair_t))) return ABEA_ERR_CUDA;
    if (host_reserve(c, c->h_rnp, ((size_t)n + 1) * sizeof(int32_t))) return ABEA_ERR_CUDA;
    char* h_seq = (char*)c->h_rseq.p;
    float* h_means = (float*)c->h_rmeans.p;
    abea_pair_t* h_pairs = (abea_pair_t*)c->h_rpairs.p; /* re-read after a late reservation (fall-backs) */
    const abea_code_t* h_codes = (const abea_code_t*)c->h_codes.p;
    volatile int32_t* h_np = (volatile int32_t*)c->h_rnp.p;
    for (int32_t i = 0; i < n; i++) h_np[i] = -1; /* "not done": the traceback stores the count when the list is out */

    abea_batch_t b;
    b.n_reads = n; b.seq = h_seq; b.seq_ptr = seq_ptr; b.read_len = read_len; b.events = nullptr;
    b.event_ptr = event_ptr; b.n_events = n_events; b.scalings = r->scalings; b.good = r->good; b.event_means = h_means;

    void* ev_alias = (c->stream_mode & 1) ? mapped_alias(h_means) : nullptr;
    abea_code_t* fin_codes = want_codes ? (abea_code_t*)mapped_alias(c->h_codes.p) : nullptr;
    abea_pair_t* fin_pairs = (!want_codes && (c->stream_mode & 2)) ? (abea_pair_t*)mapped_alias(h_pairs) : nullptr;
    int32_t* fin_np = (fin_pairs || fin_codes) ? (int32_t*)mapped_alias(c->h_rnp.p) : nullptr;
    const bool overlap = ev_alias && fin_np && n > 0 && ep > 0;

    /* workers: (1) sequences; (2) once the loader's work list exists, the means of its pieces in list order, each
     * published to the loader through its flag; (3) the finished pair lists, in the order the reads are expected to
     * finish, as their counts appear in the pinned count array */
    std::atomic<int32_t> seq_next(0), seq_left(n), item_next(0), out_next(0);
    std::atomic<int> abort_flag(0), all_packed(0);
    c->rag_items_ready.store(0);
    auto pack_read_means = [&](int32_t i, int32_t e0, int32_t e1) {
        const abea_event_t* src = r->events[i];
        float* dst = h_means + event_ptr[i];
        for (int32_t e = e0; e < e1; e++) dst[e] = src[e].mean;
    };
    double t_packed[64] = {0};
    std::atomic<int> kernels_done(0);
    const bool rag_late_copy = getenv("ABEA_RAG_LATE_COPY") != nullptr;
    const int rag_poll_us = getenv("ABEA_RAG_POLL_US") ? atoi(getenv("ABEA_RAG_POLL_US")) : 0;
    auto rag_pause = [](int us) {
        if (us <= 0) { sched_yield(); return; }
        struct timespec ts = {0, (long)us * 1000L};
        nanosleep(&ts, nullptr);
    };
    auto worker = [&](int tid_) {
        for (;;) { /* (1) */
            const int32_t i0 = seq_next.fetch_add(16);
            if (i0 >= n) break;
            const int32_t i1 = std::min(n, i0 + 16);
            for (int32_t i = i0; i < i1; i++) {
                memcpy(h_seq + seq_ptr[i], r->seq[i], (size_t)read_len[i]);
                h_seq[seq_ptr[i] + read_len[i]] = 0;
                if (!overlap) pack_read_means(i, 0, n_events[i]);
            }
            seq_left.fetch_sub(i1 - i0);
        }
        if (!overlap) return;
        while (c->rag_items_ready.load(std::memory_order_acquire) <= 0) { /* (2) */
            if (abort_flag.load()) return;
            rag_pause(0);
        }
        volatile uint32_t* flags = (volatile uint32_t*)c->h_hostready.p;
        const int64_t total_bytes = c->event_bytes, piece = c->load_piece_cur;
        for (;;) {
            const int32_t it = item_next.fetch_add(1);
            bool none_left = false;
            while (it >= c->rag_items_ready.load(std::memory_order_acquire)) { /* the list grows once (finish_load_order) */
                if (c->rag_items_final.load(std::memory_order_acquire) && it >= c->rag_items_ready.load(std::memory_order_acquire)) { none_left = true; break; }
                if (abort_flag.load()) return;
                rag_pause(0);
            }
            if (none_left) break;
            const abea_load_item_t item = c->items[(size_t)it];
            const abea_read_t& rd = c->reads[(size_t)item.read];
            const abea_load_geom_t g = abea_load_geom(rd.ev_off, rd.n_events, total_bytes, piece, (int64_t)sizeof(float));
            const int32_t e0 = (int32_t)abea_piece_first_event(g, item.piece, (int64_t)sizeof(float));
            const int32_t e1 = item.piece + 1 < g.n_pieces ? (int32_t)abea_piece_first_event(g, item.piece + 1, (int64_t)sizeof(float)) : rd.n_events;
            pack_read_means(rd.orig_index, e0, std::min(e1, rd.n_events));
            std::atomic_thread_fence(std::memory_order_release);
            flags[it] = 1u;
        }
        all_packed.fetch_add(1);
        t_packed[tid_ & 63] = now_ms();
        if (rag_late_copy) /* experiment: no copy-out while the kernels run */
            while (!kernels_done.load() && !abort_flag.load()) rag_pause(50);
        const int32_t n_sched = (int32_t)c->finish_order.size(); /* (3) */
        for (;;) {
            const int32_t j = out_next.fetch_add(1);
            if (j >= n_sched) break;
            const int32_t i = c->reads[(size_t)c->finish_order[(size_t)j]].orig_index;
            int32_t np;
            while ((np = h_np[i]) < 0) {
                if (abort_flag.load()) return;
                rag_pause(rag_poll_us);
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            r->n_pairs[i] = np;
            if (np > 0 && r->pairs[i]) {
                if (fin_codes) decode_codes(h_codes + abea_code_offset(pair_ptr[i], i), np, r->pairs[i]);
                else memcpy(r->pairs[i], h_pairs + pair_ptr[i], (size_t)np * sizeof(abea_pair_t));
            }
        }
    };
    c->pool.kick(threads, worker);
    while (seq_left.load() > 0) sched_yield();
    const double t1 = now_ms();

    int rc = ABEA_OK;
    if (overlap) {
        rc = upload_impl(c, &b, ev_alias, nullptr, true);
        if (rc == ABEA_OK && !c->streaming) rc = fail(c, ABEA_ERR_STATE, "ragged batch was not streamed");
        if (rc == ABEA_OK) rc = run_impl(c, fin_pairs, fin_np, nullptr, fin_codes);
        if (rc != ABEA_OK) abort_flag.store(1);
        kernels_done.store(1);
        const double t_run = now_ms();
        c->pool.wait();
        if (getenv("ABEA_TIME_PACK")) {
            double tp = 0;
            for (int i = 0; i < threads && i < 64; i++) tp = std::max(tp, t_packed[i]);
            fprintf(stderr, "[abea ragged] layout+seq %.3f ms, kernels done at %.3f ms, means packed at %.3f ms, unpacked at %.3f ms (%d threads, %zu items)\n",
                    t1 - t0, t_run - t0, tp - t0, now_ms() - t0, threads, (size_t)c->n_items);
        }
        if (rc != ABEA_OK && all_packed.load() == threads) {
            /* e.g. a stalled stream: the batch is complete in the staging, run it once more through the copy engine */
            rc = upload_impl(c, &b, nullptr, nullptr);
            if (rc == ABEA_OK) rc = run_impl(c, nullptr, nullptr, nullptr);
            if (rc == ABEA_OK && host_reserve(c, c->h_rpairs, (size_t)(pp + 1) * sizeof(abea_pair_t))) rc = ABEA_ERR_CUDA;
            h_pairs = (abea_pair_t*)c->h_rpairs.p;
            if (rc == ABEA_OK) rc = abea_download(c, h_pairs, pair_ptr, (int32_t*)c->h_rnp.p, nullptr);
            if (rc == ABEA_OK)
                for (int32_t i = 0; i < n; i++) {
                    r->n_pairs[i] = h_np[i];
                    if (h_np[i] > 0 && r->pairs[i]) memcpy(r->pairs[i], h_pairs + pair_ptr[i], (size_t)h_np[i] * sizeof(abea_pair_t));
                }
        }
        if (rc != ABEA_OK) return rc;
        for (int32_t i = 0; i < n; i++) /* reads the filter kept out of the schedule */
            if (c->sreads[(size_t)i].sched < 0) r->n_pairs[i] = 0;
    } else {
        c->pool.wait();
        rc = upload_impl(c, &b, nullptr, nullptr);
        if (rc == ABEA_OK) rc = run_impl(c, nullptr, nullptr, nullptr);
        if (rc == ABEA_OK && host_reserve(c, c->h_rpairs, (size_t)(pp + 1) * sizeof(abea_pair_t))) rc = ABEA_ERR_CUDA;
        h_pairs = (abea_pair_t*)c->h_rpairs.p;
        if (rc == ABEA_OK && n > 0) rc = abea_download(c, h_pairs, pair_ptr, (int32_t*)c->h_rnp.p, nullptr);
        if (rc != ABEA_OK) return rc;
        std::atomic<int32_t> nx(0);
        c->pool.kick(threads, [&](int) {
            for (;;) {
                const int32_t i = nx.fetch_add(1);
                if (i >= n) break;
                r->n_pairs[i] = h_np[i];
                if (h_np[i] > 0 && r->pairs[i]) memcpy(r->pairs[i], h_pairs + pair_ptr[i], (size_t)h_np[i] * sizeof(abea_pair_t));
            }
        });
        c->pool.wait();
    }
    c->last.pack_ms += t1 - t0;
    c->last.ragged_ms = now_ms() - t0;
    if (overlap) {
        int64_t np = 0;
        for (int32_t i = 0; i < n; i++) np += r->n_pairs[i];
        c->last.d2h_ms = 0.f;
        c->last.unpack_ms = 0.0;
        int64_t cb = 0;
        if (fin_codes)
            for (int32_t i = 0; i < n; i++) cb += code_bytes_used(r->n_pairs[i]);
        c->last.d2h_bytes = (fin_codes ? cb : np * (int64_t)sizeof(abea_pair_t)) + (int64_t)n * (int64_t)sizeof(int32_t);
    }
    if (timing) *timing = c->last;
    return ABEA_OK;
}

} /* extern "C" */
