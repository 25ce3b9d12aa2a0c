/* abea_kernels.cuh — sm_100a device code for adaptive banded event alignment (ABEA).
 *
 * Replaces the reference's three kernels (align_kernel_pre_2d / align_kernel_core_2d_shm / align_kernel_post,
 * reference src/align.cu:149-749) with a different decomposition:
 *
 *   abea_prepare_kernel      one thread per k-mer: rank -> model gather -> {scaled mean, stdv, -0.918938-log stdv,
 *                            1/stdv} (the reference does this with ONE thread per read, src/align.cu:203-209), plus a
 *                            range check of all inputs that selects the fast or the exact arithmetic per read
 *   abea_fill_kernel         one WARP per read, 4 band cells per lane, whole band state in registers; neighbours
 *                            through warp shuffles; 2-bit packed trace, one coalesced 128-B store per 4 bands;
 *                            last-column arg-max folded into the fill; persistent warps pulling reads longest-first;
 *                            traceback + QC of a read fused in, by the warp that filled it (abea_traceback_par:
 *                            128 walks per warp over 128 stretches of the path, speculative entries verified)
 *   abea_fill_wide_kernel    one CTA of four warps per read, one band cell per lane, for the reads that set the makespan
 *   abea_load_kernel         streams a pinned host batch in, in the order the fill will ask for it
 *
 * Arithmetic contract (bit-exact against the reference CPU align(), src/align.c:180-559):
 * emission in float with explicit round-to-nearest intrinsics (no FMA contraction), the three transition sums in
 * double rounded once to float, ties L > U > D, per-read lp_stay/lp_step computed on the HOST in double.
 *
 * The same source is compiled by tests/simt/ (a CPU lock-step emulator, test infrastructure) when
 * ABEA_SIMT_EMU is defined; nothing in the product path depends on that.
 */
#pragma once

#include <stdint.h>

#include "../../include/abea_types.h"

#ifndef ABEA_SIMT_EMU
#include <cuda_runtime.h>
#endif

#define ABEA_W 100            /* ALN_BANDWIDTH, reference src/f5c.h:34 */
#define ABEA_CPL 4            /* band cells per lane */
#define ABEA_LANES 25         /* lanes that own band cells (25*4 = 100) */
#define ABEA_FULL 0xffffffffu
#define ABEA_TRACE_GROUP_WORDS 32 /* one 128-B line per 4 bands: words 0..24 trace, 25..28 band event index */

#define ABEA_READ_FAST 1u /* read_flags bit: all inputs of the read are in the range the fast arithmetic is exact on */

#define ABEA_FROM_D 0u /* reference src/align.c:194-196 */
#define ABEA_FROM_U 1u
#define ABEA_FROM_L 2u

/* Per-read descriptor built by the host packer (abea_host.cu), in scheduling (longest-first) order. */
struct abea_read_t {
    int64_t seq_off;    /* first base in d_seq */
    int64_t ev_off;     /* first event mean in d_means (flat float array: only .mean of event_t is ever read, reference src/align.cu:415) */
    int64_t evs_off;    /* running sum of n_events over the schedule (flat index space of abea_prepare_kernel) */
    int64_t kp_off;     /* first k-mer in d_kparams */
    int64_t trace_off;  /* first 32-bit word of this read's trace in d_trace */
    int64_t pair_off;   /* first pair slot of the read in d_pairs == in the caller's canonical layout (capacity pair_cap) */
    double lp_stay;     /* log(p_stay), host double (reference src/align.c:214) */
    double lp_step;     /* log(1 - exp(lp_skip) - exp(lp_stay)) (src/align.c:215) */
    float scale;
    float shift;
    int32_t n_events;   /* E */
    int32_t n_kmers;    /* K = L - k + 1 */
    int32_t pair_cap;   /* E + L (reference src/f5c.c:724-726) */
    int32_t orig_index; /* index of the read in the caller's batch */
};

/* Per-read result. */
struct abea_result_t {
    double sum_emission; /* double sum of float emissions, traceback order (src/align.c:476) */
    float end_score;     /* best last-column score incl. trailing trim */
    int32_t end_event;   /* event the traceback starts from */
    int32_t n_aligned;   /* pairs before QC */
    int32_t n_pairs;     /* pairs after QC (0 = failed) */
    int32_t start_us;    /* %globaltimer (microseconds, low 31 bits) when the fill of this read began (for profiles/) */
    int32_t max_gap;
    int32_t wide;        /* 1 if the wide kernel filled this read */
    int32_t respec;      /* parallel traceback: segments whose speculative entry cell was wrong and that were walked again */
    int64_t fill_cycles; /* SM clock cycles this read spent in the band fill ... */
    int64_t trace_cycles;/* ... and in traceback + QC (per-read latency, for profiles/) */
};

__device__ __forceinline__ int32_t abea_now_us() {
#ifdef ABEA_SIMT_EMU
    return 0;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return (int32_t)((t / 1000ull) & 0x7fffffffull);
#endif
}

__device__ __forceinline__ long long abea_clock() {
#ifdef ABEA_SIMT_EMU
    return 0;
#else
    return clock64();
#endif
}

struct abea_consts_t {
    double lp_skip; /* log(1e-10) (src/align.c:212-213) */
    double lp_trim; /* log(0.01)  (src/align.c:216) */
};

/* ------------------------------------------------------------------------------------------------------------ */

/* A,C,G,T -> 0..3; anything else -> 0 (reference src/align.c:19-32 / src/align.cu:21-33) */
__device__ __forceinline__ uint32_t abea_base_rank(uint8_t b) {
    return b == 'C' ? 1u : (b == 'G' ? 2u : (b == 'T' ? 3u : 0u));
}

/* Emission log-probability, float, no contraction (reference src/align.c:108-115,137-152).
 * kp = {scale*level_mean+shift, level_stdv, -0.918938f - level_log_stdv, unused}. */
__device__ __forceinline__ float abea_emission(float x, float kp_mean, float kp_stdv, float kp_lead) {
    float a = __fdiv_rn(__fsub_rn(x, kp_mean), kp_stdv);
    return __fadd_rn(kp_lead, __fmul_rn(__fmul_rn(-0.5f, a), a));
}

/* One DP cell (reference src/align.c:378-392): double sums rounded once, ties L > U > D. */
__device__ __forceinline__ void abea_cell(float lp, float up, float left, float diag, double lp_step,
                                          double lp_stay, double lp_skip, float& score, uint32_t& from) {
    double lpd = (double)lp;
    float sd = __double2float_rn(__dadd_rn(__dadd_rn((double)diag, lp_step), lpd));
    float su = __double2float_rn(__dadd_rn(__dadd_rn((double)up, lp_stay), lpd));
    float sl = __double2float_rn(__dadd_rn((double)left, lp_skip));
    float m = sd;
    uint32_t f = ABEA_FROM_D;
    m = su > m ? su : m;
    f = (m == su) ? ABEA_FROM_U : f;
    m = sl > m ? sl : m;
    f = (m == sl) ? ABEA_FROM_L : f;
    score = m;
    from = f;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Per-read preparation: (1) k-mer parameter cache kparams[kp_off + i] = {scale*level_mean+shift, level_stdv,
 * -0.918938f - level_log_stdv, RN(1/level_stdv)} for k-mer i of each read (the reference builds this cache with ONE
 * thread per read, src/align.cu:203-209); (2) range validation of every value the fast arithmetic of the fill
 * kernel touches — a read keeps ABEA_READ_FAST in read_flags only if all of its event means and scaled level means
 * are 0 or within [2^-60, 2^16] in magnitude (so that x - mean is 0 or at least 2^-83: every intermediate of the
 * quotient correction stays exactly representable; tools/validate_fast_arith.c checks the whole admitted range against
 * the plain IEEE expressions) and all stdv are within [2^-6, 2^12] with a mantissa that is not all
 * ones (the one case Markstein's quotient correction excludes). read_flags must be pre-set to ABEA_READ_FAST.     */

__device__ __forceinline__ bool abea_sane_level(float v) {
    float a = fabsf(v);
    return (a == 0.0f) || (a >= 8.6736174e-19f /* 2^-60 */ && a <= 65536.0f); /* NaN fails both */
}
__device__ __forceinline__ bool abea_sane_stdv(float v) {
    return (v >= 0.015625f) && (v <= 4096.0f) && ((__float_as_uint(v) & 0x007fffffu) != 0x007fffffu);
}

__device__ __forceinline__ int32_t abea_find_read(const abea_read_t* __restrict__ reads, int32_t n_reads, int64_t idx,
                                                  bool by_events) {
    int32_t lo = 0, hi = n_reads - 1;
    while (lo < hi) {
        int32_t mid = (lo + hi + 1) >> 1;
        int64_t off = by_events ? reads[mid].evs_off : reads[mid].kp_off;
        if (off <= idx) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void abea_prepare_kernel(const abea_read_t* __restrict__ reads, int32_t n_reads,
                                    const uint8_t* __restrict__ seq, const abea_model_t* __restrict__ model,
                                    uint32_t kmer_size, const float* __restrict__ means,
                                    float4* __restrict__ kparams, uint32_t* __restrict__ read_flags,
                                    int64_t total_kmers, int64_t total_events) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t idx = tid; idx < total_kmers; idx += stride) {
        /* reads are laid out in kp_off order: find the read owning flat k-mer idx */
        int32_t r = abea_find_read(reads, n_reads, idx, false);
        const abea_read_t rd = reads[r];
        int64_t i = idx - rd.kp_off;
        const uint8_t* s = seq + rd.seq_off + i;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < kmer_size; j++) rank = (rank << 2) | abea_base_rank(s[j]);
        abea_model_t m = model[rank];
        float4 kp;
        kp.x = __fadd_rn(__fmul_rn(rd.scale, m.level_mean), rd.shift); /* src/align.c:137-138 */
        kp.y = m.level_stdv;
        kp.z = __fsub_rn(-0.918938f, m.level_log_stdv);               /* src/align.c:111-113 */
        kp.w = __frcp_rn(m.level_stdv);
        kparams[idx] = kp;
        if (!(abea_sane_level(kp.x) && abea_sane_stdv(kp.y))) atomicAnd(&read_flags[r], ~ABEA_READ_FAST);
    }
    /* events of scheduled reads, in schedule order (evs_off = running sum of n_events) */
    for (int64_t idx = tid; idx < total_events; idx += stride) {
        int32_t r = abea_find_read(reads, n_reads, idx, true);
        const abea_read_t rd = reads[r];
        float x = means[rd.ev_off + (idx - rd.evs_off)];
        if (!abea_sane_level(x)) atomicAnd(&read_flags[r], ~ABEA_READ_FAST);
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Streaming (abea_align_batch with pinned, mapped host buffers). The copy engine would bring the events of a batch in
 * the caller's order and the kernels could only start when the last byte has landed; instead a few CTAs of
 * abea_load_kernel read the events straight out of the caller's pinned buffer over PCIe — in the order the fill
 * warps are going to ask for them — and publish a per-read counter of landed pieces that a fill warp (or wide CTA)
 * waits on before it touches the read. The results go the other way without a copy either: the last step of the
 * traceback writes the finished pair list — as path codes (below), or whole — to mapped host memory and publishes the
 * read's count behind it. With ready == NULL and no *_final pointer the kernels run on data that is already resident
 * and leave their results in device memory only (abea_upload_batch / abea_run / abea_download). */
/* Path codes: a finished pair list in 1/32 of its size. The list is a monotone lattice path — consecutive pairs differ
 * by (+1,+1), (0,+1) or (+1,0) in (ref_pos, read_pos), the traceback's D / U / L steps read forwards (reference
 * src/align.c:452-499) — so the first pair and one bit per coordinate per step say everything: word 0 of a read's
 * region holds the first pair, word 1 + j the steps 32 j .. 32 j + 31 as two bit planes (bit t of `a`: ref_pos advances
 * at step 32 j + t; bit t of `b`: read_pos does). 8 bytes cross PCIe for 32 pairs instead of 256, and host threads that
 * would otherwise only copy the list expand it while the kernels are still running (abea_host.cu: decode_codes).
 * Read i's region starts at word (pair_off >> 5) + 2 i of the buffer: 1 + ceil((E+L-1)/32) words always fit before the
 * next read's. */
typedef abea_code_word_t abea_code_t; /* include/abea_types.h */
__host__ __device__ __forceinline__ int64_t abea_code_offset(int64_t pair_off, int32_t batch_index) {
    return (pair_off >> 5) + 2 * (int64_t)batch_index;
}

struct abea_stream_t {
    const uint32_t* ready;      /* [scheduled read] pieces of its events landed so far; NULL: everything is resident */
    abea_pair_t* pairs_final;   /* the caller's mapped host buffer (canonical layout), or NULL: the lists stay in d_pairs only */
    int32_t* n_pairs_final;     /* [batch read] pair counts in the caller's mapped host buffer, or NULL */
    abea_code_t* codes_final;   /* mapped host buffer of path codes (below), or NULL */
    int32_t* n_pairs_dev;       /* [batch read] pair counts on the device (always written) */
    uint32_t* stalled;          /* set to 1 if a wait for streamed events gave up (the host reports an error) */
    int32_t tb_mode;            /* 0: serial traceback (one walk per warp), 1: segment-parallel traceback (a walk per lane);
                                 * bit 1 (tests): the parallel form always sums the emissions in order */
    int32_t tb_margin;          /* bands a speculative walk starts above its segment */
};

/* default smallest work item of the loader: 2048 events, i.e. 48 KB of an AoS event table or 8 KB of a flat array of
 * means; the host passes the value in use (ABEA_LOAD_PIECE_KB overrides it) to both sides */
#define ABEA_LOAD_PIECE_BYTES (48 * 1024)
#define ABEA_LOAD_PIECE_BYTES_MEANS (8 * 1024)
#define ABEA_LOAD_MAX_PIECES 64           /* per read: its landed pieces are a 64-bit mask (two words of d_ready) */
#define ABEA_LOAD_THREADS 128
#define ABEA_LOAD_UNROLL 8                /* 16-B loads in flight per thread */

struct abea_load_item_t {
    int32_t read;  /* scheduled index */
    int32_t piece; /* piece of that read's line-aligned byte range */
};

/* The byte range of a read's events in the caller's pinned array (elements of esz bytes: 4 for a flat array of event
 * means, sizeof(abea_event_t) for the reference's AoS table), widened to whole 128-B lines (clamped to the buffer): a line
 * shared by two reads is copied for both, so whichever is published first the line is complete. The range is cut into
 * at most 64 pieces of at least piece_min bytes (whole lines). */
struct abea_load_geom_t {
    int64_t a, b;   /* the read's own bytes [a, b) */
    int64_t lo, hi; /* widened to lines */
    int64_t piece;  /* bytes per piece */
    int32_t n_pieces;
};
__device__ __host__ __forceinline__ abea_load_geom_t abea_load_geom(int64_t ev_off, int32_t n_events, int64_t total_bytes,
                                                                     int64_t piece_min, int64_t esz) {
    abea_load_geom_t g;
    g.a = ev_off * esz;
    g.b = g.a + (int64_t)n_events * esz;
    const int64_t h = (g.b + 127) & ~(int64_t)127;
    g.lo = g.a & ~(int64_t)127;
    g.hi = h < total_bytes ? h : total_bytes;
    int64_t per = (((g.hi - g.lo) + ABEA_LOAD_MAX_PIECES - 1) / ABEA_LOAD_MAX_PIECES + 127) & ~(int64_t)127;
    g.piece = per > piece_min ? per : piece_min;
    g.n_pieces = (int32_t)((g.hi - g.lo + g.piece - 1) / g.piece);
    return g;
}
/* first event of the read that lies (at least partly) in piece p */
__device__ __host__ __forceinline__ int64_t abea_piece_first_event(const abea_load_geom_t& g, int32_t p, int64_t esz) {
    const int64_t off = g.lo + (int64_t)p * g.piece - g.a;
    return off <= 0 ? 0 : off / esz;
}

__device__ __forceinline__ uint32_t abea_ld_acquire_u32(const uint32_t* p) {
#ifdef ABEA_SIMT_EMU
    return *p;
#else
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}

__device__ __forceinline__ uint32_t abea_ld_relaxed_u32(const uint32_t* p) {
#ifdef ABEA_SIMT_EMU
    return *p;
#else
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}

/* Streaming: per scheduled read, d_ready holds four words: [0] how many leading events of the read have landed in
 * d_events (0x7fffffff once all have), [2..3] the 64-bit mask of landed pieces it is derived from (loader only).
 * abea_wait_landed_events blocks until event `last` (and every event before it) has landed and returns the count it
 * saw, so that the caller asks again only when it runs past that: a fill warp starts a read as soon as its first
 * piece is in and chases the loader from there.
 * Called by ALL lanes of a warp (all threads of a wide CTA), warp-uniformly: every lane polls the same word (one
 * broadcast transaction). A single spinning lane would leave the warp split after the loop — the convergence barrier
 * around a NANOSLEEP loop is not a reconverging one — and a split warp executes the whole read at a quarter of the
 * speed (measured: every read that had waited ran at 4000 instead of 1000 cycles per band). The polling loads are
 * RELAXED, with one acquire at the end: an acquire load at gpu scope is LDG.STRONG + CCTL.IVALL (the SM's whole L1). */
#define ABEA_READY_WORDS 4
#define ABEA_WAIT_SPINS_MAX (1 << 22) /* x >= 1 us: a loader that has not delivered in seconds never will */
__device__ __forceinline__ int32_t abea_wait_landed_events(const uint32_t* w, int32_t last, uint32_t* stalled) {
    uint32_t v;
    for (int spins = 0;; spins++) {
        v = abea_ld_relaxed_u32(w);
        /* the vote makes the exit warp-uniform by construction */
        if (!__any_sync(ABEA_FULL, v <= (uint32_t)last)) break;
        /* never hang the GPU: flag the batch as failed and carry on (at once if another warp already gave up) */
        if (spins >= ABEA_WAIT_SPINS_MAX || ((spins & 255) == 255 && *(volatile uint32_t*)stalled != 0u)) {
            *stalled = 1u;
            v = 0x7fffffffu;
            break;
        }
#ifndef ABEA_SIMT_EMU
        __nanosleep(1000);
#else
        break; /* the emulator runs kernels one after the other: the loader has finished */
#endif
    }
    (void)abea_ld_acquire_u32(w);
    return (int32_t)v;
}

/* src: the caller's pinned (mapped) host array, 16-B aligned; dst: d_means, the flat array of event means the
 * alignment kernels read (same event indexing as the source). AOS = false: the source is a flat array of event means
 * (abea_batch_t.event_means) — 4 bytes per event cross PCIe; AOS = true: the source is the reference's event table
 * (24-B event_t, reference src/f5c.h:129-136), of which only .mean is kept: in a 16-B unit u of the table the mean is
 * .w when u % 3 == 0 (event 2(u/3)) and .y when u % 3 == 2 (event 2(u/3)+1). The means pass through here, so this is also
 * where they are range-checked for the fast arithmetic (the resident path does that in abea_prepare_kernel).
 * host_ready (or NULL): per scheduled read, a word in mapped host memory that the packer threads of abea_align_ragged
 * set once the read's means are in the pinned array — a piece is copied only after its read has been packed. */
template <bool AOS>
__global__ void __launch_bounds__(ABEA_LOAD_THREADS)
abea_load_kernel(const abea_read_t* __restrict__ reads, const abea_load_item_t* __restrict__ items, int32_t n_items,
                 const uint4* __restrict__ src, float* __restrict__ dst, int64_t total_bytes,
                 uint32_t* __restrict__ read_flags, uint32_t* __restrict__ ready, int32_t* __restrict__ counter,
                 int64_t piece_min, const volatile uint32_t* host_ready, uint32_t* stalled) {
    __shared__ int s_item;
    const int tid = threadIdx.x;
    const int64_t esz = AOS ? (int64_t)sizeof(abea_event_t) : (int64_t)sizeof(float);
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(counter, 1);
        __syncthreads();
        const int it = s_item;
        if (it >= n_items) break;
        const abea_load_item_t item = items[it];
        const abea_read_t rd = reads[item.read];
        if (host_ready) { /* the host packs the pieces in the order of this work list; normally it is far ahead */
            if (tid == 0) {
                int spins = 0;
                while (host_ready[it] == 0u) {
                    if (++spins > ABEA_WAIT_SPINS_MAX) {
                        *stalled = 1u;
                        break;
                    }
#ifndef ABEA_SIMT_EMU
                    __nanosleep(500);
#else
                    sched_yield(); /* the emulator runs this kernel on the caller's thread; the packers are real threads */
                    spins = 0;
#endif
                }
            }
            __syncthreads();
        }
        /* with a host that is still packing, a line shared with a neighbouring read may hold that read's means or not
         * yet: only the read's own means are written (the neighbour's own pieces bring its) */
        const bool own_only = host_ready != nullptr;
        const abea_load_geom_t g = abea_load_geom(rd.ev_off, rd.n_events, total_bytes, piece_min, esz);
        const int64_t a = g.a, b = g.b;
        const int64_t p0 = g.lo + (int64_t)item.piece * g.piece;
        const int64_t p1 = (p0 + g.piece < g.hi) ? p0 + g.piece : g.hi;
        const int64_t u1 = p1 >> 4;
        bool bad = false;
        for (int64_t u = (p0 >> 4) + tid; u < u1; u += ABEA_LOAD_THREADS * ABEA_LOAD_UNROLL) {
            uint4 v[ABEA_LOAD_UNROLL];
#pragma unroll
            for (int j = 0; j < ABEA_LOAD_UNROLL; j++) {
                const int64_t uj = u + (int64_t)j * ABEA_LOAD_THREADS;
                if (uj < u1) v[j] = src[uj];
            }
#pragma unroll
            for (int j = 0; j < ABEA_LOAD_UNROLL; j++) {
                const int64_t uj = u + (int64_t)j * ABEA_LOAD_THREADS;
                if (uj < u1) {
                    if (AOS) {
                        const int m3 = (int)(uj % 3);
                        if (m3 != 1) {
                            const int64_t mb = (uj << 4) + (m3 == 0 ? 12 : 4);
                            const float x = __uint_as_float(m3 == 0 ? v[j].w : v[j].y);
                            dst[2 * (uj / 3) + (m3 == 0 ? 0 : 1)] = x;
                            if (mb >= a && mb < b && !abea_sane_level(x)) bad = true;
                        }
                    } else {
                        const int64_t mb = uj << 4;
                        const uint32_t w4[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                        const bool whole = (mb >= a) && (mb + 16 <= b);
                        if (whole || !own_only) ((uint4*)dst)[uj] = v[j];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const bool mine = (mb + 4 * q >= a) && (mb + 4 * q < b);
                            if (mine && !whole && own_only) dst[(mb >> 2) + q] = __uint_as_float(w4[q]);
                            if (mine && !abea_sane_level(__uint_as_float(w4[q]))) bad = true;
                        }
                    }
                }
            }
        }
        if (!AOS && tid < 3) { /* the array ends on part of a unit: its last one to three means */
            const int64_t f = (u1 << 2) + tid;
            if ((f << 2) < p1) {
                const float x = ((const float*)src)[f];
                const bool mine = ((f << 2) >= a) && ((f << 2) < b);
                if (mine || !own_only) dst[f] = x;
                if (mine && !abea_sane_level(x)) bad = true;
            }
        }
        if (bad) atomicAnd(&read_flags[item.read], ~ABEA_READ_FAST);
        __threadfence();
        __syncthreads();
        if (tid == 0) { /* mark the piece, then publish the read's landed prefix (in events) */
            uint32_t* w = ready + (int64_t)ABEA_READY_WORDS * item.read;
            const int wi = item.piece >> 5;
            const uint32_t mine = atomicOr(&w[2 + wi], 1u << (item.piece & 31)) | (1u << (item.piece & 31));
            const uint32_t other = atomicOr(&w[2 + (wi ^ 1)], 0u);
            const uint32_t m0 = wi == 0 ? mine : other, m1 = wi == 0 ? other : mine;
            const int32_t np = (m0 != 0xffffffffu) ? (__ffs((int)~m0) - 1) : ((m1 != 0xffffffffu) ? 32 + (__ffs((int)~m1) - 1) : 64);
            const int64_t end = g.lo + (int64_t)np * g.piece;
            uint32_t nev = 0x7fffffffu;
            if (end < g.b && np < g.n_pieces) nev = end <= g.a ? 0u : (uint32_t)((end - g.a) / esz);
            __threadfence(); /* pieces whose bits were observed above are ordered before the count */
            atomicMax(&w[0], nev);
        }
    }
}

/* Event means out of a device-resident event table (the copy-engine upload of an AoS batch; the tables abea_getevents
 * leaves on the device): means[i] = events[i].mean over the whole index space, one thread per event. */
__global__ void abea_extract_means_kernel(const abea_event_t* __restrict__ events, float* __restrict__ means, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) means[i] = events[i].mean;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Band fill. One warp per read; lane j owns band offsets 4j..4j+3 (lanes 25..31 own nothing and only help with
 * loads and the trace line). Band b has lower-left (eb, kb); cell at offset o is (event eb-o, k-mer kb+o)
 * (reference src/align.c:156-161).
 *
 * Arithmetic. A band score is a float in the reference; here it is carried as the DOUBLE that has
 * exactly that float's value ("float-valued double"), so the three transition sums need no f32<->f64 conversion
 * (F2F runs on the XU pipe at 16 lanes/clk/SM on sm_100 — measured, profiles/microbench_r01.txt). Rounding a
 * double sum to float precision is done in the FP64 pipe: add and subtract C = sign(x)*1.5*2^(e+29), e = exponent
 * of x, which rounds to nearest-even at float's 24 bits exactly like cvt.rn.f32.f64 for every x whose float image
 * is normal, zero or infinite. The emission quotient (x-mean)/stdv uses the host-rounded reciprocal and two FMAs
 * (Markstein's correction), which is the correctly rounded IEEE quotient when operands are in a sane range.
 * Both shortcuts are bit-exact only on validated inputs: abea_prepare_kernel range-checks every event mean, scaled
 * level mean and stdv of a read and clears ABEA_READ_FAST otherwise; such reads take the EXACT instantiation
 * (hardware conversions, __fdiv_rn), the same arithmetic the first version of this kernel used everywhere.        */

__device__ __forceinline__ double abea_neg_inf_d() { return __hiloint2double((int)0xfff00000, 0); }

__device__ __forceinline__ float abea_load_event_mean(const float* __restrict__ ev, int32_t e, int32_t E) {
    e = e < 0 ? 0 : (e >= E ? E - 1 : e);
    return ev[e];
}

__device__ __forceinline__ float4 abea_load_kparam(const float4* __restrict__ kp, int32_t k, int32_t K) {
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    return kp[k];
}

/* x rounded to float precision, returned as a double. FAST: FP64-pipe magic constant; else hardware conversions.
 * `donor` is any float-valued double (or +-inf): its low word (29 zero bits below at most 3 mantissa bits) becomes the
 * low word of C, which keeps C an even multiple of the float ulp — so C needs no register of its own zeroed. */
template <bool FAST>
__device__ __forceinline__ double abea_round_f32(double x, double donor) {
    if (FAST) {
        int hi = __double2hiint(x);
        /* (hi & 0xfff00000) + 0x01d80000 written as shift + multiply-add: one ALU-pipe op and one FMA-pipe op
         * instead of two ALU-pipe ops (the ALU pipe is the busiest pipe of this kernel, profiles/) */
        unsigned chi;
#ifdef ABEA_SIMT_EMU
        chi = ((unsigned)hi >> 20) * 0x00100000u + 0x01d80000u;
#else
        asm("{\n\t.reg .u32 t;\n\tshr.u32 t, %1, 20;\n\tmad.lo.u32 %0, t, 1048576, 30932992;\n\t}" : "=r"(chi) : "r"(hi));
#endif
        double C = __hiloint2double((int)chi, __double2loint(donor));
        return __dadd_rn(__dadd_rn(x, C), -C);
    } else {
        return (double)__double2float_rn(x);
    }
}

/* emission log-probability (reference src/align.c:108-115,137-152); kp = {mean', stdv, lead, 1/stdv} */
template <bool FAST>
__device__ __forceinline__ float abea_emission_t(float x, const float4& kp) {
    if (FAST) {
        float t = __fsub_rn(x, kp.x);
        float q0 = __fmul_rn(t, kp.w);
        float rem = __fmaf_rn(-kp.y, q0, t);
        float a = __fmaf_rn(rem, kp.w, q0);               /* == RN(t / stdv) */
        return __fmaf_rn(__fmul_rn(a, a), -0.5f, kp.z);    /* == lead + ((-0.5f*a)*a), scaling by -0.5 is exact */
    } else {
        return abea_emission(x, kp.x, kp.y, kp.z);
    }
}

/* One DP cell on float-valued doubles (reference src/align.c:378-392): the three sums are formed in double,
 * each rounded once to float precision, then compared with ties L > U > D. */
template <bool FAST>
__device__ __forceinline__ void abea_cell_dd(double lpd, double up, double left, double diag, double lp_step,
                                             double lp_stay, double lp_skip, double& score, uint32_t& from) {
    double rd = abea_round_f32<FAST>(__dadd_rn(__dadd_rn(diag, lp_step), lpd), diag);
    double ru = abea_round_f32<FAST>(__dadd_rn(__dadd_rn(up, lp_stay), lpd), up);
    double rl = abea_round_f32<FAST>(__dadd_rn(left, lp_skip), left);
    bool isU = ru >= rd;          /* (su > max) || (max == su) */
    double m = isU ? ru : rd;
    bool isL = rl >= m;
    score = isL ? rl : m;
    from = isL ? ABEA_FROM_L : (isU ? ABEA_FROM_U : ABEA_FROM_D);
}
/* The cell as the narrow kernel computes it: the three sums in double, each narrowed to float by the hardware
 * conversion (cvt.rn.f32.f64 — on sm_100 the DOWN conversion runs at the FP64 pipe's rate, 2 warp-ops/clk/SM; it is the
 * UP conversion that crawls through the XU pipe at 0.5, profiles/microbench_r01.txt), then the reference's own float
 * logic (src/align.c:386-392): the score is the float maximum, `from` is L if the left candidate equals it, else U if
 * the up candidate does, else D — ties L > U > D. One up-conversion turns the score into the float-valued double the
 * next bands add to. 15 instructions per cell (5 DADD, 3 + 1 F2F, 2 FMNMX, 2 FSETP, 2 SEL) against 28 for the form that
 * rounds inside the FP64 pipe (abea_cell_dd<true>, still the wide kernel's: its per-band latency chain is shorter
 * without the conversions) — the opcode mixes are in profiles/fill_narrow_opcode_mix_*_r02.txt. No range assumption:
 * these ARE the reference's operations. */
__device__ __forceinline__ void abea_cell_f32(double lpd, double up, double left, double diag, double lp_step,
                                              double lp_stay, double lp_skip, double& score, uint32_t& from) {
    const float sd = __double2float_rn(__dadd_rn(__dadd_rn(diag, lp_step), lpd));
    const float su = __double2float_rn(__dadd_rn(__dadd_rn(up, lp_stay), lpd));
    const float sl = __double2float_rn(__dadd_rn(left, lp_skip));
    const float m = fmaxf(fmaxf(sd, su), sl); /* no NaN ever: the candidates are finite or -inf */
    score = (double)m;
    from = (sl == m) ? ABEA_FROM_L : ((su == m) ? ABEA_FROM_U : ABEA_FROM_D);
}

template <bool FAST>
__device__ __forceinline__ void abea_cell_d(float lp, double up, double left, double diag, double lp_step,
                                            double lp_stay, double lp_skip, double& score, uint32_t& from) {
    abea_cell_dd<FAST>((double)lp, up, left, diag, lp_step, lp_stay, lp_skip, score, from);
}

/* the wide kernel's cell: ABEA_WIDE_CELL_F32 selects the float-maximum form for an experiment build */
template <bool FAST>
__device__ __forceinline__ void abea_wide_cell(double lpd, double up, double left, double diag, double lp_step,
                                               double lp_stay, double lp_skip, double& score, uint32_t& from) {
#ifdef ABEA_WIDE_CELL_F32
    abea_cell_f32(lpd, up, left, diag, lp_step, lp_stay, lp_skip, score, from);
#else
    abea_cell_dd<FAST>(lpd, up, left, diag, lp_step, lp_stay, lp_skip, score, from);
#endif
}

/* cp.async helpers (LDGSTS on sm_100a); the CPU emulator copies synchronously */
__device__ __forceinline__ void abea_cp_async4(void* dst_smem, const void* src_gmem) {
#ifdef ABEA_SIMT_EMU
    *(uint32_t*)dst_smem = *(const uint32_t*)src_gmem;
#else
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(src_gmem) : "memory");
#endif
}
__device__ __forceinline__ void abea_cp_async16(void* dst_smem, const void* src_gmem) {
#ifdef ABEA_SIMT_EMU
    for (int i = 0; i < 4; i++) ((uint32_t*)dst_smem)[i] = ((const uint32_t*)src_gmem)[i];
#else
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src_gmem) : "memory");
#endif
}
__device__ __forceinline__ void abea_cp_async_wait_all() {
#ifndef ABEA_SIMT_EMU
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
#endif
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Traceback + QC (reference src/align.c:452-543). One warp per read. The walk is a pointer chase through the packed
 * trace, strictly downwards in band index, so the trace lines it will need are known in advance: they are streamed
 * with cp.async into a per-warp shared-memory ring of 32 lines (two chunks of 16 groups = 64 bands each; the next
 * chunk is in flight while the current one is walked), which takes HBM latency off the dependent chain. All lanes
 * walk in lock-step (the state is warp-uniform); step n's pair is parked in lane n%32 and every 32 steps the warp
 * (i) stores 32 pairs with one coalesced 256-B store, from the END of the read's capacity region backwards so the
 * list comes out ascending with no reversal pass, (ii) evaluates the 32 emissions in parallel and (iii) adds them
 * to the QC sum strictly in traceback order (the reference's summation order, src/align.c:476).                    */

#define ABEA_TB_CHUNK_GROUPS 8
#define ABEA_TB_RING_GROUPS 32

/* Shared-memory reads by 32-bit shared address: the traceback ring is reached through a pointer, and a generic
 * pointer would make the compiler rebuild the shared window address (S2R + LEA) inside the dependent loop. */
#ifdef ABEA_SIMT_EMU
typedef uint32_t* abea_sptr_t;
__device__ __forceinline__ abea_sptr_t abea_smem_base(uint32_t* p) { return p; }
__device__ __forceinline__ uint32_t abea_lds_u32(abea_sptr_t base, int32_t word) { return base[word]; }
#else
typedef uint32_t abea_sptr_t;
__device__ __forceinline__ abea_sptr_t abea_smem_base(uint32_t* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t abea_lds_u32(abea_sptr_t base, int32_t word) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + 4u * (uint32_t)word) : "memory");
    return v;
}
#endif

/* start the asynchronous copy of trace chunk `chunk` (groups 8*chunk .. 8*chunk+7 = 32 bands) into the ring */
__device__ __forceinline__ void abea_tb_prefetch(uint32_t* ring, const uint32_t* __restrict__ tr, int32_t chunk, int lane) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int idx = i * 32 + lane;                       /* 64 pieces of 16 B = 8 lines of 128 B */
        int32_t g = chunk * ABEA_TB_CHUNK_GROUPS + (idx >> 3);
        abea_cp_async16(ring + (g & (ABEA_TB_RING_GROUPS - 1)) * ABEA_TRACE_GROUP_WORDS + (idx & 7) * 4,
                        tr + (int64_t)g * ABEA_TRACE_GROUP_WORDS + (idx & 7) * 4);
    }
}

/* emissions of the pairs parked in the lanes (first `cnt` lanes valid), added to `sum` in lane order */
__device__ __forceinline__ double abea_tb_flush(double sum, int cnt, int lane, int32_t pk, int32_t pe, int32_t n_before,
                                                const float* __restrict__ ev, const float4* __restrict__ kpr,
                                                abea_pair_t* __restrict__ out, int32_t pair_cap) {
    double lpd = 0.0;
    if (lane < cnt) {
        abea_pair_t p;
        p.ref_pos = pk;
        p.read_pos = pe;
        out[pair_cap - 1 - (n_before + lane)] = p;
        float4 kp = kpr[pk];
        lpd = (double)abea_emission(ev[pe], kp.x, kp.y, kp.z);
    }
    for (int j = 0; j < cnt; j++) sum = __dadd_rn(sum, __shfl_sync(ABEA_FULL, lpd, j));
    return sum;
}

/* The path codes of a finished list (out[0 .. total), ascending, written by lanes of this warp; the caller has
 * synchronised the warp): 32 steps per trip — two coalesced loads, two ballots — and one coalesced store of 32 words
 * every 32 trips. */
__device__ __forceinline__ void abea_put_code(abea_code_t* a, abea_code_t* b, int64_t i, const abea_code_t& v) {
    if (a) a[i] = v;
    if (b) b[i] = v;
}
/* dst_a / dst_b: the read's region in the mapped host buffer and / or in the device buffer (either may be NULL) */
__device__ __forceinline__ void abea_emit_codes(abea_code_t* dst_a, abea_code_t* dst_b, const abea_pair_t* __restrict__ out,
                                                int32_t total, int lane) {
    if (total <= 0 || (!dst_a && !dst_b)) return;
    if (lane == 0) {
        abea_code_t h;
        h.a = (uint32_t)out[0].ref_pos;
        h.b = (uint32_t)out[0].read_pos;
        abea_put_code(dst_a, dst_b, 0, h);
    }
    const int32_t steps = total - 1;
    abea_code_t keep;
    keep.a = keep.b = 0u;
    int32_t j = 0;
    for (int32_t t0 = 0; t0 < steps; t0 += 32, j++) {
        const int32_t t = t0 + lane;
        int dk = 0, de = 0;
        if (t < steps) {
            const abea_pair_t p = out[t], q = out[t + 1];
            dk = q.ref_pos != p.ref_pos;
            de = q.read_pos != p.read_pos;
        }
        const uint32_t mk = __ballot_sync(ABEA_FULL, dk), me = __ballot_sync(ABEA_FULL, de);
        if ((j & 31) == lane) {
            keep.a = mk;
            keep.b = me;
        }
        if ((j & 31) == 31) abea_put_code(dst_a, dst_b, 1 + (j & ~31) + lane, keep);
    }
    if ((j & 31) != 0 && lane < (j & 31)) abea_put_code(dst_a, dst_b, 1 + (j & ~31) + lane, keep);
}
__device__ __forceinline__ void abea_emit_codes_io(const abea_stream_t& io, const abea_read_t& rd, const abea_pair_t* out,
                                                   int32_t total, int lane) {
    const int64_t off = abea_code_offset(rd.pair_off, rd.orig_index);
    abea_emit_codes(io.codes_final ? io.codes_final + off : nullptr, nullptr, out, total, lane);
}

/* Traceback + QC of one read by one warp. `ring` is the warp's 4 KB shared-memory ring (32 trace lines). The trace
 * lines were written by lanes of this warp or CTA; the caller has synchronised (__syncwarp / __syncthreads). */
__device__ __forceinline__ void abea_traceback_read(const abea_read_t& rd, int32_t ridx, int32_t end_event, uint32_t* ring,
                                                    int lane, const float* __restrict__ means,
                                                    const float4* __restrict__ kparams, const uint32_t* __restrict__ trace,
                                                    abea_pair_t* __restrict__ pairs, abea_result_t* __restrict__ results,
                                                    const abea_stream_t& io) {
    const int32_t K = rd.n_kmers;
    const float* __restrict__ ev = means + rd.ev_off;
    const float4* __restrict__ kpr = kparams + rd.kp_off;
    const uint32_t* __restrict__ tr = trace + rd.trace_off;
    abea_pair_t* out = pairs + rd.pair_off;

    int32_t ce = end_event;
    int32_t ck = K - 1;
    int32_t n = 0, gap = 0, max_gap = 0;
    int32_t last_k = ck;
    double sum = 0.0;
    int32_t pk = 0, pe = 0; /* the pair parked in this lane */

    __syncwarp();           /* the previous read's ring is no longer being read by any lane */
    /* The ring holds 4 chunks of 32 bands. Before a step at band b the chunks of b, b-1 and b-2 must have
     * landed (the step reads the trace word of b and the lower-left event index of whichever of b-1 / b-2 comes
     * next); c_ready is the lowest landed chunk, c_ready-1 is in flight. */
    int32_t b = ce + ck + 2;
    int32_t c_ready = b >> 5;
    abea_tb_prefetch(ring, tr, c_ready, lane);
    if (c_ready > 0) abea_tb_prefetch(ring, tr, c_ready - 1, lane);
    abea_cp_async_wait_all();
    __syncwarp();
    if (c_ready > 0) c_ready -= 1;
    if (c_ready > 0) abea_tb_prefetch(ring, tr, c_ready - 1, lane);
    const abea_sptr_t rs = abea_smem_base(ring);
    /* word index of band b's line in the ring: ((b>>2) & 31) * 32 == (b & 124) << 3 */
#define ABEA_TB_LINE(bb) (((bb) & 124) << 3)
    int32_t line = ABEA_TB_LINE(b);
    int32_t q8 = (b & 3) << 3;
    int32_t eb_cur = (int32_t)abea_lds_u32(rs, line + ABEA_LANES + (b & 3));

    while ((ck | ce) >= 0) {
        /* emit (reference src/align.c:458-460): park the pair in lane n%32 */
        if (lane == (n & 31)) {
            pk = ck;
            pe = ce;
        }
        n++;
        last_k = ck;
        if ((n & 31) == 0) sum = abea_tb_flush(sum, 32, lane, pk, pe, n - 32, ev, kpr, out, rd.pair_cap);

        /* inside the loop b = ce + ck + 2 >= 2 */
        const int32_t b1 = b - 1, b2 = b - 2;
        if ((b2 >> 5) < c_ready) { /* the walk is about to need the chunk that was in flight */
            abea_cp_async_wait_all();
            __syncwarp();
            c_ready -= 1;
            if (c_ready > 0) abea_tb_prefetch(ring, tr, c_ready - 1, lane);
        }
        /* everything the next step needs about bands b-1 and b-2, off the dependent chain */
        const int32_t line1 = ABEA_TB_LINE(b1), line2 = ABEA_TB_LINE(b2);
        const int32_t eb1 = (int32_t)abea_lds_u32(rs, line1 + ABEA_LANES + (b1 & 3));
        const int32_t eb2 = (int32_t)abea_lds_u32(rs, line2 + ABEA_LANES + (b2 & 3));
        /* the dependent chain: offset -> trace word -> 2 bits */
        const int32_t o = eb_cur - ce;
        const uint32_t tw = abea_lds_u32(rs, line + ((o >> 2) & 31));
        /* an out-of-band start cell is undefined behaviour in the reference; stay in bounds */
        const uint32_t from = ((uint32_t)o < (uint32_t)ABEA_W) ? ((tw >> (q8 + 2 * (o & 3))) & 3u) : ABEA_FROM_D;
        const bool isD = (from == ABEA_FROM_D), isU = (from == ABEA_FROM_U);
        ce -= (from != ABEA_FROM_L) ? 1 : 0;
        ck -= isU ? 0 : 1;
        b = isD ? b2 : b1;
        line = isD ? line2 : line1;
        q8 = (b & 3) << 3;
        eb_cur = isD ? eb2 : eb1;
        gap = (isD || isU) ? 0 : gap + 1;
        max_gap = gap > max_gap ? gap : max_gap;
    }
#undef ABEA_TB_LINE
    abea_cp_async_wait_all(); /* drain the chunk still in flight before the ring is reused */
    if ((n & 31) != 0) sum = abea_tb_flush(sum, n & 31, lane, pk, pe, n & ~31, ev, kpr, out, rd.pair_cap);

    /* QC (reference src/align.c:526-543) */
    double avg = sum / (double)n;
    bool spanned = (n > 0) && (last_k == 0);
    bool fail = (avg < -5.0) || !spanned || (max_gap > 50);
    if (lane == 0) {
        results[ridx].sum_emission = sum;
        results[ridx].n_aligned = n;
        results[ridx].n_pairs = fail ? 0 : n;
        results[ridx].max_gap = max_gap;
        io.n_pairs_dev[rd.orig_index] = fail ? 0 : n; /* db->n_event_align_pairs[i] */
    }
    /* move the list to the front of the read's capacity region (the layout the caller's buffer has): in place in
     * d_pairs (destination index t <= source index cap-n+t, batches of 32 are loaded before they are stored), and,
     * when the caller's buffer is mapped, also straight into it (posted writes over PCIe) — the device copy stays
     * for consumers on the GPU (abea_device_results, the NCCL gather of a multi-GPU driver). */
    __syncwarp();
    const int32_t src0 = rd.pair_cap - n;
    abea_pair_t* fin = io.pairs_final ? io.pairs_final + rd.pair_off : nullptr;
    if (!fail && (src0 > 0 || fin != nullptr)) {
        for (int32_t t0 = 0; t0 < n; t0 += 32) {
            const int32_t t = t0 + lane;
            abea_pair_t p;
            p.ref_pos = 0;
            p.read_pos = 0;
            if (t < n) p = out[src0 + t];
            __syncwarp();
            if (t < n) {
                if (src0 > 0) out[t] = p;
                if (fin) fin[t] = p;
            }
            __syncwarp();
        }
    }
    if (!fail) {
        __syncwarp();
        abea_emit_codes_io(io, rd, out, n, lane);
    }
    /* the count in the caller's buffer doubles as the read's "done" flag: it is written after the list, behind a
     * system-scope fence, so a host thread that sees a count >= 0 may copy the list out while other reads are still
     * being aligned (abea_align_ragged) */
    if (io.n_pairs_final) {
#ifndef ABEA_SIMT_EMU
        __threadfence_system();
#endif
        __syncwarp();
        if (lane == 0) *(volatile int32_t*)(io.n_pairs_final + rd.orig_index) = fail ? 0 : n;
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Segment-parallel traceback. The serial walk above is a chain of P dependent steps in which 31 lanes repeat what
 * lane 0 does; for the longest read of a batch it is a third of the read's latency, and over a batch a fifth of all
 * warp cycles. Here every LANE walks its own stretch of the path:
 *
 *   the band range [0, b_top] is cut into 32 segments; the lane of segment l does not know in which cell the path enters
 *   it, so it starts `margin` bands ABOVE its segment in the middle of the band (offset 50 — the adaptive band keeps
 *   the path near it) and walks down. Two backward walks that ever stand on the same cell stay together from there on,
 *   and lattice paths built from the steps (-1,-1), (-1,0), (0,-1) cannot cross without sharing a cell, so a
 *   speculative walk falls onto the true path after a few tens of bands (a detour from the best path costs an emission
 *   mismatch per event or a skip at log 1e-10). Pass 1 records, per lane, the cell in which its walk enters the segment
 *   (X), the cell in which it leaves it (Y) and the steps in between. Then the chain is verified: the true entry of the
 *   top segment is the end cell; the true entry of segment l-1 is Y of segment l PROVIDED the walk of segment l has
 *   met the true path before leaving the segment — which its lane checks by walking from the true entry side by side
 *   with its own walk until the two stand on one cell (a few tens of steps; counted in abea_result_t.respec); only a
 *   walk that leaves its segment still apart changes Y, and then the lanes below check again.
 *   With the true entries and the step counts known, every pair's position in the output is
 *   known, and pass 2 has each lane walk its segment once more from its true entry, writing the pairs where they
 *   belong (ascending, at the front of the read's region: no reversal, no slide) and evaluating the emissions.
 *
 * The QC sum (reference src/align.c:476) is a double sum of float emissions in traceback order. Every term is a
 * multiple of q = 2^(e_min - 23), e_min the smallest exponent among the terms; if sum |term| < 2^52 q, every partial
 * sum of ANY grouping is exactly representable, no addition ever rounds, and the per-lane partial sums add up to the
 * reference's value bit for bit. With the built-in models |emission| >= 0.97 (level_stdv >= 1.05), so this holds for
 * every read; when it does not (a tiny emission next to a long read), the sum is redone in traceback order from the
 * finished pair list, terms in parallel, additions in order.
 * The result is the serial walk's, bit for bit: same pairs, same sum, same max_gap, same QC verdict.                  */

struct abea_tbw_t { /* one lane's walker */
    int32_t ce, ck;  /* current cell (event, k-mer); a negative coordinate = the walk has ended (src/align.c:454) */
    int32_t eb;      /* event index of the lower-left cell of band ce + ck + 2 */
};

__device__ __forceinline__ int32_t abea_tbw_eb(const uint32_t* __restrict__ tr, int32_t b) {
    return (int32_t)tr[(int64_t)(b >> 2) * ABEA_TRACE_GROUP_WORDS + ABEA_LANES + (b & 3)];
}

/* one step back from a valid cell (ce, ck >= 0, so its band is >= 2); returns the cell's trace code */
__device__ __forceinline__ uint32_t abea_tbw_step(abea_tbw_t& w, const uint32_t* __restrict__ tr) {
    const int32_t b = w.ce + w.ck + 2;
    const uint32_t* line = tr + (int64_t)(b >> 2) * ABEA_TRACE_GROUP_WORDS;
    const int32_t o = w.eb - w.ce;
    /* the two bands the walk can go to next, off the dependent chain */
    const int32_t eb1 = abea_tbw_eb(tr, b - 1), eb2 = abea_tbw_eb(tr, b - 2);
    const uint32_t tw = line[(o >> 2) & 31];
    /* an out-of-band cell is undefined behaviour in the reference; stay in bounds */
    const uint32_t from = ((uint32_t)o < (uint32_t)ABEA_W) ? ((tw >> (((b & 3) << 3) + 2 * (o & 3))) & 3u) : ABEA_FROM_D;
    w.ce -= (from != ABEA_FROM_L) ? 1 : 0;
    w.ck -= (from != ABEA_FROM_U) ? 1 : 0;
    w.eb = (from == ABEA_FROM_D) ? eb2 : eb1;
    return from;
}

__device__ __forceinline__ void abea_prefetch_l1(const void* p) {
#ifndef ABEA_SIMT_EMU
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

/* K walkers per lane, stepped together: a step is one round trip to the trace (L2 latency under load, the lanes of a
 * warp read 32 different lines), so what a lane's time is made of is round trips, and K independent walks in flight
 * cost one. The loads of all K walkers are issued before any is used; a walker that is not `act` re-reads band 2 of
 * the read (always there) and keeps its state. K = 4 for the reads of the wide kernel (alone on their SM, bound by the
 * round trips: 95 -> 46-70 cycles per traceback step); K = 1 for the narrow warps, whose steps compete with eleven
 * other warps for issue slots and load wavefronts (with K = 4 the extra walks cost more than the round trips they
 * save: measured 10.3 -> 11.2 ms on the target config, profiles/read_cycles_partc_k4_r02_*). */
#define ABEA_TB_K4_MIN_BANDS 16384
template <int K>
__device__ __forceinline__ void abea_tbw_step_all(abea_tbw_t (&w)[K], const bool (&act)[K],
                                                  const uint32_t* __restrict__ tr, uint32_t (&from)[K]) {
    uint32_t tw[K];
    int32_t e1[K], e2[K], bb[K], oo[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int32_t b = act[k] ? (w[k].ce + w[k].ck + 2) : 2;
        const int32_t o = act[k] ? (w[k].eb - w[k].ce) : 0;
        bb[k] = b;
        oo[k] = o;
        tw[k] = tr[(int64_t)(b >> 2) * ABEA_TRACE_GROUP_WORDS + ((o >> 2) & 31)];
        e1[k] = abea_tbw_eb(tr, b - 1);
        e2[k] = abea_tbw_eb(tr, b - 2);
        /* the line the walk reaches four bands further down, at the word it is most likely to need there (the offset
         * drifts by at most one per band) and at the band event indices: in L1 by the time the walk gets to it */
        abea_prefetch_l1(tr + (int64_t)((b >= 6 ? b - 4 : 2) >> 2) * ABEA_TRACE_GROUP_WORDS + ((o >> 2) & 31));
        abea_prefetch_l1(tr + (int64_t)((b >= 6 ? b - 4 : 2) >> 2) * ABEA_TRACE_GROUP_WORDS + ABEA_LANES);
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        const uint32_t f = ((uint32_t)oo[k] < (uint32_t)ABEA_W) ? ((tw[k] >> (((bb[k] & 3) << 3) + 2 * (oo[k] & 3))) & 3u) : ABEA_FROM_D;
        from[k] = f;
        if (act[k]) {
            w[k].ce -= (f != ABEA_FROM_L) ? 1 : 0;
            w[k].ck -= (f != ABEA_FROM_U) ? 1 : 0;
            w[k].eb = (f == ABEA_FROM_D) ? e2[k] : e1[k];
        }
    }
}
__device__ __forceinline__ bool abea_tbw_above(const abea_tbw_t& w, int32_t floor_band) {
    return (w.ce | w.ck) >= 0 && w.ce + w.ck + 2 > floor_band;
}

__device__ __forceinline__ int32_t __reduce_add_sync_compat(int32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(ABEA_FULL, v, d);
    return v;
}

template <bool FAST, int K>
__device__ __forceinline__ void abea_traceback_par(const abea_read_t& rd, int32_t ridx, int32_t end_event, int lane,
                                                   const float* __restrict__ means, const float4* __restrict__ kparams,
                                                   const uint32_t* __restrict__ trace, abea_pair_t* __restrict__ pairs,
                                                   abea_result_t* __restrict__ results, const abea_stream_t& io) {
    const int32_t E = rd.n_events, NK = rd.n_kmers;
    const float* __restrict__ ev = means + rd.ev_off;
    const float4* __restrict__ kpr = kparams + rd.kp_off;
    const uint32_t* __restrict__ tr = trace + rd.trace_off;
    abea_pair_t* out = pairs + rd.pair_off;
    __syncwarp(); /* the trace lines written by the other lanes of this warp are visible */

    /* 32 K segments of S bands: segment s = lane * K + k covers bands (s S - 1, (s + 1) S - 1]; a higher segment is
     * earlier in traceback order */
    const int32_t b_top = end_event + (NK - 1) + 2;
    int32_t S = (b_top + 32 * K) / (32 * K); /* ceil((b_top + 1) / (32 K)) */
    if (S < 16) S = 16;
    const int32_t s_top = b_top / S;
    const int32_t margin = io.tb_margin;
    const int32_t eb_top = abea_tbw_eb(tr, b_top);

    /* ---- pass 1: speculative entry, steps, exit of every segment ---- */
    abea_tbw_t w[K];
    int32_t floor_b[K], n[K], ve[K], vk[K], ye[K], yk[K];
    bool act[K];
    uint32_t from[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int32_t sg = lane * K + k;
        floor_b[k] = sg * S - 1;
        n[k] = 0;
        w[k].ce = -1; w[k].ck = -1; w[k].eb = 0;
        if (sg <= s_top) {
            const int32_t st = floor_b[k] + S + margin;
            if (sg == s_top || st >= b_top) { /* close enough to the end cell: walk from it, nothing to speculate on */
                w[k].ce = end_event;
                w[k].ck = NK - 1;
                w[k].eb = eb_top;
            } else {
                const int32_t eb = abea_tbw_eb(tr, st);
                const int32_t ce = eb - ABEA_W / 2, ck = st - 2 - ce;
                if (ce >= 0 && ce < E && ck >= 0 && ck < NK) {
                    w[k].ce = ce;
                    w[k].ck = ck;
                    w[k].eb = eb;
                }
            }
        }
    }
    for (;;) { /* down to the segment's top band */
        bool any = false;
#pragma unroll
        for (int k = 0; k < K; k++) {
            act[k] = abea_tbw_above(w[k], floor_b[k] + S);
            any = any || act[k];
        }
        if (!any) break;
        abea_tbw_step_all(w, act, tr, from);
    }
#pragma unroll
    for (int k = 0; k < K; k++) { /* (ve, vk): the entry cell the segment's step count and exit cell are valid for */
        ve[k] = w[k].ce;
        vk[k] = w[k].ck;
    }
    for (;;) { /* through the segment */
        bool any = false;
#pragma unroll
        for (int k = 0; k < K; k++) {
            act[k] = abea_tbw_above(w[k], floor_b[k]);
            any = any || act[k];
            n[k] += act[k] ? 1 : 0;
        }
        if (!any) break;
        abea_tbw_step_all(w, act, tr, from);
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        ye[k] = w[k].ce;
        yk[k] = w[k].ck;
        if (lane * K + k > s_top) n[k] = 0;
    }
    __syncwarp();

    /* ---- verify: the true entry of segment s is the exit of segment s + 1 (the top one starts at the end cell) ----
     * A segment whose guess was not the true entry walks both paths side by side — always stepping the one that is
     * higher up — until they stand on the same cell: from there on they are one path, so the exit stays and only the
     * step count is corrected. That costs a few tens of steps. If they leave the segment apart, the exit changes and
     * the segments below look again; rounds repeat until nothing changes (one round, as a rule). */
    int32_t respec = 0;
    for (;;) {
        const int32_t ne0 = __shfl_down_sync(ABEA_FULL, ye[0], 1), nk0 = __shfl_down_sync(ABEA_FULL, yk[0], 1);
        abea_tbw_t A[K], B[K];
        int32_t sa[K], sb[K], ue[K], uk[K];
        bool need[K], fin[K];
        bool any_need = false;
#pragma unroll
        for (int k = 0; k < K; k++) {
            ue[k] = (k + 1 < K) ? ye[(k + 1 < K) ? k + 1 : k] : ne0;
            uk[k] = (k + 1 < K) ? yk[(k + 1 < K) ? k + 1 : k] : nk0;
            need[k] = (lane * K + k < s_top) && (ue[k] != ve[k] || uk[k] != vk[k]);
            any_need = any_need || need[k];
            fin[k] = !need[k];
            sa[k] = sb[k] = 0;
            A[k].ce = ue[k]; A[k].ck = uk[k]; A[k].eb = 0;
            B[k].ce = ve[k]; B[k].ck = vk[k]; B[k].eb = 0;
            if (need[k]) {
                if ((ue[k] | uk[k]) >= 0) A[k].eb = abea_tbw_eb(tr, ue[k] + uk[k] + 2);
                if ((ve[k] | vk[k]) >= 0) B[k].eb = abea_tbw_eb(tr, ve[k] + vk[k] + 2);
            }
        }
        bool changed = false;
        if (any_need) {
            for (;;) {
                abea_tbw_t W[K];
                bool stepA[K];
                bool any = false;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    act[k] = false;
                    stepA[k] = false;
                    W[k] = A[k];
                    if (!fin[k]) {
                        if (A[k].ce == B[k].ce && A[k].ck == B[k].ck) {
                            fin[k] = true;
                        } else {
                            const bool la = abea_tbw_above(A[k], floor_b[k]), lb = abea_tbw_above(B[k], floor_b[k]);
                            if (!la && !lb) {
                                fin[k] = true;
                            } else {
                                stepA[k] = la && (!lb || A[k].ce + A[k].ck >= B[k].ce + B[k].ck);
                                W[k] = stepA[k] ? A[k] : B[k];
                                act[k] = true;
                                any = true;
                            }
                        }
                    }
                }
                if (!any) break;
                abea_tbw_step_all(W, act, tr, from);
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if (act[k]) {
                        if (stepA[k]) { A[k] = W[k]; sa[k]++; } else { B[k] = W[k]; sb[k]++; }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (need[k]) {
                    if (A[k].ce == B[k].ce && A[k].ck == B[k].ck) {
                        n[k] = n[k] - sb[k] + sa[k]; /* the rest of B's walk is A's */
                    } else {
                        n[k] = sa[k];
                        ye[k] = A[k].ce;
                        yk[k] = A[k].ck;
                        changed = true;
                    }
                    ve[k] = ue[k];
                    vk[k] = uk[k];
                    respec++;
                }
            }
        }
        if (!__any_sync(ABEA_FULL, changed)) break;
    }
    respec = __reduce_add_sync_compat(respec);

    /* ---- positions: steps of all segments above (earlier in traceback order) ---- */
    int32_t lane_n = 0;
#pragma unroll
    for (int k = 0; k < K; k++) lane_n += n[k];
    int32_t incl = lane_n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int32_t v = __shfl_down_sync(ABEA_FULL, incl, d);
        if (lane + d < 32) incl += v;
    }
    const int32_t total = __shfl_sync(ABEA_FULL, incl, 0);
    int32_t pos[K];
    {
        int32_t before = incl - lane_n; /* the lanes above */
#pragma unroll
        for (int k = K - 1; k >= 0; k--) {
            pos[k] = total - 1 - before;
            before += n[k];
        }
    }

    /* ---- pass 2: pairs, emissions, gap runs ---- */
    double sum = 0.0, asum = 0.0;
    uint32_t emin = 0xffu;                             /* smallest biased exponent among the non-zero emissions */
    int32_t lead_run[K], max_run[K], run[K], last_kk[K]; /* runs of FROM_L: at the head of a segment, anywhere, at its tail */
    bool all_l[K];
    int32_t nmax = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        w[k].ce = ve[k];
        w[k].ck = vk[k];
        w[k].eb = (n[k] > 0) ? abea_tbw_eb(tr, ve[k] + vk[k] + 2) : 0;
        lead_run[k] = max_run[k] = run[k] = 0;
        last_kk[k] = 0;
        all_l[k] = true;
        nmax = n[k] > nmax ? n[k] : nmax;
    }
    for (int32_t j = 0; j < nmax; j++) {
        float x[K];
        float4 kp[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            act[k] = j < n[k];
            const int32_t e = act[k] ? w[k].ce : 0, q = act[k] ? w[k].ck : 0;
            x[k] = ev[e];
            kp[k] = kpr[q];
            abea_prefetch_l1(ev + (e >= 12 ? e - 12 : 0)); /* the walk moves down both arrays a step at a time */
            abea_prefetch_l1(kpr + (q >= 6 ? q - 6 : 0));
            if (act[k]) {
                abea_pair_t p;
                p.ref_pos = w[k].ck;
                p.read_pos = w[k].ce;
                out[pos[k] - j] = p;
                last_kk[k] = w[k].ck;
            }
        }
        abea_tbw_step_all(w, act, tr, from);
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (act[k]) {
                const float lp = abea_emission_t<FAST>(x[k], kp[k]);
                const uint32_t ex = (__float_as_uint(lp) >> 23) & 0xffu;
                if (lp != 0.0f && ex < emin) emin = ex;
                const double lpd = (double)lp;
                sum = __dadd_rn(sum, lpd);
                asum = __dadd_rn(asum, fabs(lpd));
                if (from[k] == ABEA_FROM_L) {
                    run[k]++;
                    if (run[k] > max_run[k]) max_run[k] = run[k];
                } else {
                    if (all_l[k]) lead_run[k] = run[k];
                    all_l[k] = false;
                    run[k] = 0;
                }
            }
        }
    }
    /* a lane's segments in traceback order (k descending) as one stretch: steps, leading / longest / trailing run */
    int32_t L_n = 0, L_lead = 0, L_max = 0, L_tail = 0, L_lastk = 0;
    bool L_all = true, L_has = false;
#pragma unroll
    for (int k = K - 1; k >= 0; k--) {
        if (n[k] > 0) {
            if (all_l[k]) {
                L_tail += n[k];
                if (L_all) L_lead = L_tail;
                if (L_tail > L_max) L_max = L_tail;
            } else {
                const int32_t joined = L_tail + lead_run[k];
                if (L_all) L_lead = joined;
                if (joined > L_max) L_max = joined;
                if (max_run[k] > L_max) L_max = max_run[k];
                L_tail = run[k];
                L_all = false;
            }
            L_n += n[k];
            L_lastk = last_kk[k];
            L_has = true;
        }
    }
    (void)L_has;
    __syncwarp();

    /* ---- combine across lanes, traceback order = lanes descending ---- */
    int32_t max_gap = 0, carry = 0;
    for (int32_t l = 31; l >= 0; l--) {
        const int32_t ln = __shfl_sync(ABEA_FULL, L_n, l);
        const int32_t llead = __shfl_sync(ABEA_FULL, L_lead, l), lmax = __shfl_sync(ABEA_FULL, L_max, l);
        const int32_t ltail = __shfl_sync(ABEA_FULL, L_tail, l);
        const bool lall = __shfl_sync(ABEA_FULL, L_all ? 1 : 0, l) != 0;
        if (ln == 0) continue;
        if (lall) {
            carry += ln;
            if (carry > max_gap) max_gap = carry;
        } else {
            if (carry + llead > max_gap) max_gap = carry + llead;
            if (lmax > max_gap) max_gap = lmax;
            carry = ltail;
        }
    }
    /* the last pair emitted (the lowest lane that has steps) */
    const uint32_t have = __ballot_sync(ABEA_FULL, L_n > 0);
    const int low = have ? (__ffs((int)have) - 1) : 0;
    const int32_t last_k = __shfl_sync(ABEA_FULL, L_lastk, low);

    /* ---- the QC sum ---- */
    double a_tot = asum;
    uint32_t e_all = emin;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a_tot = __dadd_rn(a_tot, __shfl_xor_sync(ABEA_FULL, a_tot, d));
        const uint32_t oe = __shfl_xor_sync(ABEA_FULL, e_all, d);
        e_all = oe < e_all ? oe : e_all;
    }
    /* no addition rounds if sum|x| < 2^52 * 2^(e_min - 23) (one bit spare for the rounding of a_tot itself); e_all is
     * biased by 127 and 0 means a denormal term: then the ordered sum below decides */
    bool exact = false;
    if (e_all == 0xffu) {
        exact = true; /* every term is zero */
    } else if (e_all > 0u) {
        const int32_t lim_exp = (int32_t)e_all - 127 - 23 + 52; /* a_tot must stay below 2^lim_exp */
        exact = a_tot < __hiloint2double((lim_exp + 1023) << 20, 0);
    }
    if (io.tb_mode & 2) exact = false; /* tests: always take the ordered sum */
    if (exact) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum = __dadd_rn(sum, __shfl_xor_sync(ABEA_FULL, sum, d));
    } else { /* traceback order = descending position in the finished list; terms in parallel, additions in order */
        __syncwarp();
        sum = 0.0;
        for (int32_t t0 = 0; t0 < total; t0 += 32) {
            const int32_t cnt = (total - t0 < 32) ? total - t0 : 32;
            double lpd = 0.0;
            if (lane < cnt) {
                const abea_pair_t p = out[total - 1 - (t0 + lane)];
                const float4 kp = kpr[p.ref_pos];
                lpd = (double)abea_emission(ev[p.read_pos], kp.x, kp.y, kp.z);
            }
            for (int j = 0; j < cnt; j++) sum = __dadd_rn(sum, __shfl_sync(ABEA_FULL, lpd, j));
        }
    }

    /* QC (reference src/align.c:526-543) */
    const double avg = sum / (double)total;
    const bool spanned = (total > 0) && (last_k == 0);
    const bool fail = (avg < -5.0) || !spanned || (max_gap > 50);
    const int32_t rs = respec;
    if (lane == 0) {
        results[ridx].sum_emission = sum;
        results[ridx].n_aligned = total;
        results[ridx].n_pairs = fail ? 0 : total;
        results[ridx].max_gap = max_gap;
        results[ridx].respec = rs;
        io.n_pairs_dev[rd.orig_index] = fail ? 0 : total; /* db->n_event_align_pairs[i] */
    }
    /* the list already sits at the front of the read's region in d_pairs; a mapped caller buffer gets a coalesced copy */
    __syncwarp();
    if (io.pairs_final) {
        abea_pair_t* fin = io.pairs_final + rd.pair_off;
        if (!fail)
            for (int32_t t = lane; t < total; t += 32) fin[t] = out[t];
    }
    if (!fail) abea_emit_codes_io(io, rd, out, total, lane);
    if (io.n_pairs_final) {
#ifndef ABEA_SIMT_EMU
        __threadfence_system();
#endif
        __syncwarp();
        if (lane == 0) *(volatile int32_t*)(io.n_pairs_final + rd.orig_index) = fail ? 0 : total;
    }
}

/* A streamed read filled by the FAST instantiation may have lost its FAST bit while it was being filled: the loader
 * clears the bit when an out-of-range event mean passes through it, BEFORE it publishes that piece, and the fill has
 * consumed every piece by now — so the bit is final here. Such a read is filled again, last, by the EXACT instantiation,
 * which overwrites every result; but a host thread takes a read's list the moment its count appears in the mapped
 * count array, so the first, invalid result must never be published there. */
template <bool FAST, bool STREAM>
__device__ __forceinline__ abea_stream_t abea_publishable(const abea_stream_t& io, const uint32_t* flag) {
    abea_stream_t r = io;
    if (FAST && STREAM) {
        if ((abea_ld_acquire_u32(flag) & ABEA_READ_FAST) == 0u) {
            r.pairs_final = nullptr;
            r.n_pairs_final = nullptr;
            r.codes_final = nullptr;
        }
    }
    return r;
}

/* the traceback of one read by the warp that filled it (warp 0 of a wide CTA) */
template <bool FAST, int K>
__device__ __forceinline__ void abea_traceback(const abea_read_t& rd, int32_t ridx, int32_t end_event, uint32_t* ring, int lane,
                                               const float* __restrict__ means, const float4* __restrict__ kparams,
                                               const uint32_t* __restrict__ trace, abea_pair_t* __restrict__ pairs,
                                               abea_result_t* __restrict__ results, const abea_stream_t& io) {
    if (io.tb_mode == 0) {
        abea_traceback_read(rd, ridx, end_event, ring, lane, means, kparams, trace, pairs, results, io);
        if (lane == 0) results[ridx].respec = 0;
    } else {
        abea_traceback_par<FAST, K>(rd, ridx, end_event, lane, means, kparams, trace, pairs, results, io);
    }
}

/* One band of scores as held by a lane, with the two neighbour-lane cells next to its four ("halos"). */
struct abea_band_t {
    double R[ABEA_CPL];
    double lo; /* lane-1's R[3]  (offset 4*lane-1), -inf for lane 0  */
    double hi; /* lane+1's R[0]  (offset 4*lane+4), -inf for lane 24 */
};

#define ABEA_RING 64 /* entries of the per-warp shared-memory rings (two chunks of 32) */

struct abea_fill_smem_t {
    float ev[ABEA_RING];   /* event means, slot = event index & 63 */
    float4 kp[ABEA_RING];  /* k-mer parameters, slot = k-mer index & 63 */
    /* streaming: looked at once per 32 events, so they live here and not in registers of the band loop */
    const uint32_t* ready_w; /* the read's words of d_ready */
    int32_t ev_landed;       /* leading events of the read known to be in d_events */
    int32_t pad;
};

/* asynchronously stage chunk `chunk` (indices 32*chunk .. +31, clamped into the read) of the event means / k-mer
 * parameters into the warp's ring: the "TMA/async-copy staging of the active band window" of the design */
__device__ __forceinline__ void abea_stage_events(abea_fill_smem_t* sm, const float* __restrict__ ev, int32_t chunk,
                                                  int32_t E, int lane) {
    int32_t e = chunk * 32 + lane;
    int32_t ec = e < 0 ? 0 : (e >= E ? E - 1 : e);
    abea_cp_async4(&sm->ev[e & (ABEA_RING - 1)], &ev[ec]);
}
__device__ __forceinline__ void abea_stage_kparams(abea_fill_smem_t* sm, const float4* __restrict__ kpr, int32_t chunk,
                                                   int32_t K, int lane) {
    int32_t k = chunk * 32 + lane;
    int32_t kc = k < 0 ? 0 : (k >= K ? K - 1 : k);
    abea_cp_async16(&sm->kp[k & (ABEA_RING - 1)], &kpr[kc]);
}

/* The four (previous move, this move) geometries, each fully specialised so that every neighbour is a fixed
 * register: RIGHT: up = b-1[o+1], left = b-1[o]; DOWN: up = b-1[o], left = b-1[o-1];
 * diag = b-2[o+1] (right,right), b-2[o-1] (down,down), else b-2[o]. A = band b-1, B = band b-2. */
template <bool FAST, bool RIGHT, bool PREV_RIGHT>
__device__ __forceinline__ void abea_band_cells(const float* x, const float4* kp, const abea_band_t& A,
                                                const abea_band_t& B, double lp_step, double lp_stay, double lp_skip,
                                                double* Rn, uint32_t* fr) {
#pragma unroll
    for (int c = 0; c < ABEA_CPL; c++) {
        double up, left, diag;
        if (RIGHT) {
            up = (c < ABEA_CPL - 1) ? A.R[c + 1 < ABEA_CPL ? c + 1 : c] : A.hi;
            left = A.R[c];
        } else {
            up = A.R[c];
            left = (c > 0) ? A.R[c > 0 ? c - 1 : 0] : A.lo;
        }
        if (RIGHT && PREV_RIGHT) diag = (c < ABEA_CPL - 1) ? B.R[c + 1 < ABEA_CPL ? c + 1 : c] : B.hi;
        else if (!RIGHT && !PREV_RIGHT) diag = (c > 0) ? B.R[c > 0 ? c - 1 : 0] : B.lo;
        else diag = B.R[c];
        float lp = abea_emission_t<FAST>(x[c], kp[c]);
        if (FAST) abea_cell_f32((double)lp, up, left, diag, lp_step, lp_stay, lp_skip, Rn[c], fr[c]);
        else abea_cell_d<FAST>(lp, up, left, diag, lp_step, lp_stay, lp_skip, Rn[c], fr[c]);
    }
}

/* Per-read constants and cursors of the fill, warp-uniform. */
struct abea_fill_ctx_t {
    const float* ev;
    const float4* kpr;
    uint32_t* tr;
    double lp_stay, lp_step, lp_skip, lp_trim;
    int32_t E, K;
    int32_t NB;
    int32_t eb, kb;       /* lower-left of the current band */
    int32_t b;            /* band being filled */
    int32_t safe;         /* further bands guaranteed to be interior */
    uint32_t tword;       /* this lane's trace bytes of the current 4-band group */
    int32_t eb_keep;      /* lanes 25..28: event index of the group's bands */
    double best_s;        /* best end cell seen by this lane */
    int32_t best_e;
    float x_next;         /* mean of event eb+1 (enters at offset 0 on the next down move) */
    float4 kp_next;       /* parameters of k-mer kb+100 (enters at offset 99 on the next right move) */
    bool prev_right;
};

/* One band: A holds band b-1, B holds band b-2 and receives band b. `right` is this band's move; returns the next
 * band's move (Suzuki's rule, reference src/align.c:304-322), decided as soon as this band's scores exist so that its
 * shuffle/vote latency overlaps the trace bookkeeping. */
template <bool FAST, bool STREAM>
__device__ __forceinline__ bool abea_fill_step(abea_fill_ctx_t& cx, float* x, float4* kp, const abea_band_t& A,
                                               abea_band_t& B, abea_fill_smem_t* sm, bool right, int lane,
                                               uint32_t* io_stalled) {
    const double NEG = abea_neg_inf_d();
    double Rn[ABEA_CPL];
    uint32_t fr[ABEA_CPL];
    if (right) {
        cx.kb += 1;
        /* k-mer window slides towards lower offsets; k-mer kb+99 enters at offset 99 */
        float4 nb;
        nb.x = __shfl_down_sync(ABEA_FULL, kp[0].x, 1);
        nb.y = __shfl_down_sync(ABEA_FULL, kp[0].y, 1);
        nb.z = __shfl_down_sync(ABEA_FULL, kp[0].z, 1);
        nb.w = __shfl_down_sync(ABEA_FULL, kp[0].w, 1);
#pragma unroll
        for (int c = 0; c < ABEA_CPL - 1; c++) kp[c] = kp[c + 1];
        kp[ABEA_CPL - 1] = (lane == ABEA_LANES - 1) ? cx.kp_next : nb;
        const int32_t kn = cx.kb + ABEA_W; /* next k-mer to enter */
        if ((kn & 31) == 0) {              /* first use of a chunk that was in flight: land it, start the next one */
            abea_cp_async_wait_all();
            __syncwarp();
            abea_stage_kparams(sm, cx.kpr, (kn >> 5) + 1, cx.K, lane);
        }
        cx.kp_next = sm->kp[kn & (ABEA_RING - 1)];
        if (cx.prev_right) abea_band_cells<FAST, true, true>(x, kp, A, B, cx.lp_step, cx.lp_stay, cx.lp_skip, Rn, fr);
        else abea_band_cells<FAST, true, false>(x, kp, A, B, cx.lp_step, cx.lp_stay, cx.lp_skip, Rn, fr);
    } else {
        cx.eb += 1;
        /* event window slides towards higher offsets; event eb enters at offset 0 */
        float nb = __shfl_up_sync(ABEA_FULL, x[ABEA_CPL - 1], 1);
#pragma unroll
        for (int c = ABEA_CPL - 1; c > 0; c--) x[c] = x[c - 1];
        x[0] = (lane == 0) ? cx.x_next : nb;
        const int32_t en = cx.eb + 1;
        if ((en & 31) == 0) {
            abea_cp_async_wait_all();
            __syncwarp();
            /* streaming: the chunk about to be staged (events en+32 .. en+63) may not have crossed PCIe yet */
            if (STREAM) {
                const int32_t last = (en + 63 < cx.E) ? en + 63 : cx.E - 1;
                if (last >= sm->ev_landed) {
                    const int32_t got = abea_wait_landed_events(sm->ready_w, last, io_stalled);
                    __syncwarp();
                    if (lane == 0) sm->ev_landed = got;
                    __syncwarp();
                }
            }
            abea_stage_events(sm, cx.ev, (en >> 5) + 1, cx.E, lane);
        }
        cx.x_next = sm->ev[en & (ABEA_RING - 1)];
        if (cx.prev_right) abea_band_cells<FAST, false, true>(x, kp, A, B, cx.lp_step, cx.lp_stay, cx.lp_skip, Rn, fr);
        else abea_band_cells<FAST, false, false>(x, kp, A, B, cx.lp_step, cx.lp_stay, cx.lp_skip, Rn, fr);
    }
    cx.prev_right = right;

    /* --- band edges: validity window, trim column, end column. Interior bands (all 100 cells valid, trim and end
     * columns outside the band) skip this; `safe` counts how many more bands are certainly interior. --- */
    if (cx.safe > 0) {
        cx.safe -= 1;
    } else {
        const int32_t kb = cx.kb, eb = cx.eb, E = cx.E, K = cx.K;
        const bool interior = (kb >= 0) && (kb + ABEA_W < K) && (eb >= ABEA_W - 1) && (eb <= E - 1);
        if (interior) {
            /* each band advances exactly one of kb, eb by one */
            int32_t s1 = K - ABEA_W - 1 - kb, s2 = E - 1 - eb;
            cx.safe = s1 < s2 ? s1 : s2;
        } else {
            /* offsets whose event and k-mer exist (reference src/align.c:337-346) */
            int32_t lo = -kb;
            if (eb - (E - 1) > lo) lo = eb - (E - 1);
            if (lo < 0) lo = 0;
            int32_t hi = K - kb;
            if (eb + 1 < hi) hi = eb + 1;
            if (hi > ABEA_W) hi = ABEA_W;
            /* trim column: k-mer -1 (reference src/align.c:324-333) */
            const int32_t to = -1 - kb;
            const int32_t te = eb - to;
            const bool trim_in = (to >= 0) && (to < ABEA_W) && (te >= 0) && (te < E);
            const double trim_s = (double)__double2float_rn(__dmul_rn(cx.lp_trim, (double)(te + 1)));
            /* end column: k-mer K-1 (reference src/align.c:429-445) */
            const int32_t oe = (K - 1) - kb;
#pragma unroll
            for (int c = 0; c < ABEA_CPL; c++) {
                int32_t o = ABEA_CPL * lane + c;
                bool valid = (o >= lo) && (o < hi);
                Rn[c] = valid ? Rn[c] : NEG;
                fr[c] = valid ? fr[c] : 0u;
                if (o == to && trim_in) {
                    Rn[c] = trim_s;
                    fr[c] = ABEA_FROM_U;
                }
                if (o == oe && valid) {
                    int32_t e = eb - o;
                    double s = (double)__double2float_rn(__dadd_rn(Rn[c], __dmul_rn((double)(E - e), cx.lp_trim)));
                    if (s > cx.best_s) {
                        cx.best_s = s;
                        cx.best_e = e;
                    }
                }
            }
        }
    }

    /* --- the new band replaces band b-2; its halos and the next move are requested right away --- */
#pragma unroll
    for (int c = 0; c < ABEA_CPL; c++) B.R[c] = Rn[c];
    double hi = __shfl_down_sync(ABEA_FULL, Rn[0], 1);
    double lo = __shfl_up_sync(ABEA_FULL, Rn[ABEA_CPL - 1], 1);
    double ll = __shfl_sync(ABEA_FULL, Rn[0], 0);
    B.hi = (lane >= ABEA_LANES - 1) ? NEG : hi;
    B.lo = (lane == 0) ? NEG : lo;
    const double ur = Rn[ABEA_CPL - 1]; /* meaningful on lane 24 */
    bool my_right = (ll == NEG && ur == NEG) ? (((cx.b + 1) & 1) == 1) : (ll < ur);
    const bool next_right = __any_sync(ABEA_FULL, (lane == ABEA_LANES - 1) && my_right) != 0;

    /* --- trace: 2 bits per cell, one byte per lane per band, one 128-B line per 4 bands --- */
    const uint32_t byte = fr[0] | (fr[1] << 2) | (fr[2] << 4) | (fr[3] << 6);
    const int q = cx.b & 3;
    cx.tword |= byte << (8 * q);
    if (lane == ABEA_LANES + q) cx.eb_keep = cx.eb;
    if (q == 3 || cx.b == cx.NB - 1) {
        cx.tr[(int64_t)(cx.b >> 2) * ABEA_TRACE_GROUP_WORDS + lane] = (lane < ABEA_LANES) ? cx.tword : (uint32_t)cx.eb_keep;
        cx.tword = 0u;
    }
    cx.b += 1;
    return next_right;
}

/* Scheduling of the persistent narrow warps. A CTA has ABEA_NARROW_WARPS_MAX or fewer warps; warp w sits on SM
 * sub-partition w%4. Warps 0..3 are PRIMARY: they pull reads from the head of the longest-first queue. The others
 * are SECONDARY: they pull from the TAIL (shortest first) and, while the primary of their sub-partition is filling a
 * read longer than long_thr bands, they wait between reads — so the reads that set the makespan run alone on their
 * sub-partition (~655 instead of ~780 cycles per band, profiles/README.md) at the cost of idling a few of the 592
 * sub-partitions' second slots. queue[0] counts pulls, queue[1] the head, queue[2] the tail. */
#ifndef ABEA_NARROW_WARPS_MAX
#define ABEA_NARROW_WARPS_MAX 12
#endif
/* Register budget. A resident batch: the CTA's 12 warps may use the whole register file (launch bound 384 threads ->
 * 162 registers, 8 % fewer instructions than at 124: cfg2 11.08 -> 10.81 ms). A streamed batch: the CTAs of
 * abea_load_kernel must fit on the same SMs BESIDE the persistent narrow CTAs (with 162 x 384 registers taken they do
 * not, the narrow grid then waits for the loader to finish and the end-to-end time goes from 13.5 to 20.7 ms — measured),
 * so the STREAM instantiations keep the bound of 512 threads (124 registers). */
#define ABEA_NARROW_BOUND(STREAM) ((STREAM) ? 32 * 16 : 32 * ABEA_NARROW_WARPS_MAX)

/* A paused secondary warp must cost the warps that work next to it as little as possible: every poll is a dozen
 * instructions through the same issue port (at a fixed 2 us request the polls were 6.5 % of all instructions issued,
 * profiles/fill_narrow_opcode_mix_r01.txt — the hardware wakes a sleeper early), so the interval doubles up to
 * ~32 us; a read that was waited for that long is a long one, and 32 us late is nothing against it. */
__device__ __forceinline__ void abea_backoff(unsigned& ns) {
#ifndef ABEA_SIMT_EMU
    __nanosleep(ns);
    if (ns < 32768u) ns <<= 1;
#else
    (void)ns;
#endif
}

template <bool FAST, bool STREAM>
__global__ void __launch_bounds__(ABEA_NARROW_BOUND(STREAM))
abea_fill_kernel(const abea_read_t* __restrict__ reads, int32_t n_reads, const float* __restrict__ means,
                 const float4* __restrict__ kparams, const uint32_t* __restrict__ read_flags,
                 uint32_t* __restrict__ trace, abea_pair_t* __restrict__ pairs, abea_result_t* __restrict__ results,
                 abea_stream_t io, abea_consts_t cst, int32_t* __restrict__ queue,
                 int32_t first, int32_t long_thr, int32_t policy) {
    /* dynamic shared memory: per warp one abea_fill_smem_t and one 4-KB traceback ring */
#ifdef ABEA_SIMT_EMU
    unsigned char* dyn = (unsigned char*)simt::g_dynsmem;
#else
    extern __shared__ __align__(16) unsigned char abea_dyn_smem[];
    unsigned char* dyn = abea_dyn_smem;
#endif
    __shared__ int long_flag[4]; /* only ever touched with atomics: a flag, not data */
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    abea_fill_smem_t* sm = (abea_fill_smem_t*)dyn + wid;
    uint32_t* tb_ring = (uint32_t*)(dyn + (size_t)nwarps * sizeof(abea_fill_smem_t)) +
                        (size_t)wid * (ABEA_TB_RING_GROUPS * ABEA_TRACE_GROUP_WORDS);
    const bool primary = wid < 4;
    const int slot = wid & 3;
    /* policy 1 (longest-first for every warp): the first reads of the primary warps are assigned statically — the
     * longest reads, spread over the SMs — and every sub-partition knows from the start whether its primary has a long
     * one, so that no secondary warp starts beside a long read before its primary has had the time to say so */
    const int32_t n_pri = 4 * (int32_t)gridDim.x;
    bool first_pull = (policy == 1) && primary;
    if (threadIdx.x < 4) {
        int flag = 0;
        if (policy == 1) {
            const int32_t r = first + (int32_t)threadIdx.x * (int32_t)gridDim.x + (int32_t)blockIdx.x;
            if (r < n_reads) flag = (reads[r].n_events + reads[r].n_kmers + 2 > long_thr) ? 1 : 0;
        }
        long_flag[threadIdx.x] = flag;
    }
    __syncthreads();
    const double NEG = abea_neg_inf_d();

    for (;;) {
        if (!primary) { /* let the long read on this sub-partition run alone */
            unsigned ns = 1024u;
            for (;;) {
                int busy = 0;
                if (lane == 0) busy = atomicOr(&long_flag[slot], 0);
                if (__shfl_sync(ABEA_FULL, busy, 0) == 0) break;
                abea_backoff(ns);
            }
        }
        int32_t ridx = n_reads;
        if (first_pull) {
            first_pull = false;
            ridx = first + slot * (int32_t)gridDim.x + (int32_t)blockIdx.x;
        } else {
            if (lane == 0) { /* reads [0, first) are filled by the wide kernel */
                if (policy == 1) {
                    ridx = first + n_pri + atomicAdd(queue + 1, 1);
                } else if (first + atomicAdd(queue, 1) < n_reads) {
                    ridx = primary ? first + atomicAdd(queue + 1, 1) : n_reads - 1 - atomicAdd(queue + 2, 1);
                }
            }
            ridx = __shfl_sync(ABEA_FULL, ridx, 0);
        }
        if (ridx >= n_reads || ridx < 0) break;
        const abea_read_t rd = reads[ridx];
        /* streaming: the read starts as soon as its first events (0..95: the first window and two chunks of the
         * ring) have landed, and from there chases the loader piece by piece (abea_fill_step) */
        if (STREAM) {
            const uint32_t* w = io.ready + (int64_t)ABEA_READY_WORDS * ridx;
            const int32_t got = abea_wait_landed_events(w, rd.n_events > 96 ? 95 : rd.n_events - 1, io.stalled);
            __syncwarp();
            if (lane == 0) {
                sm->ready_w = w;
                sm->ev_landed = got;
            }
        }
        __syncwarp();
        /* each instantiation takes only the reads validated for its arithmetic. Streaming: the loader clears the FAST
         * bit of a read when an out-of-range event mean passes through it, always BEFORE it publishes that piece; a
         * read whose bit is cleared after this test is filled here in vain and again, last, by the EXACT instantiation
         * (which runs after this kernel and overwrites every result of the read). */
        if (((abea_ld_acquire_u32(read_flags + ridx) & ABEA_READ_FAST) != 0u) != FAST) {
            if (primary && lane == 0) atomicExch(&long_flag[slot], 0);
            continue;
        }

        if (primary && lane == 0) atomicExch(&long_flag[slot], (rd.n_events + rd.n_kmers + 2 > long_thr) ? 1 : 0);
        const long long t_start = abea_clock();
        if (lane == 0) results[ridx].start_us = abea_now_us();
        abea_fill_ctx_t cx;
        cx.E = rd.n_events;
        cx.K = rd.n_kmers;
        cx.NB = cx.E + cx.K + 2; /* the host rejects reads with E + K + 2 >= 2^31 */
        cx.ev = means + rd.ev_off;
        cx.kpr = kparams + rd.kp_off;
        cx.tr = trace + rd.trace_off;
        cx.lp_stay = rd.lp_stay;
        cx.lp_step = rd.lp_step;
        cx.lp_skip = cst.lp_skip;
        cx.lp_trim = cst.lp_trim;
        /* band 1 geometry (reference src/align.c:277-279): e0=49,k0=-51 ; band 1 = move_down(band 0) */
        cx.eb = ABEA_W / 2;
        cx.kb = -1 - ABEA_W / 2;
        cx.b = 2;
        cx.safe = 0;
        cx.prev_right = false; /* band 1 was a down move */
        cx.best_s = NEG;
        cx.best_e = 0x7fffffff;
        cx.tword = (lane == (ABEA_W / 2) / ABEA_CPL) ? (ABEA_FROM_U << (8 + 2 * ((ABEA_W / 2) % ABEA_CPL))) : 0u;
        cx.eb_keep = (lane == 25) ? (ABEA_W / 2 - 1) : ((lane == 26) ? ABEA_W / 2 : 0);

        /* stage the chunks holding the first entering event (eb+1) and k-mer (kb+100), prefetch the ones after */
        __syncwarp();
        const int32_t ec0 = (cx.eb + 1) >> 5, kc0 = (cx.kb + ABEA_W) >> 5;
        abea_stage_events(sm, cx.ev, ec0, cx.E, lane);
        abea_stage_kparams(sm, cx.kpr, kc0, cx.K, lane);
        abea_cp_async_wait_all();
        __syncwarp();
        abea_stage_events(sm, cx.ev, ec0 + 1, cx.E, lane);
        abea_stage_kparams(sm, cx.kpr, kc0 + 1, cx.K, lane);
        cx.x_next = sm->ev[(cx.eb + 1) & (ABEA_RING - 1)];
        cx.kp_next = sm->kp[(cx.kb + ABEA_W) & (ABEA_RING - 1)];

        /* sliding windows for band 1: x[c] = mean of event eb-o ; kp[c] = params of k-mer kb+o, o = 4*lane+c */
        float x[ABEA_CPL];
        float4 kp[ABEA_CPL];
        abea_band_t P, Q; /* P = band 1, Q = band 0 */
#pragma unroll
        for (int c = 0; c < ABEA_CPL; c++) {
            int o = ABEA_CPL * lane + c;
            x[c] = abea_load_event_mean(cx.ev, cx.eb - o, cx.E);
            kp[c] = abea_load_kparam(cx.kpr, cx.kb + o, cx.K);
            Q.R[c] = (o == ABEA_W / 2) ? 0.0 : NEG;                                        /* src/align.c:284 */
            P.R[c] = (o == ABEA_W / 2) ? (double)__double2float_rn(cx.lp_trim) : NEG;       /* src/align.c:290 */
        }
        /* the only finite cells of bands 0 and 1 sit at offset 50 (lane 12, c = 2): every halo is -inf */
        P.lo = P.hi = Q.lo = Q.hi = NEG;

        /* move of band 2: both extreme cells of band 1 are -inf, so it alternates: band 2 is even -> down */
        bool right = false;
        while (cx.b < cx.NB) {
            right = abea_fill_step<FAST, STREAM>(cx, x, kp, P, Q, sm, right, lane, io.stalled); /* Q <- band b */
            if (cx.b >= cx.NB) break;
            right = abea_fill_step<FAST, STREAM>(cx, x, kp, Q, P, sm, right, lane, io.stalled); /* P <- band b */
        }
        abea_cp_async_wait_all(); /* nothing may still be landing in the ring when the next read reuses it */

        /* merge per-lane best end cells: max score, ties to the smaller event (first strict max in event order) */
        double best_s = cx.best_s;
        int32_t best_e = cx.best_e;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            double os = __shfl_xor_sync(ABEA_FULL, best_s, d);
            int32_t oe2 = __shfl_xor_sync(ABEA_FULL, best_e, d);
            if (os > best_s || (os == best_s && oe2 < best_e)) {
                best_s = os;
                best_e = oe2;
            }
        }
        const int32_t end_event = (best_e == 0x7fffffff) ? 0 : best_e;
        if (lane == 0) {
            results[ridx].end_score = __double2float_rn(best_s);
            results[ridx].end_event = end_event;
        }
        /* traceback + QC of this read by the same warp, while the other warps keep filling: the walk of a long read
         * is a serial chain too, and fusing it here takes it off the tail of the batch */
        __syncwarp();
        const long long t_fill = abea_clock();
        /* four walks per lane pay off once the stretches are long (RNA reads of 20 k events: 39.4 -> 35.7 ms at cfg4);
         * on reads of a few thousand bands the margins of 128 short stretches cost more than the round trips saved */
        const abea_stream_t io_r = abea_publishable<FAST, STREAM>(io, read_flags + ridx);
        if (cx.NB >= ABEA_TB_K4_MIN_BANDS)
            abea_traceback<FAST, 4>(rd, ridx, end_event, tb_ring, lane, means, kparams, trace, pairs, results, io_r);
        else
            abea_traceback<FAST, 1>(rd, ridx, end_event, tb_ring, lane, means, kparams, trace, pairs, results, io_r);
        if (lane == 0) {
            results[ridx].wide = 0;
            results[ridx].fill_cycles = t_fill - t_start;
            results[ridx].trace_cycles = abea_clock() - t_fill;
        }
        if (primary && lane == 0) atomicExch(&long_flag[slot], 0);
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Band fill, WIDE form: one CTA (4 warps) per read, ONE band cell per lane. A read is a serial chain of NB bands, so
 * the longest reads of a batch set the makespan; the reference sends them to CPU threads (src/f5c.cu:440-452), here
 * they get four warps instead of one, which cuts the per-band critical path (one cell instead of four per lane, no
 * register windows to slide). Warp w owns offsets 28w..28w+27 (warp 3: 84..99), so that four lanes form one word of
 * the same 128-B trace line layout the narrow kernel writes. Events and k-mer parameters of the whole band window live
 * in shared-memory rings (256 entries, 64-entry chunks staged by cp.async) and are simply re-addressed every band.
 * Every lane publishes its cell of the band to a shared-memory copy (double-buffered by band parity); neighbours and
 * the two cells of Suzuki's rule are read from it behind a split-phase mbarrier (below). Arithmetic is the same
 * templates as the narrow kernel.                                                                                  */

/* Split-phase CTA barrier (mbarrier in shared memory): a warp ARRIVES as soon as its boundary cells of the band are
 * published and WAITS only when it needs the other warps' cells at the top of the next band; everything in between
 * (next band's speculative emissions, shuffles, trace bookkeeping) overlaps the other warps' arrival. One arrival per
 * warp. The CPU emulator, which runs the warps of a block in lock-step, turns the wait into a plain block barrier. */
__device__ __forceinline__ void abea_mbar_init(uint64_t* bar, unsigned count) {
#ifndef ABEA_SIMT_EMU
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
#else
    (void)bar; (void)count;
#endif
}
__device__ __forceinline__ void abea_mbar_inval(uint64_t* bar) {
#ifndef ABEA_SIMT_EMU
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(a) : "memory");
#else
    (void)bar;
#endif
}
__device__ __forceinline__ void abea_mbar_arrive(uint64_t* bar) {
#ifndef ABEA_SIMT_EMU
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(a) : "memory");
#else
    (void)bar;
#endif
}
__device__ __forceinline__ void abea_mbar_wait(uint64_t* bar, unsigned parity) {
#ifndef ABEA_SIMT_EMU
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "ABEA_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra ABEA_MBAR_DONE;\n\t"
        "bra ABEA_MBAR_WAIT;\n"
        "ABEA_MBAR_DONE:\n\t}" ::"r"(a), "r"(parity) : "memory");
#else
    (void)bar; (void)parity;
    __syncthreads();
#endif
}

#define ABEA_WIDE_WARPS 4
#define ABEA_WIDE_SPAN 28   /* offsets per warp */
#define ABEA_WRING 256      /* ring entries (window of 100 + chunks in flight) */
#define ABEA_WCHUNK 64
#define ABEA_WBAND 120      /* entries of a published band: [0] = offset -1 (-inf), [1..100] = the cells, [101] = offset 100 (-inf),
                             * the rest is only ever read by lanes that own no cell */

struct abea_wide_smem_t {
    float ev[ABEA_WRING];
    float4 kp[ABEA_WRING];
    double band[2][ABEA_WBAND]; /* [band parity][1 + offset]: the scores of bands b-1 and b, read by every lane's neighbours */
    double red_s[ABEA_WIDE_WARPS];
    int32_t red_e[ABEA_WIDE_WARPS];
    int32_t ridx;
    uint64_t bar; /* the per-band split-phase barrier */
};

__device__ __forceinline__ void abea_wide_stage(abea_wide_smem_t* sm, const float* __restrict__ ev,
                                                const float4* __restrict__ kpr, int32_t echunk, int32_t kchunk,
                                                int32_t E, int32_t K, int tid) {
    if (tid < ABEA_WCHUNK) {
        if (echunk >= 0) {
            int32_t e = echunk * ABEA_WCHUNK + tid;
            int32_t ec = e >= E ? E - 1 : e;
            abea_cp_async4(&sm->ev[e & (ABEA_WRING - 1)], &ev[ec]);
        }
        if (kchunk >= 0) {
            int32_t k = kchunk * ABEA_WCHUNK + tid;
            int32_t kc = k >= K ? K - 1 : k;
            abea_cp_async16(&sm->kp[k & (ABEA_WRING - 1)], &kpr[kc]);
        }
    }
}

/* A lane's view of one band: its own cell and the cells at offset-1 / offset+1 (read from the published band). */
struct abea_wband_t {
    double R, lo, hi;
};

/* Per-read state of a wide lane. Everything but the scores, the window values and the trace bits is CTA-uniform. */
struct abea_wide_ctx_t {
    const float* ev;
    const float4* kpr;
    uint32_t* tr;
    const uint32_t* ready_w; /* streaming: the read's words of d_ready */
    uint32_t* stalled;
    double lp_stay, lp_step, lp_skip, lp_trim;
    int32_t E, K, NB;
    int32_t eb, kb;          /* lower-left of the current band */
    int32_t b;               /* band being filled */
    int32_t safe;            /* further bands guaranteed to be interior */
    int32_t e_cnt, k_cnt;    /* down / right moves until the chunk in flight is entered (land it, stage the next) */
    int32_t echunk_hi, kchunk_hi; /* highest chunk staged; it may still be in flight */
    int32_t ev_landed;       /* streaming: leading events of the read known to be in d_events */
    uint32_t acc;            /* this lane's trace bits of the current 4-band group */
    int32_t eb_keep;         /* warp 3, lanes 28..31: event index of the group's bands */
    double best_s;           /* best end cell seen by this lane */
    int32_t best_e;
    /* the lane's event / k-mer of band b-1 and, speculatively, what it would hold after either move: a right move keeps
     * the event and takes k-mer kb+1+o, a down move keeps the k-mer and takes event eb+1-o */
    float x_cur, x_dn;
    float4 kp_cur, kp_rt;
    double lpd_rt, lpd_dn;   /* emission of the lane's cell of the next band after a right / down move */
    bool prev_right;
};

/* One band of the wide fill for band index b with b % 4 == Q. A holds band b-1 (its R in a register since the last
 * step, its neighbours read here from the published copy), B holds band b-2 and receives band b. */
template <bool FAST, bool STREAM, int Q>
__device__ __forceinline__ void abea_wide_step(abea_wide_ctx_t& cx, abea_wide_smem_t& sm, abea_wband_t& A, abea_wband_t& B,
                                               const int tid, const int o, const int oa, const bool active,
                                               const uint32_t sh0) {
    const double NEG = abea_neg_inf_d();
    constexpr int PAR = Q & 1;      /* parity of band b: where it is published, and the phase parity of "b-1 is published" */
    constexpr int PP = PAR ^ 1;     /* parity of band b-1 */
    /* ---- wait until every warp has published band b-1, then: neighbours, the two corner cells, this band's move ---- */
    abea_mbar_wait(&sm.bar, PAR);
    const double* bp = sm.band[PP];
    A.lo = bp[o];
    A.hi = bp[o + 2];
    const double ll = bp[1], ur = bp[ABEA_W];
    /* src/align.c:304-314: both corners -inf -> right iff b is odd, else right iff ll < ur. With both at -inf the
     * comparison is false, so for an even band it alone decides; for an odd one the two -inf tests are integer compares of
     * the high words (a band score is never NaN), off the FP64 pipe and in parallel with the comparison */
    bool right = ll < ur;
    if (PAR == 1) right = right || ((__double2hiint(ll) == (int)0xfff00000) && (__double2hiint(ur) == (int)0xfff00000));

    /* ---- the cell (same arithmetic templates as the narrow kernel), specialised by (this move, previous move) ---- */
    double Rn;
    uint32_t fr;
    if (right) {
        cx.kb += 1;
        cx.kp_cur = cx.kp_rt;
        /* keep the rings ahead of what the next band may touch (event eb+1, k-mer kb+100): when that index is the first
         * of the chunk in flight, land it and start the next one. Chunk c+1 reuses the ring slots of chunk c-3, which
         * left the 100-wide window long before. */
        if (--cx.k_cnt == 0) {
            cx.k_cnt = ABEA_WCHUNK;
            abea_cp_async_wait_all();
            __syncthreads();
            cx.kchunk_hi += 1;
            abea_wide_stage(&sm, cx.ev, cx.kpr, -1, cx.kchunk_hi, cx.E, cx.K, tid);
        }
        /* a right move changes only the k-mer a further right move would bring in: its load is issued here, ahead of
         * the cell; the event a down move would bring in (x_dn) is still the one loaded after the last down move */
        cx.kp_rt = sm.kp[(cx.kb + 1 + oa) & (ABEA_WRING - 1)];
        if (cx.prev_right) abea_wide_cell<FAST>(cx.lpd_rt, A.hi, A.R, B.hi, cx.lp_step, cx.lp_stay, cx.lp_skip, Rn, fr);
        else abea_wide_cell<FAST>(cx.lpd_rt, A.hi, A.R, B.R, cx.lp_step, cx.lp_stay, cx.lp_skip, Rn, fr);
    } else {
        cx.eb += 1;
        cx.x_cur = cx.x_dn;
        if (--cx.e_cnt == 0) {
            cx.e_cnt = ABEA_WCHUNK;
            abea_cp_async_wait_all();
            __syncthreads();
            cx.echunk_hi += 1;
            if (STREAM) { /* the chunk about to be staged may not have crossed PCIe yet */
                const int32_t last = (cx.echunk_hi * ABEA_WCHUNK + ABEA_WCHUNK - 1 < cx.E) ? cx.echunk_hi * ABEA_WCHUNK + ABEA_WCHUNK - 1 : cx.E - 1;
                if (last >= cx.ev_landed) cx.ev_landed = abea_wait_landed_events(cx.ready_w, last, cx.stalled);
            }
            abea_wide_stage(&sm, cx.ev, cx.kpr, cx.echunk_hi, -1, cx.E, cx.K, tid);
        }
        cx.x_dn = sm.ev[(cx.eb + 1 - oa) & (ABEA_WRING - 1)]; /* likewise: only the event of a further down move changes */
        if (cx.prev_right) abea_wide_cell<FAST>(cx.lpd_dn, A.R, A.lo, B.R, cx.lp_step, cx.lp_stay, cx.lp_skip, Rn, fr);
        else abea_wide_cell<FAST>(cx.lpd_dn, A.R, A.lo, B.lo, cx.lp_step, cx.lp_stay, cx.lp_skip, Rn, fr);
    }
    cx.prev_right = right;

    /* ---- band edges (validity window, trim column, end column); interior bands skip this ---- */
    if (cx.safe > 0) {
        cx.safe -= 1;
    } else {
        const int32_t kb = cx.kb, eb = cx.eb, E = cx.E, K = cx.K;
        const bool interior = (kb >= 0) && (kb + ABEA_W < K) && (eb >= ABEA_W - 1) && (eb <= E - 1);
        if (interior) {
            int32_t s1 = K - ABEA_W - 1 - kb, s2 = E - 1 - eb;
            cx.safe = s1 < s2 ? s1 : s2;
        } else {
            int32_t lo = -kb;
            if (eb - (E - 1) > lo) lo = eb - (E - 1);
            if (lo < 0) lo = 0;
            int32_t hi = K - kb;
            if (eb + 1 < hi) hi = eb + 1;
            if (hi > ABEA_W) hi = ABEA_W;
            const int32_t to = -1 - kb;
            const int32_t te = eb - to;
            const bool trim_in = (to >= 0) && (to < ABEA_W) && (te >= 0) && (te < E);
            const double trim_s = (double)__double2float_rn(__dmul_rn(cx.lp_trim, (double)(te + 1)));
            const int32_t oe = (K - 1) - kb;
            const bool valid = active && (o >= lo) && (o < hi);
            Rn = valid ? Rn : NEG;
            fr = valid ? fr : 0u;
            if (active && o == to && trim_in) {
                Rn = trim_s;
                fr = ABEA_FROM_U;
            }
            if (o == oe && valid) {
                const int32_t e = eb - o;
                double s = (double)__double2float_rn(__dadd_rn(Rn, __dmul_rn((double)(E - e), cx.lp_trim)));
                if (s > cx.best_s) {
                    cx.best_s = s;
                    cx.best_e = e;
                }
            }
        }
    }

    /* ---- publish the band; everything below is independent of the other warps and overlaps their arrival ---- */
    if (active) sm.band[PAR][o + 1] = Rn;
    __syncwarp(); /* orders the other lanes' stores before lane 0's arrival */
    if ((tid & 31) == 0) abea_mbar_arrive(&sm.bar);
    B.R = Rn;

    /* speculative emissions of band b+1 for both moves (lanes past offset 99 read the ring at offset 99 so that they
     * never touch a chunk that is still in flight) */
    cx.lpd_rt = (double)abea_emission_t<FAST>(cx.x_cur, cx.kp_rt);
    cx.lpd_dn = (double)abea_emission_t<FAST>(cx.x_dn, cx.kp_cur);

    /* trace: this lane's 2 bits go to bit 8q + 2(o&3) of word o>>2 of the 128-B line of the 4-band group */
    cx.acc += (fr << sh0) << (8 * Q);
    if (tid == 32 * 3 + 28 + Q) cx.eb_keep = cx.eb;
    if (Q == 3 || cx.b == cx.NB - 1) {
        uint32_t word = cx.acc | __shfl_xor_sync(ABEA_FULL, cx.acc, 1);
        word |= __shfl_xor_sync(ABEA_FULL, word, 2);
        uint32_t* line = cx.tr + (int64_t)(cx.b >> 2) * ABEA_TRACE_GROUP_WORDS;
        if (active && (tid & 3) == 0) line[o >> 2] = word;
        if (tid >= 32 * 3 + 28) line[ABEA_LANES + (tid - (32 * 3 + 28))] = (uint32_t)cx.eb_keep;
        cx.acc = 0u;
    }
    cx.b += 1;
}

template <bool FAST, bool STREAM>
__global__ void __launch_bounds__(32 * ABEA_WIDE_WARPS)
abea_fill_wide_kernel(const abea_read_t* __restrict__ reads, int32_t n_wide, const float* __restrict__ means,
                      const float4* __restrict__ kparams, const uint32_t* __restrict__ read_flags,
                      uint32_t* __restrict__ trace, abea_pair_t* __restrict__ pairs, abea_result_t* __restrict__ results,
                      abea_stream_t io, abea_consts_t cst, int32_t* __restrict__ queue) {
    __shared__ __align__(16) abea_wide_smem_t sm;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int w = tid >> 5;
    const int o = ABEA_WIDE_SPAN * w + lane;                       /* this lane's band offset */
    const bool active = (lane < ABEA_WIDE_SPAN) && (o < ABEA_W);
    const int oa = o < ABEA_W ? o : ABEA_W - 1;
    const uint32_t sh0 = 2u * (uint32_t)(o & 3);
    const double NEG = abea_neg_inf_d();
    for (int i = tid; i < 2 * ABEA_WBAND; i += 32 * ABEA_WIDE_WARPS) (&sm.band[0][0])[i] = NEG; /* incl. the -inf sentinels */
    bool bar_live = false;

    for (;;) {
        __syncthreads(); /* nobody is still waiting on, or arriving at, the previous read's barrier */
        if (tid == 0) {
            sm.ridx = atomicAdd(queue, 1);
            /* a fresh barrier per read: its phase parity then follows the band index (band b waits for parity b & 1) */
            if (bar_live) abea_mbar_inval(&sm.bar);
            abea_mbar_init(&sm.bar, ABEA_WIDE_WARPS);
            bar_live = true;
        }
        __syncthreads();
        const int32_t ridx = sm.ridx;
        if (ridx >= n_wide) break;
        const abea_read_t rd = reads[ridx];
        abea_wide_ctx_t cx;
        /* streaming: start when the first three chunks of the ring (events 0..191) have landed, then chase the loader */
        cx.ev_landed = 0x7fffffff;
        cx.ready_w = nullptr;
        cx.stalled = io.stalled;
        if (STREAM) {
            cx.ready_w = io.ready + (int64_t)ABEA_READY_WORDS * ridx;
            cx.ev_landed = abea_wait_landed_events(cx.ready_w, rd.n_events > 3 * ABEA_WCHUNK ? 3 * ABEA_WCHUNK - 1 : rd.n_events - 1,
                                                   io.stalled);
        }
        __syncthreads();
        if (((abea_ld_acquire_u32(read_flags + ridx) & ABEA_READ_FAST) != 0u) != FAST) continue;

        const long long t_start = abea_clock();
        if (tid == 0) results[ridx].start_us = abea_now_us();
        cx.E = rd.n_events;
        cx.K = rd.n_kmers;
        cx.NB = cx.E + cx.K + 2;
        cx.ev = means + rd.ev_off;
        cx.kpr = kparams + rd.kp_off;
        cx.tr = trace + rd.trace_off;
        cx.lp_stay = rd.lp_stay;
        cx.lp_step = rd.lp_step;
        cx.lp_skip = cst.lp_skip;
        cx.lp_trim = cst.lp_trim;
        cx.eb = ABEA_W / 2;          /* band 1 (reference src/align.c:277-279) */
        cx.kb = -1 - ABEA_W / 2;
        cx.b = 2;
        cx.safe = 0;
        cx.prev_right = false;       /* band 1 was a down move */
        cx.best_s = NEG;
        cx.best_e = 0x7fffffff;
        /* stage the chunks covering events [0, eb+64) and k-mers [0, kb+99+64): chunk 0 and 1 of each (indices < 0 are
         * never valid cells; the ring slots they alias hold clamped garbage that the validity mask discards) */
        abea_wide_stage(&sm, cx.ev, cx.kpr, 0, 0, cx.E, cx.K, tid);
        abea_wide_stage(&sm, cx.ev, cx.kpr, 1, 1, cx.E, cx.K, tid);
        /* band 1 as published: -inf but for the trim cell at offset 50 (src/align.c:290) */
        const double R1 = (o == ABEA_W / 2) ? (double)__double2float_rn(cx.lp_trim) : NEG;
        if (active) sm.band[1][o + 1] = R1;
        abea_cp_async_wait_all();
        __syncthreads();
        if (lane == 0) abea_mbar_arrive(&sm.bar); /* "band 1 is published": the first wait (band 2, parity 0) falls through */
        abea_wide_stage(&sm, cx.ev, cx.kpr, 2, 2, cx.E, cx.K, tid); /* in flight while chunks 0 and 1 are consumed */
        cx.echunk_hi = 2;
        cx.kchunk_hi = 2;
        /* event eb+1 / k-mer kb+100 first enter the chunk in flight (chunk 2) at index 128 */
        cx.e_cnt = 2 * ABEA_WCHUNK - (cx.eb + 1);
        cx.k_cnt = 2 * ABEA_WCHUNK - (cx.kb + ABEA_W);

        abea_wband_t P, Qb; /* P = band 1, Qb = band 0 (0 at offset 50, src/align.c:284); P's neighbours are read in step 2 */
        P.R = active ? R1 : NEG;
        P.lo = P.hi = NEG;
        Qb.R = (o == ABEA_W / 2) ? 0.0 : NEG;
        Qb.lo = (o - 1 == ABEA_W / 2) ? 0.0 : NEG;
        Qb.hi = (o + 1 == ABEA_W / 2) ? 0.0 : NEG;

        cx.acc = (o == ABEA_W / 2) ? (ABEA_FROM_U << (8 + 2 * (o & 3))) : 0u; /* band 1's trim cell */
        cx.eb_keep = (tid == 32 * 3 + 28) ? (ABEA_W / 2 - 1) : ((tid == 32 * 3 + 29) ? ABEA_W / 2 : 0);

        cx.x_cur = sm.ev[(cx.eb - oa) & (ABEA_WRING - 1)];
        cx.kp_cur = sm.kp[(cx.kb + oa) & (ABEA_WRING - 1)];
        cx.x_dn = sm.ev[(cx.eb + 1 - oa) & (ABEA_WRING - 1)];
        cx.kp_rt = sm.kp[(cx.kb + 1 + oa) & (ABEA_WRING - 1)];
        cx.lpd_rt = (double)abea_emission_t<FAST>(cx.x_cur, cx.kp_rt);
        cx.lpd_dn = (double)abea_emission_t<FAST>(cx.x_dn, cx.kp_cur);

        /* bands 2, 3, 4, 5, ...: four steps per trip so that the trace position and the barrier parity are static, and
         * the two band registers swap roles instead of being copied */
        for (;;) {
            abea_wide_step<FAST, STREAM, 2>(cx, sm, P, Qb, tid, o, oa, active, sh0); /* Qb <- band b */
            if (cx.b >= cx.NB) break;
            abea_wide_step<FAST, STREAM, 3>(cx, sm, Qb, P, tid, o, oa, active, sh0); /* P <- band b */
            if (cx.b >= cx.NB) break;
            abea_wide_step<FAST, STREAM, 0>(cx, sm, P, Qb, tid, o, oa, active, sh0);
            if (cx.b >= cx.NB) break;
            abea_wide_step<FAST, STREAM, 1>(cx, sm, Qb, P, tid, o, oa, active, sh0);
            if (cx.b >= cx.NB) break;
        }
        abea_cp_async_wait_all();
        __syncthreads(); /* the trace lines of the whole read are written before warp 0 walks them */

        /* best end cell over the CTA: max score, ties to the smaller event */
        double best_s = cx.best_s;
        int32_t best_e = cx.best_e;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            double os = __shfl_xor_sync(ABEA_FULL, best_s, d);
            int32_t oe2 = __shfl_xor_sync(ABEA_FULL, best_e, d);
            if (os > best_s || (os == best_s && oe2 < best_e)) { best_s = os; best_e = oe2; }
        }
        if (lane == 0) { sm.red_s[w] = best_s; sm.red_e[w] = best_e; }
        __syncthreads();
        if (w == 0) { /* warp 0 merges and then walks the trace; the k-mer ring (4 KB) becomes its trace ring */
            for (int i = 1; i < ABEA_WIDE_WARPS; i++) {
                double os = sm.red_s[i];
                int32_t oe2 = sm.red_e[i];
                if (os > best_s || (os == best_s && oe2 < best_e)) { best_s = os; best_e = oe2; }
            }
            const int32_t end_event = (best_e == 0x7fffffff) ? 0 : best_e;
            if (lane == 0) {
                results[ridx].end_score = __double2float_rn(best_s);
                results[ridx].end_event = end_event;
            }
            const long long t_fill = abea_clock();
            abea_traceback<FAST, 4>(rd, ridx, end_event, (uint32_t*)sm.kp, lane, means, kparams, trace, pairs, results,
                                    abea_publishable<FAST, STREAM>(io, read_flags + ridx));
            if (lane == 0) {
                results[ridx].wide = 1;
                results[ridx].fill_cycles = t_fill - t_start;
                results[ridx].trace_cycles = abea_clock() - t_fill;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Dense result list for consumers that move results between GPUs (the NCCL gather of a multi-GPU driver): the pair
 * lists of a batch leave the alignment in the caller's capacity layout (read i at the prefix sum of E+L, n_pairs[i]
 * of its E+L slots used); these two kernels pack them back to back. offsets[i] = sum of n_pairs[0..i), offsets[n] =
 * the total: one block, each thread scans a contiguous slice, the slice sums are scanned through shared memory. */
#define ABEA_SCAN_THREADS 1024
__global__ void __launch_bounds__(ABEA_SCAN_THREADS)
abea_pair_offsets_kernel(const int32_t* __restrict__ n_pairs, int32_t n, int64_t* __restrict__ offsets) {
    __shared__ int64_t part[ABEA_SCAN_THREADS];
    const int t = threadIdx.x;
    const int32_t per = (n + ABEA_SCAN_THREADS - 1) / ABEA_SCAN_THREADS;
    const int32_t i0 = t * per, i1 = (i0 + per < n) ? i0 + per : n;
    int64_t sum = 0;
    for (int32_t i = i0; i < i1; i++) sum += n_pairs[i] > 0 ? n_pairs[i] : 0;
    part[t] = sum;
    __syncthreads();
    if (t == 0) { /* 1024 additions: not worth a parallel scan */
        int64_t acc = 0;
        for (int j = 0; j < ABEA_SCAN_THREADS; j++) {
            const int64_t v = part[j];
            part[j] = acc;
            acc += v;
        }
        offsets[n] = acc;
    }
    __syncthreads();
    int64_t acc = part[t];
    for (int32_t i = i0; i < i1; i++) {
        offsets[i] = acc;
        acc += n_pairs[i] > 0 ? n_pairs[i] : 0;
    }
}

/* one warp per read (grid-stride): dst[offsets[i] .. +n_pairs[i]) = pairs[cap_ptr[i] .. +n_pairs[i]) */
__global__ void abea_compact_pairs_kernel(const abea_pair_t* __restrict__ pairs, const int64_t* __restrict__ cap_ptr,
                                          const int32_t* __restrict__ n_pairs, const int64_t* __restrict__ offsets,
                                          int32_t n, abea_pair_t* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int32_t warps = (int32_t)((gridDim.x * blockDim.x) >> 5);
    for (int32_t i = (int32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < n; i += warps) {
        const int32_t np = n_pairs[i];
        const abea_pair_t* src = pairs + cap_ptr[i];
        abea_pair_t* d = dst + offsets[i];
        for (int32_t j = lane; j < np; j += 32) d[j] = src[j];
    }
}

/* The path codes of a whole batch from the pair lists the traceback left in d_pairs (capacity layout): one warp per read.
 * This is how abea_device_codes produces them — on demand, off every read's critical path (emitting them from the
 * traceback cost the longest read of a batch 9 cycles per step: cfg3 38.6 -> 39.1 ms). */
__global__ void abea_pairs_to_codes_kernel(const abea_pair_t* __restrict__ pairs, const int64_t* __restrict__ cap_ptr,
                                           const int32_t* __restrict__ n_pairs, int32_t n, abea_code_t* __restrict__ codes) {
    const int lane = threadIdx.x & 31;
    const int32_t warps = (int32_t)((gridDim.x * blockDim.x) >> 5);
    for (int32_t i = (int32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < n; i += warps) {
        const int32_t np = n_pairs[i];
        if (np > 0) abea_emit_codes(nullptr, codes + abea_code_offset(cap_ptr[i], i), pairs + cap_ptr[i], np, lane);
    }
}

/* Path codes back to a dense pair list on the device (the receiving end of the multi-GPU exchange, or any consumer
 * that was handed codes): one warp per read, 32 steps per trip — lane t's pair is the base of the word plus the number
 * of set bits below bit t+1 in each plane. dst[offsets[i] .. +n_pairs[i]) = read i's pairs. */
__global__ void abea_expand_codes_kernel(const abea_code_t* __restrict__ codes, const int64_t* __restrict__ cap_ptr,
                                         const int32_t* __restrict__ n_pairs, const int64_t* __restrict__ offsets,
                                         int32_t n, abea_pair_t* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int32_t warps = (int32_t)((gridDim.x * blockDim.x) >> 5);
    for (int32_t i = (int32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < n; i += warps) {
        const int32_t np = n_pairs[i];
        if (np <= 0) continue;
        const abea_code_t* w = codes + abea_code_offset(cap_ptr[i], i);
        abea_pair_t* d = dst + offsets[i];
        int32_t k = (int32_t)w[0].a, e = (int32_t)w[0].b;
        if (lane == 0) {
            abea_pair_t p;
            p.ref_pos = k;
            p.read_pos = e;
            d[0] = p;
        }
        const uint32_t below = 0xffffffffu >> (31 - lane); /* bits 0..lane */
        for (int32_t j0 = 1; j0 < np; j0 += 32) {
            const abea_code_t c = w[1 + ((j0 - 1) >> 5)];
            const int32_t j = j0 + lane;
            if (j < np) {
                abea_pair_t p;
                p.ref_pos = k + __popc(c.a & below);
                p.read_pos = e + __popc(c.b & below);
                d[j] = p;
            }
            k += __popc(c.a);
            e += __popc(c.b);
        }
    }
}
