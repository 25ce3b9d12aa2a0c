/* abea_kernels.cuh — sm_100a device code for adaptive banded event alignment (ABEA).
 *
 * Replaces the reference's three kernels (align_kernel_pre_2d / align_kernel_core_2d_shm / align_kernel_post,
 * reference src/align.cu:149-749) with a different decomposition:
 *
 *   abea_kmer_params_kernel  one thread per k-mer: rank -> model gather -> {scaled mean, stdv, -0.918938-log stdv}
 *                            (the reference does this with ONE thread per read, src/align.cu:203-209)
 *   abea_fill_kernel         one WARP per read, 4 band cells per lane, whole band state in registers; neighbours
 *                            through warp shuffles; 2-bit packed trace, one coalesced 128-B store per 4 bands;
 *                            last-column arg-max folded into the fill; persistent warps pulling reads longest-first
 *   abea_traceback_kernel    one warp per read walking the packed trace, pairs written in ascending order
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

#define ABEA_FROM_D 0u /* reference src/align.c:194-196 */
#define ABEA_FROM_U 1u
#define ABEA_FROM_L 2u

/* Per-read descriptor built by the host packer (abea_host.cu), in scheduling (longest-first) order. */
struct abea_read_t {
    int64_t seq_off;    /* first base in d_seq */
    int64_t ev_off;     /* first event in d_events (AoS abea_event_t) */
    int64_t kp_off;     /* first k-mer in d_kparams */
    int64_t trace_off;  /* first 32-bit word of this read's trace in d_trace */
    int64_t pair_off;   /* first pair slot in d_pairs (capacity pair_cap) */
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
    int32_t pair_start;  /* pairs live at d_pairs[pair_off + pair_start .. + n_aligned), ascending */
    int32_t max_gap;
};

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
/* k-mer parameter cache: kparams[kp_off + i] for k-mer i of each read.                                           */

__global__ void abea_kmer_params_kernel(const abea_read_t* __restrict__ reads, int32_t n_reads,
                                        const uint8_t* __restrict__ seq, const abea_model_t* __restrict__ model,
                                        uint32_t kmer_size, float4* __restrict__ kparams, int64_t total_kmers) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_kmers; idx += stride) {
        /* reads are laid out in kp_off order: find the read owning flat k-mer idx */
        int32_t lo = 0, hi = n_reads - 1;
        while (lo < hi) {
            int32_t mid = (lo + hi + 1) >> 1;
            if (reads[mid].kp_off <= idx) lo = mid; else hi = mid - 1;
        }
        const abea_read_t rd = reads[lo];
        int64_t i = idx - rd.kp_off;
        if (i >= rd.n_kmers) continue; /* padding slot */
        const uint8_t* s = seq + rd.seq_off + i;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < kmer_size; j++) rank = (rank << 2) | abea_base_rank(s[j]);
        abea_model_t m = model[rank];
        float4 kp;
        kp.x = __fadd_rn(__fmul_rn(rd.scale, m.level_mean), rd.shift); /* src/align.c:137-138 */
        kp.y = m.level_stdv;
        kp.z = __fsub_rn(-0.918938f, m.level_log_stdv);               /* src/align.c:111-113 */
        kp.w = 0.0f;
        kparams[idx] = kp;
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Band fill. One warp per read; lane j owns band offsets 4j..4j+3 (lanes 25..31 own nothing and only help with
 * loads and the trace line). Band b has lower-left (eb, kb); cell at offset o is (event eb-o, k-mer kb+o)
 * (reference src/align.c:156-161).                                                                               */

__device__ __forceinline__ float abea_neg_inf() { return __int_as_float(0xff800000); }

__device__ __forceinline__ float abea_load_event_mean(const abea_event_t* __restrict__ ev, int32_t e, int32_t E) {
    e = e < 0 ? 0 : (e >= E ? E - 1 : e);
    return ev[e].mean;
}

__device__ __forceinline__ float4 abea_load_kparam(const float4* __restrict__ kp, int32_t k, int32_t K) {
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    return kp[k];
}

__global__ void __launch_bounds__(128)
abea_fill_kernel(const abea_read_t* __restrict__ reads, int32_t n_reads, const abea_event_t* __restrict__ events,
                 const float4* __restrict__ kparams, uint32_t* __restrict__ trace, abea_result_t* __restrict__ results,
                 abea_consts_t cst, int32_t* __restrict__ queue) {
    const int lane = threadIdx.x & 31;
    const float NEG = abea_neg_inf();

    for (;;) {
        int32_t ridx = 0;
        if (lane == 0) ridx = atomicAdd(queue, 1);
        ridx = __shfl_sync(ABEA_FULL, ridx, 0);
        if (ridx >= n_reads) break;

        const abea_read_t rd = reads[ridx];
        const int32_t E = rd.n_events, K = rd.n_kmers;
        const int64_t NB = (int64_t)E + (int64_t)K + 2;
        const abea_event_t* __restrict__ ev = events + rd.ev_off;
        const float4* __restrict__ kpr = kparams + rd.kp_off;
        uint32_t* __restrict__ tr = trace + rd.trace_off;
        const double lp_stay = rd.lp_stay, lp_step = rd.lp_step, lp_skip = cst.lp_skip, lp_trim = cst.lp_trim;

        /* band 1 geometry (reference src/align.c:277-279): e0=49,k0=-51 ; band 1 = move_down(band 0) */
        int32_t eb = ABEA_W / 2, kb = -1 - ABEA_W / 2;

        /* sliding windows for band 1: x[c] = mean of event eb-o ; kp[c] = params of k-mer kb+o, o = 4*lane+c */
        float x[ABEA_CPL];
        float4 kp[ABEA_CPL];
#pragma unroll
        for (int c = 0; c < ABEA_CPL; c++) {
            int o = ABEA_CPL * lane + c;
            x[c] = abea_load_event_mean(ev, eb - o, E);
            kp[c] = abea_load_kparam(kpr, kb + o, K);
        }
        /* register chunk buffers for the elements that enter the window: lane i holds event (ebase+i) and
         * k-mer (kbase+i); the element needed next is fetched with one shuffle, a chunk is refilled every 32 moves */
        int32_t ebase = eb + 1;           /* next event to enter at offset 0 is eb+1 */
        int32_t kbase = kb + ABEA_W;      /* next k-mer to enter at offset 99 is kb+100 */
        float evbuf = abea_load_event_mean(ev, ebase + lane, E);
        float evnext = abea_load_event_mean(ev, ebase + 32 + lane, E);
        float4 kbuf = abea_load_kparam(kpr, kbase + lane, K);
        float4 knext = abea_load_kparam(kpr, kbase + 32 + lane, K);

        /* scores of band b-1 (S1) and b-2 (S2) */
        float S1[ABEA_CPL], S2[ABEA_CPL];
#pragma unroll
        for (int c = 0; c < ABEA_CPL; c++) {
            int o = ABEA_CPL * lane + c;
            S2[c] = (o == ABEA_W / 2) ? 0.0f : NEG;                           /* band 0: start cell (src/align.c:284) */
            S1[c] = (o == ABEA_W / 2) ? __double2float_rn(lp_trim) : NEG;      /* band 1: first trim (src/align.c:290) */
        }
        float edge_prev = NEG;   /* neighbour-lane score of band b-2 fetched one step earlier */
        bool prev_right = false; /* band 1 was a down move */

        /* trace line accumulators for the current group of 4 bands */
        uint32_t tword = (lane == (ABEA_W / 2) / ABEA_CPL) ? (ABEA_FROM_U << (8 + 2 * ((ABEA_W / 2) % ABEA_CPL))) : 0u;
        int32_t eb_keep = (lane == 25) ? (ABEA_W / 2 - 1) : ((lane == 26) ? ABEA_W / 2 : 0);

        /* best end cell (reference src/align.c:424-445), kept per lane, merged at the end */
        float best_s = NEG;
        int32_t best_e = 0x7fffffff;

        for (int64_t b = 2; b < NB; b++) {
            /* --- Suzuki's rule on band b-1's two extreme cells (reference src/align.c:304-322) --- */
            float ll = __shfl_sync(ABEA_FULL, S1[0], 0);
            float ur = __shfl_sync(ABEA_FULL, S1[ABEA_CPL - 1], ABEA_LANES - 1);
            bool right = (ll == NEG && ur == NEG) ? ((b & 1) == 1) : (ll < ur);

            float up[ABEA_CPL], left[ABEA_CPL], diag[ABEA_CPL];
            float edge;
            if (right) {
                kb += 1;
                /* k-mer window slides towards lower offsets; the new k-mer kb+99 enters at offset 99 */
                float4 in = kbuf;
                in.x = __shfl_sync(ABEA_FULL, kbuf.x, (kb + ABEA_W - 1) - kbase);
                in.y = __shfl_sync(ABEA_FULL, kbuf.y, (kb + ABEA_W - 1) - kbase);
                in.z = __shfl_sync(ABEA_FULL, kbuf.z, (kb + ABEA_W - 1) - kbase);
                float4 nb;
                nb.x = __shfl_down_sync(ABEA_FULL, kp[0].x, 1);
                nb.y = __shfl_down_sync(ABEA_FULL, kp[0].y, 1);
                nb.z = __shfl_down_sync(ABEA_FULL, kp[0].z, 1);
                nb.w = 0.0f;
                in.w = 0.0f;
#pragma unroll
                for (int c = 0; c < ABEA_CPL - 1; c++) kp[c] = kp[c + 1];
                kp[ABEA_CPL - 1] = (lane == ABEA_LANES - 1) ? in : nb;
                if ((kb + ABEA_W) - kbase == 32) { /* chunk exhausted */
                    kbase += 32;
                    kbuf = knext;
                    knext = abea_load_kparam(kpr, kbase + 32 + lane, K);
                }
                /* neighbours: up = band b-1 at o+1, left = band b-1 at o */
                edge = __shfl_down_sync(ABEA_FULL, S1[0], 1);
                if (lane >= ABEA_LANES - 1) edge = NEG;
#pragma unroll
                for (int c = 0; c < ABEA_CPL; c++) {
                    up[c] = (c < ABEA_CPL - 1) ? S1[c + 1] : edge;
                    left[c] = S1[c];
                }
                if (prev_right) { /* right,right: diag = band b-2 at o+1 */
#pragma unroll
                    for (int c = 0; c < ABEA_CPL; c++) diag[c] = (c < ABEA_CPL - 1) ? S2[c + 1] : edge_prev;
                } else {          /* down,right: diag = band b-2 at o */
#pragma unroll
                    for (int c = 0; c < ABEA_CPL; c++) diag[c] = S2[c];
                }
            } else {
                eb += 1;
                /* event window slides towards higher offsets; the new event eb enters at offset 0 */
                float in = __shfl_sync(ABEA_FULL, evbuf, eb - ebase);
                float nb = __shfl_up_sync(ABEA_FULL, x[ABEA_CPL - 1], 1);
#pragma unroll
                for (int c = ABEA_CPL - 1; c > 0; c--) x[c] = x[c - 1];
                x[0] = (lane == 0) ? in : nb;
                if ((eb + 1) - ebase == 32) {
                    ebase += 32;
                    evbuf = evnext;
                    evnext = abea_load_event_mean(ev, ebase + 32 + lane, E);
                }
                /* neighbours: up = band b-1 at o, left = band b-1 at o-1 */
                edge = __shfl_up_sync(ABEA_FULL, S1[ABEA_CPL - 1], 1);
                if (lane == 0) edge = NEG;
#pragma unroll
                for (int c = 0; c < ABEA_CPL; c++) {
                    up[c] = S1[c];
                    left[c] = (c > 0) ? S1[c - 1] : edge;
                }
                if (prev_right) { /* right,down: diag = band b-2 at o */
#pragma unroll
                    for (int c = 0; c < ABEA_CPL; c++) diag[c] = S2[c];
                } else {          /* down,down: diag = band b-2 at o-1 */
#pragma unroll
                    for (int c = 0; c < ABEA_CPL; c++) diag[c] = (c > 0) ? S2[c - 1] : edge_prev;
                }
            }

            /* --- cells --- */
            float Sn[ABEA_CPL];
            uint32_t fr[ABEA_CPL];
#pragma unroll
            for (int c = 0; c < ABEA_CPL; c++) {
                float lp = abea_emission(x[c], kp[c].x, kp[c].y, kp[c].z);
                abea_cell(lp, up[c], left[c], diag[c], lp_step, lp_stay, lp_skip, Sn[c], fr[c]);
            }

            /* --- band edges: validity window, trim column, end column (interior bands skip all of this) --- */
            const bool interior = (kb >= 0) && (kb + ABEA_W < K) && (eb >= ABEA_W - 1) && (eb <= E - 1);
            if (!interior) {
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
                const float trim_s = __double2float_rn(__dmul_rn(lp_trim, (double)(te + 1)));
                /* end column: k-mer K-1 (reference src/align.c:429-445) */
                const int32_t oe = (K - 1) - kb;
#pragma unroll
                for (int c = 0; c < ABEA_CPL; c++) {
                    int32_t o = ABEA_CPL * lane + c;
                    bool valid = (o >= lo) && (o < hi);
                    Sn[c] = valid ? Sn[c] : NEG;
                    fr[c] = valid ? fr[c] : 0u;
                    if (o == to && trim_in) {
                        Sn[c] = trim_s;
                        fr[c] = ABEA_FROM_U;
                    }
                    if (o == oe && valid) {
                        int32_t e = eb - o;
                        float s = __double2float_rn(__dadd_rn((double)Sn[c], __dmul_rn((double)(E - e), lp_trim)));
                        if (s > best_s) {
                            best_s = s;
                            best_e = e;
                        }
                    }
                }
            }

            /* --- trace: 2 bits per cell, one byte per lane per band, one 128-B line per 4 bands --- */
            uint32_t byte = fr[0] | (fr[1] << 2) | (fr[2] << 4) | (fr[3] << 6);
            const int q = (int)(b & 3);
            tword |= byte << (8 * q);
            if (lane == ABEA_LANES + q) eb_keep = eb;
            if (q == 3 || b == NB - 1) {
                tr[(b >> 2) * ABEA_TRACE_GROUP_WORDS + lane] = (lane < ABEA_LANES) ? tword : (uint32_t)eb_keep;
                tword = 0u;
            }

            /* rotate */
#pragma unroll
            for (int c = 0; c < ABEA_CPL; c++) {
                S2[c] = S1[c];
                S1[c] = Sn[c];
            }
            edge_prev = edge;
            prev_right = right;
        }

        /* merge per-lane best end cells: max score, ties to the smaller event (first strict max in event order) */
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            float os = __shfl_xor_sync(ABEA_FULL, best_s, d);
            int32_t oe2 = __shfl_xor_sync(ABEA_FULL, best_e, d);
            if (os > best_s || (os == best_s && oe2 < best_e)) {
                best_s = os;
                best_e = oe2;
            }
        }
        if (lane == 0) {
            results[ridx].end_score = best_s;
            results[ridx].end_event = (best_e == 0x7fffffff) ? 0 : best_e;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Traceback + QC (reference src/align.c:452-543). One warp per read, all lanes walk in lock-step over the 128-B
 * trace line of the current 4-band group held one word per lane; pairs are written from the END of the read's
 * capacity region backwards, so they come out ascending (no reversal pass).                                      */

__global__ void __launch_bounds__(128)
abea_traceback_kernel(const abea_read_t* __restrict__ reads, int32_t n_reads, const abea_event_t* __restrict__ events,
                      const float4* __restrict__ kparams, const uint32_t* __restrict__ trace,
                      abea_pair_t* __restrict__ pairs, abea_result_t* __restrict__ results, int32_t* __restrict__ queue) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        int32_t ridx = 0;
        if (lane == 0) ridx = atomicAdd(queue, 1);
        ridx = __shfl_sync(ABEA_FULL, ridx, 0);
        if (ridx >= n_reads) break;

        const abea_read_t rd = reads[ridx];
        const int32_t K = rd.n_kmers;
        const abea_event_t* __restrict__ ev = events + rd.ev_off;
        const float4* __restrict__ kpr = kparams + rd.kp_off;
        const uint32_t* __restrict__ tr = trace + rd.trace_off;
        abea_pair_t* __restrict__ out = pairs + rd.pair_off;

        int32_t ce = results[ridx].end_event;
        int32_t ck = K - 1;
        int32_t n = 0, gap = 0, max_gap = 0;
        int32_t last_k = ck;
        double sum = 0.0;
        int64_t cur_group = -1;
        uint32_t w = 0;
        while (ck >= 0 && ce >= 0) {
            /* emit (reference src/align.c:458-460) */
            if (lane == 0) {
                abea_pair_t p;
                p.ref_pos = ck;
                p.read_pos = ce;
                out[rd.pair_cap - 1 - n] = p;
            }
            n++;
            last_k = ck;
            float4 kp = kpr[ck];
            sum = __dadd_rn(sum, (double)abea_emission(ev[ce].mean, kp.x, kp.y, kp.z));

            int64_t b = (int64_t)ce + (int64_t)ck + 2;
            int64_t g = b >> 2;
            if (g != cur_group) {
                w = tr[g * ABEA_TRACE_GROUP_WORDS + lane];
                cur_group = g;
            }
            int q = (int)(b & 3);
            int32_t ebb = (int32_t)__shfl_sync(ABEA_FULL, w, ABEA_LANES + q);
            int32_t o = ebb - ce;
            /* an out-of-band start cell is undefined behaviour in the reference; stay in bounds */
            uint32_t tw = __shfl_sync(ABEA_FULL, w, (o >> 2) & 31);
            uint32_t from = (o >= 0 && o < ABEA_W) ? ((tw >> (8 * q + 2 * (o & 3))) & 3u) : ABEA_FROM_D;
            if (from == ABEA_FROM_D) {
                ck--; ce--; gap = 0;
            } else if (from == ABEA_FROM_U) {
                ce--; gap = 0;
            } else {
                ck--; gap++;
                max_gap = gap > max_gap ? gap : max_gap;
            }
        }
        /* QC (reference src/align.c:526-543) */
        double avg = sum / (double)n;
        bool spanned = (n > 0) && (last_k == 0);
        bool fail = (avg < -5.0) || !spanned || (max_gap > 50);
        if (lane == 0) {
            results[ridx].sum_emission = sum;
            results[ridx].n_aligned = n;
            results[ridx].n_pairs = fail ? 0 : n;
            results[ridx].pair_start = rd.pair_cap - n;
            results[ridx].max_gap = max_gap;
        }
    }
}
