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

#define ABEA_READ_FAST 1u /* read_flags bit: all inputs of the read are in the range the fast arithmetic is exact on */

#define ABEA_FROM_D 0u /* reference src/align.c:194-196 */
#define ABEA_FROM_U 1u
#define ABEA_FROM_L 2u

/* Per-read descriptor built by the host packer (abea_host.cu), in scheduling (longest-first) order. */
struct abea_read_t {
    int64_t seq_off;    /* first base in d_seq */
    int64_t ev_off;     /* first event in d_events (AoS abea_event_t) */
    int64_t evs_off;    /* running sum of n_events over the schedule (flat index space of abea_prepare_kernel) */
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
/* Per-read preparation: (1) k-mer parameter cache kparams[kp_off + i] = {scale*level_mean+shift, level_stdv,
 * -0.918938f - level_log_stdv, RN(1/level_stdv)} for k-mer i of each read (the reference builds this cache with ONE
 * thread per read, src/align.cu:203-209); (2) range validation of every value the fast arithmetic of the fill
 * kernel touches — a read keeps ABEA_READ_FAST in read_flags only if all of its event means and scaled level means
 * are 0 or within [2^-6, 2^16] in magnitude and all stdv are within [2^-6, 2^12] with a mantissa that is not all
 * ones (the one case Markstein's quotient correction excludes). read_flags must be pre-set to ABEA_READ_FAST.     */

__device__ __forceinline__ bool abea_sane_level(float v) {
    float a = fabsf(v);
    return (a == 0.0f) || (a >= 0.015625f && a <= 65536.0f); /* NaN fails both */
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
                                    uint32_t kmer_size, const abea_event_t* __restrict__ events,
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
        float x = events[rd.ev_off + (idx - rd.evs_off)].mean;
        if (!abea_sane_level(x)) atomicAnd(&read_flags[r], ~ABEA_READ_FAST);
    }
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

__device__ __forceinline__ float abea_load_event_mean(const abea_event_t* __restrict__ ev, int32_t e, int32_t E) {
    e = e < 0 ? 0 : (e >= E ? E - 1 : e);
    return ev[e].mean;
}

__device__ __forceinline__ float4 abea_load_kparam(const float4* __restrict__ kp, int32_t k, int32_t K) {
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    return kp[k];
}

/* x rounded to float precision, returned as a double. FAST: FP64-pipe magic constant; else hardware conversions. */
template <bool FAST>
__device__ __forceinline__ double abea_round_f32(double x) {
    if (FAST) {
        int hi = __double2hiint(x);
        double C = __hiloint2double((int)(((unsigned)hi & 0xfff00000u) + 0x01d80000u), 0);
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
__device__ __forceinline__ void abea_cell_d(float lp, double up, double left, double diag, double lp_step,
                                            double lp_stay, double lp_skip, double& score, uint32_t& from) {
    double lpd = (double)lp;
    double rd = abea_round_f32<FAST>(__dadd_rn(__dadd_rn(diag, lp_step), lpd));
    double ru = abea_round_f32<FAST>(__dadd_rn(__dadd_rn(up, lp_stay), lpd));
    double rl = abea_round_f32<FAST>(__dadd_rn(left, lp_skip));
    bool isU = ru >= rd;          /* (su > max) || (max == su) */
    double m = isU ? ru : rd;
    bool isL = rl >= m;
    score = isL ? rl : m;
    from = isL ? ABEA_FROM_L : (isU ? ABEA_FROM_U : ABEA_FROM_D);
}

struct abea_band_state {
    float x[ABEA_CPL];        /* event means of the lane's cells */
    float4 kp[ABEA_CPL];      /* k-mer parameters of the lane's cells */
    double R1[ABEA_CPL];      /* scores of band b-1 */
    double R2[ABEA_CPL];      /* scores of band b-2 */
    double halo_prev;         /* neighbour-lane score of band b-2 fetched one band earlier */
};

/* The four (previous move, this move) geometries, each fully specialised so that every neighbour is a fixed
 * register: RIGHT: up = b-1[o+1], left = b-1[o]; DOWN: up = b-1[o], left = b-1[o-1];
 * diag = b-2[o+1] (right,right), b-2[o-1] (down,down), else b-2[o]. */
template <bool FAST, bool RIGHT, bool PREV_RIGHT>
__device__ __forceinline__ void abea_band_cells(abea_band_state& st, int lane, double lp_step, double lp_stay,
                                                double lp_skip, double* Rn, uint32_t* fr) {
    const double NEG = abea_neg_inf_d();
    double halo;
    if (RIGHT) {
        halo = __shfl_down_sync(ABEA_FULL, st.R1[0], 1);
        if (lane >= ABEA_LANES - 1) halo = NEG;
    } else {
        halo = __shfl_up_sync(ABEA_FULL, st.R1[ABEA_CPL - 1], 1);
        if (lane == 0) halo = NEG;
    }
#pragma unroll
    for (int c = 0; c < ABEA_CPL; c++) {
        double up, left, diag;
        if (RIGHT) {
            up = (c < ABEA_CPL - 1) ? st.R1[c + 1 < ABEA_CPL ? c + 1 : c] : halo;
            left = st.R1[c];
        } else {
            up = st.R1[c];
            left = (c > 0) ? st.R1[c > 0 ? c - 1 : 0] : halo;
        }
        if (RIGHT && PREV_RIGHT) diag = (c < ABEA_CPL - 1) ? st.R2[c + 1 < ABEA_CPL ? c + 1 : c] : st.halo_prev;
        else if (!RIGHT && !PREV_RIGHT) diag = (c > 0) ? st.R2[c > 0 ? c - 1 : 0] : st.halo_prev;
        else diag = st.R2[c];
        float lp = abea_emission_t<FAST>(st.x[c], st.kp[c]);
        abea_cell_d<FAST>(lp, up, left, diag, lp_step, lp_stay, lp_skip, Rn[c], fr[c]);
    }
    st.halo_prev = halo;
}

template <bool FAST>
__global__ void __launch_bounds__(128)
abea_fill_kernel(const abea_read_t* __restrict__ reads, int32_t n_reads, const abea_event_t* __restrict__ events,
                 const float4* __restrict__ kparams, const uint32_t* __restrict__ read_flags,
                 uint32_t* __restrict__ trace, abea_result_t* __restrict__ results, abea_consts_t cst,
                 int32_t* __restrict__ queue) {
    const int lane = threadIdx.x & 31;
    const double NEG = abea_neg_inf_d();

    for (;;) {
        int32_t ridx = 0;
        if (lane == 0) ridx = atomicAdd(queue, 1);
        ridx = __shfl_sync(ABEA_FULL, ridx, 0);
        if (ridx >= n_reads) break;
        /* each instantiation takes only the reads validated for its arithmetic */
        if (((read_flags[ridx] & ABEA_READ_FAST) != 0u) != FAST) continue;

        const abea_read_t rd = reads[ridx];
        const int32_t E = rd.n_events, K = rd.n_kmers;
        const int64_t NB = (int64_t)E + (int64_t)K + 2;
        const abea_event_t* __restrict__ ev = events + rd.ev_off;
        const float4* __restrict__ kpr = kparams + rd.kp_off;
        uint32_t* __restrict__ tr = trace + rd.trace_off;
        const double lp_stay = rd.lp_stay, lp_step = rd.lp_step, lp_skip = cst.lp_skip, lp_trim = cst.lp_trim;

        /* band 1 geometry (reference src/align.c:277-279): e0=49,k0=-51 ; band 1 = move_down(band 0) */
        int32_t eb = ABEA_W / 2, kb = -1 - ABEA_W / 2;

        abea_band_state st;
#pragma unroll
        for (int c = 0; c < ABEA_CPL; c++) {
            int o = ABEA_CPL * lane + c;
            st.x[c] = abea_load_event_mean(ev, eb - o, E);
            st.kp[c] = abea_load_kparam(kpr, kb + o, K);
            st.R2[c] = (o == ABEA_W / 2) ? 0.0 : NEG;                                        /* src/align.c:284 */
            st.R1[c] = (o == ABEA_W / 2) ? (double)__double2float_rn(lp_trim) : NEG;          /* src/align.c:290 */
        }
        st.halo_prev = NEG;
        /* register chunk buffers for the elements that enter the window: lane i holds event (ebase+i) and
         * k-mer (kbase+i); the element needed next is fetched with a shuffle, a chunk is refilled every 32 moves */
        int32_t ebase = eb + 1;
        int32_t kbase = kb + ABEA_W;
        float evbuf = abea_load_event_mean(ev, ebase + lane, E);
        float evnext = abea_load_event_mean(ev, ebase + 32 + lane, E);
        float4 kbuf = abea_load_kparam(kpr, kbase + lane, K);
        float4 knext = abea_load_kparam(kpr, kbase + 32 + lane, K);
        bool prev_right = false; /* band 1 was a down move */

        uint32_t tword = (lane == (ABEA_W / 2) / ABEA_CPL) ? (ABEA_FROM_U << (8 + 2 * ((ABEA_W / 2) % ABEA_CPL))) : 0u;
        int32_t eb_keep = (lane == 25) ? (ABEA_W / 2 - 1) : ((lane == 26) ? ABEA_W / 2 : 0);

        double best_s = NEG;
        int32_t best_e = 0x7fffffff;

        for (int64_t b = 2; b < NB; b++) {
            /* --- Suzuki's rule (reference src/align.c:304-322): lane 24 owns ur, fetches ll, votes --- */
            double ll = __shfl_sync(ABEA_FULL, st.R1[0], 0);
            bool my_right = (ll == NEG && st.R1[ABEA_CPL - 1] == NEG) ? ((b & 1) == 1) : (ll < st.R1[ABEA_CPL - 1]);
            const bool right = __any_sync(ABEA_FULL, (lane == ABEA_LANES - 1) && my_right) != 0;

            double Rn[ABEA_CPL];
            uint32_t fr[ABEA_CPL];
            if (right) {
                kb += 1;
                /* k-mer window slides towards lower offsets; the new k-mer kb+99 enters at offset 99 */
                const int src = (kb + ABEA_W - 1) - kbase;
                float4 in, nb;
                in.x = __shfl_sync(ABEA_FULL, kbuf.x, src);
                in.y = __shfl_sync(ABEA_FULL, kbuf.y, src);
                in.z = __shfl_sync(ABEA_FULL, kbuf.z, src);
                in.w = __shfl_sync(ABEA_FULL, kbuf.w, src);
                nb.x = __shfl_down_sync(ABEA_FULL, st.kp[0].x, 1);
                nb.y = __shfl_down_sync(ABEA_FULL, st.kp[0].y, 1);
                nb.z = __shfl_down_sync(ABEA_FULL, st.kp[0].z, 1);
                nb.w = __shfl_down_sync(ABEA_FULL, st.kp[0].w, 1);
#pragma unroll
                for (int c = 0; c < ABEA_CPL - 1; c++) st.kp[c] = st.kp[c + 1];
                st.kp[ABEA_CPL - 1] = (lane == ABEA_LANES - 1) ? in : nb;
                if ((kb + ABEA_W) - kbase == 32) { /* chunk exhausted */
                    kbase += 32;
                    kbuf = knext;
                    knext = abea_load_kparam(kpr, kbase + 32 + lane, K);
                }
                if (prev_right) abea_band_cells<FAST, true, true>(st, lane, lp_step, lp_stay, lp_skip, Rn, fr);
                else abea_band_cells<FAST, true, false>(st, lane, lp_step, lp_stay, lp_skip, Rn, fr);
            } else {
                eb += 1;
                /* event window slides towards higher offsets; the new event eb enters at offset 0 */
                float in = __shfl_sync(ABEA_FULL, evbuf, eb - ebase);
                float nb = __shfl_up_sync(ABEA_FULL, st.x[ABEA_CPL - 1], 1);
#pragma unroll
                for (int c = ABEA_CPL - 1; c > 0; c--) st.x[c] = st.x[c - 1];
                st.x[0] = (lane == 0) ? in : nb;
                if ((eb + 1) - ebase == 32) {
                    ebase += 32;
                    evbuf = evnext;
                    evnext = abea_load_event_mean(ev, ebase + 32 + lane, E);
                }
                if (prev_right) abea_band_cells<FAST, false, true>(st, lane, lp_step, lp_stay, lp_skip, Rn, fr);
                else abea_band_cells<FAST, false, false>(st, lane, lp_step, lp_stay, lp_skip, Rn, fr);
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
                const double trim_s = (double)__double2float_rn(__dmul_rn(lp_trim, (double)(te + 1)));
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
                        double s = (double)__double2float_rn(__dadd_rn(Rn[c], __dmul_rn((double)(E - e), lp_trim)));
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
                st.R2[c] = st.R1[c];
                st.R1[c] = Rn[c];
            }
            prev_right = right;
        }

        /* merge per-lane best end cells: max score, ties to the smaller event (first strict max in event order) */
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            double os = __shfl_xor_sync(ABEA_FULL, best_s, d);
            int32_t oe2 = __shfl_xor_sync(ABEA_FULL, best_e, d);
            if (os > best_s || (os == best_s && oe2 < best_e)) {
                best_s = os;
                best_e = oe2;
            }
        }
        if (lane == 0) {
            results[ridx].end_score = __double2float_rn(best_s);
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
