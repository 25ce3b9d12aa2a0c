/* events_kernels.cuh — event detection on the device.
 *
 *   abea_events_kernel   getevents = detect_events over the whole signal   reference src/events.c:562-582
 *                        (its trim_and_segment_raw result is discarded, :572)
 *                          compute_sum_sumsq          src/events.c:297-307
 *                          compute_tstat (x2)         src/events.c:320-372
 *                          short_long_peak_detector   src/events.c:379-448
 *                          create_events / _event     src/events.c:463-515
 *                        + the pA conversion of event_single               src/f5c.c:692-696
 *
 * The event table feeds the alignment, which is bit-exact integer work, so every float the reference stores is
 * reproduced exactly. A signal is cut into CHUNKS of `cl` samples (1024 by default); five kernels:
 *  1 abea_events_sums_kernel (one warp per read): the cumulative sums are ordered double additions (the square is
 *    a FLOAT product, as in the reference). When every partial sum is exactly representable — all terms are
 *    multiples of one power of two q and n * max|term| < q * 2^53, which the kernel checks per read from the
 *    exponents of its samples and which holds for any realistic signal (picoamperes in [16, 256): 47 of 53 bits for
 *    a million samples) — no addition ever rounds, so the order of the additions cannot matter and the sums are
 *    formed by a warp scan. Otherwise (e.g. samples of 1e-6 next to samples of 200) the lanes stage the terms, lane
 *    0 (sum) and lane 1 (sum of squares) run the two 8-cycle DADD chains in the reference's order and overwrite the
 *    terms with the running sums. Either way all lanes store the running sums coalesced.
 *  2 abea_events_tstat_kernel (one thread per sample): the two windowed t-statistics, with the reference's mix of
 *    float and double operations spelled out with _rn intrinsics.
 *  3 abea_events_spec_kernel (one thread per chunk): the short/long peak detector is a sequential state machine
 *    over the whole signal — 290 cycles per sample for a lone thread, 12.7 ms for the longest signal of a batch when
 *    one thread walks it (measured, profiles/events_stage_r01.txt). But its joint state RENEWS at every boundary
 *    the short detector emits: at that sample the short detector restarts from (no peak, value = current sample)
 *    and has just silenced the long one with constants that depend only on the emitted peak (src/events.c:423-446),
 *    so nothing before the boundary matters any more. Every chunk is therefore walked SPECULATIVELY from the initial
 *    state, all chunks of all reads in parallel, recording what it emits and when.
 *  4 abea_events_stitch_kernel (one thread per read): walks the chunks in order with the TRUE state: each chunk is
 *    re-walked from the true state of its predecessor only until the true walk emits a short-detector boundary
 *    that the speculative walk emitted at the same sample with the same peak — from there on the two walks are
 *    identical, so the speculative emissions after that point and the speculative end state are the true ones.
 *    Typically a few events (tens of samples) per chunk; a chunk that never synchronises is simply walked to its end.
 *  5 abea_events_create_kernel (one warp per read): concatenates per chunk [re-walked emissions] + [speculative
 *    emissions after the synchronisation point] into the boundary list and builds the events in parallel.
 * The detector is written without branches (compares and selects), so the threads of a warp stay in lock-step.
 * Signals shorter than 100 samples (an assert in the reference's trim_raw_by_mad) give 0 events; a signal with no
 * peak (the reference reads peaks[-1]) gives one event over the whole signal.
 */
#pragma once

#define EVT_WARPS 4
#define EVT_CHUNK 128

struct abea_sig_t {
    int64_t raw_off;   /* first sample in d_raw */
    int64_t sum_off;   /* first entry of the read's n+1 cumulative sums in d_sum / d_sumsq */
    int64_t ts_off;    /* first entry of the read's t-statistics in d_ts1 / d_ts2 (a multiple of 4: float4 loads) */
    int64_t cap_off;   /* first slot of the read in d_peaks / d_events_cap (capacity n/2 + 2) */
    int64_t chunk_off; /* first chunk of the read in the chunk arrays */
    int32_t n_samples;
    int32_t cap;       /* n/2 + 2 */
    float offset;      /* pA = (raw + offset) * raw_unit (src/f5c.c:692-696); raw_unit == 0: the samples are pA already */
    float raw_unit;    /* range / digitisation, divided on the host in float */
};

/* One chunk of a read's signal: the unit of the speculative detector. Built by the host. */
struct abea_chunk_t {
    int32_t read;  /* index of the read in the batch */
    int32_t idx;   /* chunk number within the read: samples [idx * cl, min(n, (idx + 1) * cl)) */
};

struct abea_det_param_t { /* src/events.c:52-63 */
    int32_t w1, w2;
    float thr1, thr2, peak_height;
};

/* a / b in double, correctly rounded. For the window lengths of the reference's two parameter sets (3, 6, 7, 14) the
 * quotient is formed from the rounded reciprocal and two FMAs (Markstein's correction) instead of the ~35-instruction
 * division sequence; checked against a / b on 1.2e9 operands spanning 2^-60..2^60 and the fixed-point grid the
 * cumulative sums live on: no mismatch (profiles/events_stage_r01.txt). Any other divisor takes the division. */
__device__ __forceinline__ double evt_div(double a, double b, double rcp, bool fast) {
    if (fast) {
        const double q = __dmul_rn(a, rcp);
        const double r = __fma_rn(-b, q, a);
        return __fma_rn(r, rcp, q);
    }
    return a / b;
}

/* compute_tstat (src/events.c:320-372) for one sample. sum / sumsq point at the read's cumulative sums. */
__device__ __forceinline__ float evt_tstat(const double* __restrict__ sum, const double* __restrict__ sumsq, int32_t n,
                                           int32_t i, int32_t w) {
    if (n < 2 * w || w < 2) return 0.f;
    if (i < w || i > n - w) return 0.f;
    const float wf = (float)w;
    const bool fast = (w == 3) || (w == 6) || (w == 7) || (w == 14);
    const double wd = (double)wf, rcp = 1.0 / wd;
    const double s_i = sum[i], q_i = sumsq[i];
    double sum1 = s_i, sumsq1 = q_i;
    if (i > w) {
        sum1 = __dadd_rn(sum1, -sum[i - w]);
        sumsq1 = __dadd_rn(sumsq1, -sumsq[i - w]);
    }
    const float sum2 = __double2float_rn(__dadd_rn(sum[i + w], -s_i));
    const float sumsq2 = __double2float_rn(__dadd_rn(sumsq[i + w], -q_i));
    const float mean1 = __double2float_rn(evt_div(sum1, wd, rcp, fast));
    const float mean2 = __fdiv_rn(sum2, wf);
    /* sumsq1 / w - mean1 * mean1 + sumsq2 / w - mean2 * mean2: left to right in double, the products and the
     * second quotient are float operations promoted afterwards */
    double cv = __dadd_rn(evt_div(sumsq1, wd, rcp, fast), -(double)__fmul_rn(mean1, mean1));
    cv = __dadd_rn(cv, (double)__fdiv_rn(sumsq2, wf));
    cv = __dadd_rn(cv, -(double)__fmul_rn(mean2, mean2));
    float combined_var = __double2float_rn(cv);
    combined_var = fmaxf(combined_var, 1.17549435e-38f); /* FLT_MIN */
    const float delta_mean = __fsub_rn(mean2, mean1);
    return __fdiv_rn(fabsf(delta_mean), __fsqrt_rn(__fdiv_rn(combined_var, wf)));
}

struct evt_det_t {
    int32_t masked_to;
    int32_t peak_pos; /* -1 = none yet */
    float peak_value;
    int32_t valid;
};

/* joint state of the two detectors */
struct evt_state_t {
    evt_det_t s, l;
};

#define EVT_FMAX 3.402823466e+38f /* FLT_MAX */
#define EVT_SHORT_BIT 0x80000000u /* tag of a boundary emitted by the short detector (in the `time` word) */

__device__ __forceinline__ evt_state_t evt_initial_state() { /* src/events.c:531-551 */
    evt_state_t st;
    st.s.masked_to = 0; st.s.peak_pos = -1; st.s.peak_value = EVT_FMAX; st.s.valid = 0;
    st.l = st.s;
    return st;
}

/* ---- 1. cumulative sums (src/events.c:297-307) ---- */
__global__ void __launch_bounds__(32 * EVT_WARPS)
abea_events_sums_kernel(const abea_sig_t* __restrict__ sigs, int32_t n_reads, const float* __restrict__ raw,
                        double* __restrict__ d_sum, double* __restrict__ d_sumsq) {
    __shared__ __align__(16) double stage[EVT_WARPS][2][EVT_CHUNK];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t r = blockIdx.x * EVT_WARPS + w;
    if (r >= n_reads) return;
    const abea_sig_t sg = sigs[r];
    const int32_t n = sg.n_samples;
    if (n < 100) return;
    const float* x = raw + sg.raw_off;
    double* sum = d_sum + sg.sum_off;
    double* sumsq = d_sumsq + sg.sum_off;
    double* t0 = stage[w][0];
    double* t1 = stage[w][1];
    const bool convert = sg.raw_unit != 0.f;
    if (lane == 0) {
        sum[0] = 0.0;
        sumsq[0] = 0.0;
    }
    /* ---- optimistic pass: warp scan, valid iff no addition can round (checked from the samples' exponents, which
     * are collected on the way); each lane owns 4 consecutive samples of a 128-sample round ---- */
    {
        int32_t e_min = 1000, e_max = -1000;
        double c0 = 0.0, c1 = 0.0; /* carries */
        float nx[EVT_CHUNK / 32];
#pragma unroll
        for (int u = 0; u < EVT_CHUNK / 32; u++) {
            const int32_t i = lane * (EVT_CHUNK / 32) + u;
            nx[u] = (i < n) ? x[i] : 0.f;
        }
        for (int32_t base = 0; base < n; base += EVT_CHUNK) {
            const int32_t i0 = base + lane * (EVT_CHUNK / 32);
            float cur[EVT_CHUNK / 32];
#pragma unroll
            for (int u = 0; u < EVT_CHUNK / 32; u++) cur[u] = nx[u];
#pragma unroll
            for (int u = 0; u < EVT_CHUNK / 32; u++) { /* next round in flight during the scan */
                const int32_t i = i0 + EVT_CHUNK + u;
                nx[u] = (i < n) ? x[i] : 0.f;
            }
            double a[EVT_CHUNK / 32], q[EVT_CHUNK / 32];
            double l0 = 0.0, l1 = 0.0;
#pragma unroll
            for (int u = 0; u < EVT_CHUNK / 32; u++) {
                float sv = cur[u];
                if (i0 + u < n) {
                    if (convert) sv = __fmul_rn(__fadd_rn(sv, sg.offset), sg.raw_unit);
                    const uint32_t bits = __float_as_uint(sv);
                    const int32_t be = (int32_t)((bits >> 23) & 0xffu);
                    if ((bits & 0x7fffffffu) != 0u) {                  /* not +-0 */
                        const int32_t e = (be == 0) ? -149 : be - 127;  /* denormals: as fine as floats get */
                        e_min = e < e_min ? e : e_min;
                        e_max = e > e_max ? e : e_max;
                    }
                } else {
                    sv = 0.f;
                }
                l0 = __dadd_rn(l0, (double)sv);
                l1 = __dadd_rn(l1, (double)__fmul_rn(sv, sv));
                a[u] = l0;
                q[u] = l1;
            }
            double s0 = l0, s1 = l1; /* inclusive scan of the lane totals */
            for (int d = 1; d < 32; d <<= 1) {
                const double u0 = __shfl_up_sync(ABEA_FULL, s0, d), u1 = __shfl_up_sync(ABEA_FULL, s1, d);
                if (lane >= d) {
                    s0 = __dadd_rn(s0, u0);
                    s1 = __dadd_rn(s1, u1);
                }
            }
            const double p0 = __dadd_rn(c0, __dadd_rn(s0, -l0)), p1 = __dadd_rn(c1, __dadd_rn(s1, -l1)); /* exclusive */
#pragma unroll
            for (int u = 0; u < EVT_CHUNK / 32; u++) {
                if (i0 + u < n) {
                    sum[i0 + u + 1] = __dadd_rn(p0, a[u]);
                    sumsq[i0 + u + 1] = __dadd_rn(p1, q[u]);
                }
            }
            c0 = __dadd_rn(c0, __shfl_sync(ABEA_FULL, s0, 31));
            c1 = __dadd_rn(c1, __shfl_sync(ABEA_FULL, s1, 31));
        }
        for (int d = 16; d >= 1; d >>= 1) {
            const int32_t a2 = __shfl_xor_sync(ABEA_FULL, e_min, d), b2 = __shfl_xor_sync(ABEA_FULL, e_max, d);
            e_min = a2 < e_min ? a2 : e_min;
            e_max = b2 > e_max ? b2 : e_max;
        }
        int32_t lg = 0;
        while (((int64_t)1 << lg) < (int64_t)n) lg++;
        /* terms of the sum: multiples of 2^(e_min-23), below 2^(e_max+1); of the sum of squares (float products):
         * multiples of 2^(2 e_min - 23), below 2^(2 e_max + 2); both well inside the double range. Inf / NaN
         * (exponent 128) never pass. If every partial sum is exact, the scan's association gave the reference's values. */
        const bool exact = (e_max < 128) && (e_max >= e_min) && ((e_max + 1 + lg) - (e_min - 23) <= 53) &&
                           ((2 * e_max + 2 + lg) - (2 * e_min - 23) <= 53);
        if (exact || e_max < e_min) return; /* (all samples zero: every sum is 0) */
        __syncwarp();
    }
    /* ---- some addition may round: redo the sums in the reference's order ---- */
    double acc = 0.0; /* lane 0: sum, lane 1: sum of squares */
    float v[EVT_CHUNK / 32];
    for (int u = 0; u < EVT_CHUNK / 32; u++) {
        const int32_t i = lane + 32 * u;
        v[u] = (i < n) ? x[i] : 0.f;
    }
    for (int32_t base = 0; base < n; base += EVT_CHUNK) {
        for (int u = 0; u < EVT_CHUNK / 32; u++) {
            float s = v[u];
            if (convert) s = __fmul_rn(__fadd_rn(s, sg.offset), sg.raw_unit); /* src/f5c.c:695 */
            t0[lane + 32 * u] = (double)s;
            t1[lane + 32 * u] = (double)__fmul_rn(s, s);
        }
        for (int u = 0; u < EVT_CHUNK / 32; u++) { /* next round, in flight while this one is summed */
            const int32_t i = base + EVT_CHUNK + lane + 32 * u;
            if (i < n) v[u] = x[i];
        }
        __syncwarp();
        const int32_t cnt = (n - base < EVT_CHUNK) ? n - base : EVT_CHUNK;
        if (lane < 2) {
            double* t = lane ? t1 : t0;
            int32_t j = 0;
            for (; j + 4 <= cnt; j += 4) {
                const double a0 = t[j], a1 = t[j + 1], a2 = t[j + 2], a3 = t[j + 3];
                const double p0 = __dadd_rn(acc, a0);
                const double p1 = __dadd_rn(p0, a1);
                const double p2 = __dadd_rn(p1, a2);
                acc = __dadd_rn(p2, a3);
                t[j] = p0;
                t[j + 1] = p1;
                t[j + 2] = p2;
                t[j + 3] = acc;
            }
            for (; j < cnt; j++) {
                acc = __dadd_rn(acc, t[j]);
                t[j] = acc;
            }
        }
        __syncwarp();
        for (int u = 0; u < EVT_CHUNK / 32; u++) {
            const int32_t j = lane + 32 * u;
            if (j < cnt) {
                sum[base + 1 + j] = t0[j];
                sumsq[base + 1 + j] = t1[j];
            }
        }
        __syncwarp();
    }
}

/* ---- 2. t-statistics (src/events.c:320-372), both window lengths; one block per chunk ---- */
__global__ void abea_events_tstat_kernel(const abea_sig_t* __restrict__ sigs, const abea_chunk_t* __restrict__ chunks,
                                         int32_t cl, const double* __restrict__ d_sum, const double* __restrict__ d_sumsq,
                                         float* __restrict__ d_ts1, float* __restrict__ d_ts2, abea_det_param_t P) {
    const abea_chunk_t ck = chunks[blockIdx.x];
    const abea_sig_t sg = sigs[ck.read];
    const int32_t n = sg.n_samples;
    if (n < 100) return;
    const double* sum = d_sum + sg.sum_off;
    const double* sumsq = d_sumsq + sg.sum_off;
    const int32_t lo = ck.idx * cl, hi = (lo + cl < n) ? lo + cl : n;
    for (int32_t i = lo + (int32_t)threadIdx.x; i < hi; i += (int32_t)blockDim.x) {
        d_ts1[sg.ts_off + i] = evt_tstat(sum, sumsq, n, i, P.w1);
        d_ts2[sg.ts_off + i] = evt_tstat(sum, sumsq, n, i, P.w2);
    }
}

/* The state machine of src/events.c:379-448 for ONE detector and ONE sample, written without branches. `dom` (short
 * detector only) reports that the long detector must be silenced (:423-431); `emit` that peak `pp` is a boundary
 * (:438-446). */
__device__ __forceinline__ void evt_step(evt_det_t& d, const float v, const int32_t i, const float thr, const int32_t half,
                                         const float peak_height, bool& dom, bool& emit, int32_t& pp) {
    const bool act = d.masked_to < i;
    const bool nopeak = d.peak_pos < 0;
    const bool lower = v < d.peak_value;
    const bool rise = __fsub_rn(v, d.peak_value) > peak_height;
    const bool higher = v > d.peak_value;
    const bool c2 = act && !nopeak;                                   /* CASE 2: in an existing peak (:408) */
    const bool setval = act && (nopeak ? (lower || rise) : higher);
    const bool setpos = act && (nopeak ? (!lower && rise) : higher);
    const float pv = setval ? v : d.peak_value;
    pp = setpos ? i : d.peak_pos;
    const bool over = pv > thr;
    dom = c2 && over;
    const int32_t valid = d.valid | ((c2 && over && (__fsub_rn(pv, v) > peak_height)) ? 1 : 0);
    emit = c2 && (valid != 0) && ((i - pp) > half);
    d.peak_pos = emit ? -1 : pp;
    d.peak_value = emit ? v : pv;
    d.valid = emit ? 0 : valid;
}

/* Both detectors over one sample, in the reference's order (short first). Returns the boundaries emitted (0..2) in
 * pos[] / tag[] (tag = sample index, with EVT_SHORT_BIT for the short detector). */
__device__ __forceinline__ int evt_sample(evt_state_t& st, const float v1, const float v2, const int32_t i,
                                          const abea_det_param_t& P, int32_t* pos, uint32_t* tag) {
    bool dom, emit, dom2;
    int32_t pp;
    int k = 0;
    evt_step(st.s, v1, i, P.thr1, P.w1 / 2, P.peak_height, dom, emit, pp);
    st.l.masked_to = dom ? pp + P.w1 : st.l.masked_to; /* the short detector silences the long one */
    st.l.peak_pos = dom ? -1 : st.l.peak_pos;
    st.l.peak_value = dom ? EVT_FMAX : st.l.peak_value;
    st.l.valid = dom ? 0 : st.l.valid;
    if (emit) {
        pos[k] = pp;
        tag[k] = (uint32_t)i | EVT_SHORT_BIT;
        k++;
    }
    evt_step(st.l, v2, i, P.thr2, P.w2 / 2, P.peak_height, dom2, emit, pp);
    if (emit) {
        pos[k] = pp;
        tag[k] = (uint32_t)i;
        k++;
    }
    return k;
}

/* ---- 3. speculative walk: one thread per chunk, from the initial state ---- */
__global__ void __launch_bounds__(64)
abea_events_spec_kernel(const abea_sig_t* __restrict__ sigs, const abea_chunk_t* __restrict__ chunks, int32_t n_chunks,
                        int32_t cl, int32_t capc, const float* __restrict__ d_ts1, const float* __restrict__ d_ts2,
                        int2* __restrict__ spec, int32_t* __restrict__ spec_cnt, evt_state_t* __restrict__ spec_end,
                        abea_det_param_t P) {
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const abea_chunk_t ck = chunks[c];
    const abea_sig_t sg = sigs[ck.read];
    const int32_t n = sg.n_samples;
    if (n < 100) return;
    const int32_t lo = ck.idx * cl, hi = (lo + cl < n) ? lo + cl : n;
    const float4* ts1 = (const float4*)(d_ts1 + sg.ts_off + lo); /* lo and ts_off are multiples of 4 */
    const float4* ts2 = (const float4*)(d_ts2 + sg.ts_off + lo);
    int2* out = spec + (int64_t)c * capc;
    evt_state_t st = evt_initial_state();
    int32_t cnt = 0;
    const int32_t n4 = (hi - lo + 3) >> 2;
    float4 a = ts1[0], b = ts2[0];
    for (int32_t q = 0; q < n4; q++) {
        const float4 an = (q + 1 < n4) ? ts1[q + 1] : a; /* next group in flight while this one is walked */
        const float4 bn = (q + 1 < n4) ? ts2[q + 1] : b;
        const float va[4] = {a.x, a.y, a.z, a.w};
        const float vb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int32_t i = lo + 4 * q + u;
            if (i < hi) {
                int32_t pos[2];
                uint32_t tag[2];
                const int k = evt_sample(st, va[u], vb[u], i, P, pos, tag);
                for (int e = 0; e < k; e++) {
                    if (cnt < capc) out[cnt] = make_int2(pos[e], (int)tag[e]);
                    cnt++;
                }
            }
        }
        a = an;
        b = bn;
    }
    spec_cnt[c] = cnt;
    spec_end[c] = st;
}

/* ---- 4. stitch: one thread per read (reads listed longest first in order[]), chunks in sequence ---- */
__global__ void __launch_bounds__(32)
abea_events_stitch_kernel(const abea_sig_t* __restrict__ sigs, const int32_t* __restrict__ order, int32_t n_reads,
                          int32_t cl, int32_t capc, const float* __restrict__ d_ts1, const float* __restrict__ d_ts2,
                          const int2* __restrict__ spec, const int32_t* __restrict__ spec_cnt,
                          const evt_state_t* __restrict__ spec_end, int2* __restrict__ fix, int32_t* __restrict__ fix_cnt,
                          int32_t* __restrict__ sync_idx, int32_t* __restrict__ n_events, abea_det_param_t P) {
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_reads) return;
    const int32_t r = order[t];
    const abea_sig_t sg = sigs[r];
    const int32_t n = sg.n_samples;
    if (n < 100) {
        n_events[r] = 0;
        return;
    }
    const int32_t nc = (n + cl - 1) / cl;
    const float* ts1 = d_ts1 + sg.ts_off;
    const float* ts2 = d_ts2 + sg.ts_off;
    const int64_t c0 = sg.chunk_off;
    /* chunk 0 was walked from the true initial state */
    int64_t total = spec_cnt[c0];
    bool overflow = spec_cnt[c0] > capc;
    fix_cnt[c0] = 0;
    sync_idx[c0] = 0;
    evt_state_t st = spec_end[c0];
    for (int32_t c = 1; c < nc; c++) {
        const int64_t cc = c0 + c;
        const int32_t lo = c * cl, hi = (lo + cl < n) ? lo + cl : n;
        const int2* sp = spec + cc * capc;
        const int32_t ns = spec_cnt[cc] < capc ? spec_cnt[cc] : capc;
        int2* fx = fix + cc * capc;
        int32_t nf = 0, cur = 0, sync = -1;
        for (int32_t i = lo; i < hi && sync < 0; i++) {
            int32_t pos[2];
            uint32_t tag[2];
            const int k = evt_sample(st, ts1[i], ts2[i], i, P, pos, tag);
            for (int e = 0; e < k; e++) {
                if (nf < capc) fx[nf] = make_int2(pos[e], (int)tag[e]);
                nf++;
                if ((tag[e] & EVT_SHORT_BIT) && sync < 0) {
                    /* did the speculative walk emit the same short-detector boundary at this very sample? */
                    while (cur < ns && ((uint32_t)sp[cur].y & ~EVT_SHORT_BIT) < (uint32_t)i) cur++;
                    int32_t m = cur;
                    while (m < ns && ((uint32_t)sp[m].y & ~EVT_SHORT_BIT) == (uint32_t)i) {
                        if ((uint32_t)sp[m].y == tag[e] && sp[m].x == pos[e]) sync = m;
                        m++;
                    }
                }
            }
            /* (the long detector cannot emit at the synchronising sample: the short one has just reset it) */
        }
        overflow = overflow || nf > capc || spec_cnt[cc] > capc;
        fix_cnt[cc] = nf;
        if (sync >= 0) {
            sync_idx[cc] = sync + 1;       /* speculative entries from here on are the true ones */
            total += nf + (ns - (sync + 1));
            st = spec_end[cc];
        } else {
            sync_idx[cc] = ns;             /* never synchronised: the re-walk covered the whole chunk */
            total += nf;
        }
    }
    n_events[r] = (overflow || total >= sg.cap - 1) ? -1 : (int32_t)total + 1;
}

/* ---- 5. boundary list + events (src/events.c:463-515): [0, p0), [p0, p1), ..., [p_last, n) ---- */
__global__ void __launch_bounds__(32 * EVT_WARPS)
abea_events_create_kernel(const abea_sig_t* __restrict__ sigs, int32_t n_reads, int32_t cl, int32_t capc,
                          const double* __restrict__ d_sum, const double* __restrict__ d_sumsq,
                          const int2* __restrict__ spec, const int32_t* __restrict__ spec_cnt,
                          const int2* __restrict__ fix, const int32_t* __restrict__ fix_cnt,
                          const int32_t* __restrict__ sync_idx, int32_t* __restrict__ d_peaks,
                          abea_event_t* __restrict__ d_events, const int32_t* __restrict__ n_events) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t r = blockIdx.x * EVT_WARPS + w;
    if (r >= n_reads) return;
    const abea_sig_t sg = sigs[r];
    const int32_t n = sg.n_samples;
    const int32_t n_ev = n_events[r];
    if (n_ev <= 0) return;
    const double* sum = d_sum + sg.sum_off;
    const double* sumsq = d_sumsq + sg.sum_off;
    int32_t* peaks = d_peaks + sg.cap_off;
    abea_event_t* ev = d_events + sg.cap_off;
    const int32_t nc = (n + cl - 1) / cl;
    /* per chunk: [re-walked emissions] ++ [speculative emissions from sync_idx on], chunks in order */
    int32_t off = 0;
    for (int32_t cb = 0; cb < nc; cb += 32) {
        const int32_t c = cb + lane;
        int32_t nf = 0, s0 = 0, ns = 0;
        if (c < nc) {
            const int64_t cc = sg.chunk_off + c;
            nf = fix_cnt[cc];
            s0 = sync_idx[cc];
            ns = spec_cnt[cc];
        }
        const int32_t mine = nf + (ns - s0);
        int32_t incl = mine;
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t up = __shfl_up_sync(ABEA_FULL, incl, d);
            if (lane >= d) incl += up;
        }
        int32_t o = off + incl - mine;
        if (c < nc) {
            const int64_t cc = sg.chunk_off + c;
            const int2* fx = fix + cc * capc;
            const int2* sp = spec + cc * capc;
            for (int32_t j = 0; j < nf; j++) peaks[o++] = fx[j].x;
            for (int32_t j = s0; j < ns; j++) peaks[o++] = sp[j].x;
        }
        off += __shfl_sync(ABEA_FULL, incl, 31);
    }
    __syncwarp();
    for (int32_t e = lane; e < n_ev; e += 32) {
        const int32_t start = (e == 0) ? 0 : peaks[e - 1];
        const int32_t end = (e == n_ev - 1) ? n : peaks[e];
        const float length = (float)(end - start);
        const float mean = __fdiv_rn(__double2float_rn(__dadd_rn(sum[end], -sum[start])), length);
        const float deltasqr = __double2float_rn(__dadd_rn(sumsq[end], -sumsq[start]));
        const float var = __fsub_rn(__fdiv_rn(deltasqr, length), __fmul_rn(mean, mean));
        const float stdv = __fsqrt_rn(fmaxf(var, 0.0f));
        unsigned long long* o = (unsigned long long*)(ev + e); /* 24 bytes: start | length, mean | stdv, 0 */
        o[0] = (unsigned long long)start;
        o[1] = ((unsigned long long)__float_as_uint(mean) << 32) | (unsigned long long)__float_as_uint(length);
        o[2] = (unsigned long long)__float_as_uint(stdv);
    }
}

/* Capacity layout (read i at cap_off, n/2 + 2 slots) -> the caller's compact layout (read i at event_ptr[i]). */
__global__ void abea_events_compact_kernel(const abea_sig_t* __restrict__ sigs, int32_t n_reads,
                                           const abea_event_t* __restrict__ d_events, const int32_t* __restrict__ n_events,
                                           const int64_t* __restrict__ event_ptr, abea_event_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_reads) return;
    const int32_t n = n_events[r];
    const unsigned long long* src = (const unsigned long long*)(d_events + sigs[r].cap_off);
    unsigned long long* dst = (unsigned long long*)(out + event_ptr[r]);
    for (int64_t i = lane; i < 3 * (int64_t)(n > 0 ? n : 0); i += 32) dst[i] = src[i];
}
