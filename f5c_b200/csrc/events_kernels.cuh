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
 * reproduced exactly. Three kernels:
 *  - abea_events_sums_kernel (one warp per read): the cumulative sums are ordered double additions (the square is a
 *    FLOAT product, as in the reference): the lanes stage the terms, lane 0 (sum) and lane 1 (sum of squares) run
 *    the two 8-cycle DADD chains and overwrite the terms with the running sums, which all lanes then store
 *    coalesced; then the two t-statistics, independent per sample, are evaluated by the 32 lanes with the
 *    reference's mix of float and double operations spelled out with _rn intrinsics, and written to d_ts;
 *  - abea_events_detect_kernel (one THREAD per read, reads sorted by length so that a warp's 32 reads end
 *    together): the peak detector is a sequential state machine over the two t-statistic streams; written without
 *    branches it runs 32 reads in lock-step, and a warp with only one lane at work — what the first version did —
 *    is avoided (that version issued 75 instructions per sample at 1/32 lane occupancy: 14.8 ms per 81 M samples);
 *  - abea_events_create_kernel (one warp per read): events are built in parallel, one lane per event.
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
    int32_t n_samples;
    int32_t cap;       /* n/2 + 2 */
    float offset;      /* pA = (raw + offset) * raw_unit (src/f5c.c:692-696); raw_unit == 0: the samples are pA already */
    float raw_unit;    /* range / digitisation, divided on the host in float */
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

__global__ void __launch_bounds__(32 * EVT_WARPS)
abea_events_sums_kernel(const abea_sig_t* __restrict__ sigs, int32_t n_reads, const float* __restrict__ raw,
                        double* __restrict__ d_sum, double* __restrict__ d_sumsq, float* __restrict__ d_ts1,
                        float* __restrict__ d_ts2, abea_det_param_t P) {
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

    /* ---- cumulative sums (src/events.c:297-307) ---- */
    if (lane == 0) {
        sum[0] = 0.0;
        sumsq[0] = 0.0;
    }
    {
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

    /* ---- t-statistics (src/events.c:320-372), both window lengths ---- */
    float* ts1 = d_ts1 + sg.ts_off;
    float* ts2 = d_ts2 + sg.ts_off;
    for (int32_t i = lane; i < n; i += 32) {
        ts1[i] = evt_tstat(sum, sumsq, n, i, P.w1);
        ts2[i] = evt_tstat(sum, sumsq, n, i, P.w2);
    }
}

/* The state machine of src/events.c:379-448 for ONE detector and ONE sample, written without branches: compares and
 * selects only, so that 32 reads advance in lock-step. `dom` (short detector only) reports that the long detector
 * must be silenced (:423-431); `emit` that peak `pp` is a boundary (:438-446). */
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

/* One thread per read; order[] lists the reads longest first, so the 32 reads of a warp have similar lengths. */
__global__ void __launch_bounds__(32)
abea_events_detect_kernel(const abea_sig_t* __restrict__ sigs, const int32_t* __restrict__ order, int32_t n_reads,
                          const float* __restrict__ d_ts1, const float* __restrict__ d_ts2,
                          int32_t* __restrict__ d_peaks, int32_t* __restrict__ n_events, abea_det_param_t P) {
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_reads) return;
    const int32_t r = order[t];
    const abea_sig_t sg = sigs[r];
    const int32_t n = sg.n_samples;
    if (n < 100) {
        n_events[r] = 0;
        return;
    }
    const float4* ts1 = (const float4*)(d_ts1 + sg.ts_off);
    const float4* ts2 = (const float4*)(d_ts2 + sg.ts_off);
    int32_t* peaks = d_peaks + sg.cap_off;
    const float FMAX = 3.402823466e+38f;
    evt_det_t ds, dl; /* short and long detector */
    ds.masked_to = 0; ds.peak_pos = -1; ds.peak_value = FMAX; ds.valid = 0;
    dl = ds;
    int32_t n_peaks = 0;
    const int32_t half1 = P.w1 / 2, half2 = P.w2 / 2;
    const int32_t cap1 = sg.cap - 1;
    const int32_t n4 = (n + 3) >> 2; /* the arrays are padded to a multiple of 4 */
    float4 a = ts1[0], b = ts2[0];
    for (int32_t q = 0; q < n4; q++) {
        const float4 an = (q + 1 < n4) ? ts1[q + 1] : a; /* next group in flight while this one is walked */
        const float4 bn = (q + 1 < n4) ? ts2[q + 1] : b;
        const float va[4] = {a.x, a.y, a.z, a.w};
        const float vb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int32_t i = 4 * q + u;
            if (i < n) {
                bool dom, emit;
                int32_t pp;
                evt_step(ds, va[u], i, P.thr1, half1, P.peak_height, dom, emit, pp);
                dl.masked_to = dom ? pp + P.w1 : dl.masked_to; /* the short detector silences the long one */
                dl.peak_pos = dom ? -1 : dl.peak_pos;
                dl.peak_value = dom ? FMAX : dl.peak_value;
                dl.valid = dom ? 0 : dl.valid;
                if (emit && n_peaks < cap1) peaks[n_peaks] = pp;
                n_peaks += emit ? 1 : 0;
                bool dom2;
                evt_step(dl, vb[u], i, P.thr2, half2, P.peak_height, dom2, emit, pp);
                if (emit && n_peaks < cap1) peaks[n_peaks] = pp;
                n_peaks += emit ? 1 : 0;
            }
        }
        a = an;
        b = bn;
    }
    /* more boundaries than any real signal has: refuse rather than truncate */
    n_events[r] = (n_peaks >= cap1) ? -1 : n_peaks + 1;
}

__global__ void __launch_bounds__(32 * EVT_WARPS)
abea_events_create_kernel(const abea_sig_t* __restrict__ sigs, int32_t n_reads, const double* __restrict__ d_sum,
                          const double* __restrict__ d_sumsq, const int32_t* __restrict__ d_peaks,
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
    const int32_t* peaks = d_peaks + sg.cap_off;
    abea_event_t* ev = d_events + sg.cap_off;
    /* ---- events (src/events.c:463-515): [0, p0), [p0, p1), ..., [p_last, n) ---- */
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
