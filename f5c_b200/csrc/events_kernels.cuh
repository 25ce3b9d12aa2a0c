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
 * One warp per read. The event table feeds the alignment, which is bit-exact integer work, so every float the
 * reference stores is reproduced exactly:
 *  - the cumulative sums are ordered double additions (the square is a FLOAT product, as in the reference): the
 *    lanes stage the terms, lane 0 (sum) and lane 1 (sum of squares) run the two 8-cycle DADD chains and overwrite
 *    the terms with the running sums, which all lanes then store coalesced;
 *  - the t-statistics are independent per sample: 128 samples are evaluated by the 32 lanes at a time (both window
 *    lengths), with the reference's mix of float and double operations spelled out with _rn intrinsics;
 *  - the peak detector is a sequential state machine over the two t-statistic streams: lane 0 walks the 128 staged
 *    values while the other lanes have already fetched the sums of the next round;
 *  - events are then built in parallel, one lane per event.
 * Signals shorter than 100 samples (an assert in the reference's trim_raw_by_mad) give 0 events; a signal with no
 * peak (the reference reads peaks[-1]) gives one event over the whole signal.
 */
#pragma once

#define EVT_WARPS 4
#define EVT_CHUNK 128

struct abea_sig_t {
    int64_t raw_off;   /* first sample in d_raw */
    int64_t sum_off;   /* first entry of the read's n+1 cumulative sums in d_sum / d_sumsq */
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

/* compute_tstat (src/events.c:320-372) for one sample. sum / sumsq point at the read's cumulative sums. */
__device__ __forceinline__ float evt_tstat(const double* __restrict__ sum, const double* __restrict__ sumsq, int32_t n,
                                           int32_t i, int32_t w) {
    if (n < 2 * w || w < 2) return 0.f;
    if (i < w || i > n - w) return 0.f;
    const float wf = (float)w;
    const double s_i = sum[i], q_i = sumsq[i];
    double sum1 = s_i, sumsq1 = q_i;
    if (i > w) {
        sum1 = __dadd_rn(sum1, -sum[i - w]);
        sumsq1 = __dadd_rn(sumsq1, -sumsq[i - w]);
    }
    const float sum2 = __double2float_rn(__dadd_rn(sum[i + w], -s_i));
    const float sumsq2 = __double2float_rn(__dadd_rn(sumsq[i + w], -q_i));
    const float mean1 = __double2float_rn(sum1 / (double)wf);
    const float mean2 = __fdiv_rn(sum2, wf);
    /* sumsq1 / w - mean1 * mean1 + sumsq2 / w - mean2 * mean2: left to right in double, the products and the
     * second quotient are float operations promoted afterwards */
    double cv = __dadd_rn(sumsq1 / (double)wf, -(double)__fmul_rn(mean1, mean1));
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
abea_events_kernel(const abea_sig_t* __restrict__ sigs, int32_t n_reads, const float* __restrict__ raw,
                   double* __restrict__ d_sum, double* __restrict__ d_sumsq, int32_t* __restrict__ d_peaks,
                   abea_event_t* __restrict__ d_events, int32_t* __restrict__ n_events, abea_det_param_t P) {
    __shared__ __align__(16) double stage[EVT_WARPS][2][EVT_CHUNK];
    __shared__ float tstage[EVT_WARPS][2][EVT_CHUNK];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t r = blockIdx.x * EVT_WARPS + w;
    if (r >= n_reads) return;
    const abea_sig_t sg = sigs[r];
    const int32_t n = sg.n_samples;
    if (n < 100) {
        if (lane == 0) n_events[r] = 0;
        return;
    }
    const float* x = raw + sg.raw_off;
    double* sum = d_sum + sg.sum_off;
    double* sumsq = d_sumsq + sg.sum_off;
    int32_t* peaks = d_peaks + sg.cap_off;
    abea_event_t* ev = d_events + sg.cap_off;
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

    /* ---- t-statistics + peak detection (src/events.c:320-448) ---- */
    float* ts1 = tstage[w][0];
    float* ts2 = tstage[w][1];
    evt_det_t ds, dl; /* short and long detector (lane 0) */
    ds.masked_to = 0; ds.peak_pos = -1; ds.peak_value = 3.402823466e+38f; ds.valid = 0; /* FLT_MAX */
    dl = ds;
    int32_t n_peaks = 0;
    const int32_t half1 = P.w1 / 2, half2 = P.w2 / 2;
    for (int32_t base = 0; base < n; base += EVT_CHUNK) {
        for (int u = 0; u < EVT_CHUNK / 32; u++) {
            const int32_t i = base + lane + 32 * u;
            float a = 0.f, b = 0.f;
            if (i < n) {
                a = evt_tstat(sum, sumsq, n, i, P.w1);
                b = evt_tstat(sum, sumsq, n, i, P.w2);
            }
            ts1[lane + 32 * u] = a;
            ts2[lane + 32 * u] = b;
        }
        __syncwarp();
        if (lane == 0) {
            const int32_t cnt = (n - base < EVT_CHUNK) ? n - base : EVT_CHUNK;
            for (int32_t j = 0; j < cnt; j++) {
                const int32_t i = base + j;
                /* short detector */
                if (ds.masked_to < i) {
                    const float v = ts1[j];
                    if (ds.peak_pos == -1) {
                        if (v < ds.peak_value) {
                            ds.peak_value = v;
                        } else if (__fsub_rn(v, ds.peak_value) > P.peak_height) {
                            ds.peak_value = v;
                            ds.peak_pos = i;
                        }
                    } else {
                        if (v > ds.peak_value) {
                            ds.peak_value = v;
                            ds.peak_pos = i;
                        }
                        if (ds.peak_value > P.thr1) { /* dominate the long detector (:423-431) */
                            dl.masked_to = ds.peak_pos + P.w1;
                            dl.peak_pos = -1;
                            dl.peak_value = 3.402823466e+38f;
                            dl.valid = 0;
                        }
                        if (__fsub_rn(ds.peak_value, v) > P.peak_height && ds.peak_value > P.thr1) ds.valid = 1;
                        if (ds.valid && (i - ds.peak_pos) > half1) {
                            if (n_peaks < sg.cap - 1) peaks[n_peaks] = ds.peak_pos;
                            n_peaks++;
                            ds.peak_pos = -1;
                            ds.peak_value = v;
                            ds.valid = 0;
                        }
                    }
                }
                /* long detector */
                if (dl.masked_to < i) {
                    const float v = ts2[j];
                    if (dl.peak_pos == -1) {
                        if (v < dl.peak_value) {
                            dl.peak_value = v;
                        } else if (__fsub_rn(v, dl.peak_value) > P.peak_height) {
                            dl.peak_value = v;
                            dl.peak_pos = i;
                        }
                    } else {
                        if (v > dl.peak_value) {
                            dl.peak_value = v;
                            dl.peak_pos = i;
                        }
                        if (__fsub_rn(dl.peak_value, v) > P.peak_height && dl.peak_value > P.thr2) dl.valid = 1;
                        if (dl.valid && (i - dl.peak_pos) > half2) {
                            if (n_peaks < sg.cap - 1) peaks[n_peaks] = dl.peak_pos;
                            n_peaks++;
                            dl.peak_pos = -1;
                            dl.peak_value = v;
                            dl.valid = 0;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    n_peaks = __shfl_sync(ABEA_FULL, n_peaks, 0);
    if (n_peaks >= sg.cap - 1) { /* more boundaries than any real signal has: refuse rather than truncate */
        if (lane == 0) n_events[r] = -1;
        return;
    }

    /* ---- events (src/events.c:463-515): [0, p0), [p0, p1), ..., [p_last, n) ---- */
    const int32_t n_ev = n_peaks + 1;
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
    if (lane == 0) n_events[r] = n_ev;
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
    for (int64_t i = lane; i < 3 * (int64_t)n; i += 32) dst[i] = src[i];
}
