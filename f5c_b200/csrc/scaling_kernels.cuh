/* scaling_kernels.cuh — the two stages either side of ABEA, on the device.
 *
 *   abea_mom_kernel      estimate_scalings_using_mom            reference src/align.c:58-106 (call site src/f5c.c:709-711)
 *                        + the RNA event reversal that follows  src/f5c.c:713-721
 *   abea_scaling_kernel  scaling_single = postalign             src/f5c.c:736-807, src/align.c:561-660
 *                        + recalibrate_model + read flags       src/align.c:665-773
 *
 * Both are per-read reductions whose results feed bit-exact integer work (the alignment itself, and the k-mer -> event
 * map), so they reproduce the reference's floating-point results exactly, not approximately. The reference adds
 * its terms one after another into double accumulators; double addition is not associative, so a tree reduction
 * would differ in the last bits. Here the TERMS are produced in parallel by the 32 lanes of the warp that owns the
 * read (loads, k-mer ranks, model gathers, the double division and products), staged in shared memory in list
 * order, and the ADDITIONS are done in that order by one lane per accumulator: an 8-cycle DADD chain per term
 * (profiles/microbench_r01.txt), which for the longest read of a batch is still two orders of magnitude below its
 * band fill. Built with -fmad=false like everything else, so a*b+c is never contracted; `/` and sqrt on doubles are
 * IEEE-correct on the device. log(var) is left to the host (glibc), see abea_scaling_download.
 *
 * One warp per read, reads in the caller's order. Compiles unchanged for the CPU SIMT emulator (tests/simt).
 */
#pragma once

#define SCL_WARPS 4     /* warps (reads in flight) per CTA */
#define SCL_CHUNK 128   /* terms staged per round of the method-of-moments sums (4 per lane) */

/* Per-read descriptor of the scaling stages, in the CALLER's order (every read of the batch has one). */
struct abea_sread_t {
    int64_t seq_off;   /* first base in d_seq */
    int64_t ev_off;    /* first event mean in d_means */
    int64_t map_off;   /* first entry of the read's base_to_event_map in d_maps (prefix sum of max(K, 0)) */
    int64_t pair_off;  /* first pair slot in d_pairs (canonical capacity layout) */
    int32_t n_events;  /* E */
    int32_t read_len;  /* L */
    int32_t sched;     /* position in the ABEA schedule (d_reads), -1 if the read was not scheduled */
    int32_t usable;    /* good read with E >= 1 and L >= k: the reference would have called event_single on it */
};

__device__ __forceinline__ uint32_t scl_kmer_rank(const uint8_t* __restrict__ s, uint32_t k) {
    uint32_t r = 0;
    for (uint32_t j = 0; j < k; j++) r = (r << 2) | abea_base_rank(s[j]);
    return r;
}

/* acc + t[0] + t[1] + ... in exactly that order: the terms are fetched four at a time so that only the additions are
 * on the dependent chain (8 cycles each); no padding terms are ever added. */
__device__ __forceinline__ double scl_chain_d(const double* t, int32_t cnt, double acc) {
    int32_t j = 0;
    for (; j + 8 <= cnt; j += 8) {
        const double v0 = t[j], v1 = t[j + 1], v2 = t[j + 2], v3 = t[j + 3];
        const double v4 = t[j + 4], v5 = t[j + 5], v6 = t[j + 6], v7 = t[j + 7];
        acc = __dadd_rn(acc, v0);
        acc = __dadd_rn(acc, v1);
        acc = __dadd_rn(acc, v2);
        acc = __dadd_rn(acc, v3);
        acc = __dadd_rn(acc, v4);
        acc = __dadd_rn(acc, v5);
        acc = __dadd_rn(acc, v6);
        acc = __dadd_rn(acc, v7);
    }
    for (; j < cnt; j++) acc = __dadd_rn(acc, t[j]);
    return acc;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* N2. shift = mean(event) - mean(level); scale = mean((event - shift)^2) / mean(level^2), all sums sequential in
 * double (src/align.c:68-95). Three staged passes: events, k-mers, events again (the second needs shift). The next
 * round's values are loaded into registers before the current round is summed, so the loads hide behind the chain. */
__global__ void __launch_bounds__(32 * SCL_WARPS)
abea_mom_kernel(const abea_sread_t* __restrict__ sreads, int32_t n_reads, const uint8_t* __restrict__ seq,
                float* __restrict__ means, const abea_model_t* __restrict__ model, uint32_t kmer_size,
                abea_scalings_t* __restrict__ scalings, abea_read_t* __restrict__ reads, int32_t reverse_events) {
    __shared__ __align__(16) double stage[SCL_WARPS][2][SCL_CHUNK];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t r = blockIdx.x * SCL_WARPS + w;
    if (r >= n_reads) return;
    const abea_sread_t rd = sreads[r];
    if (!rd.usable) return;
    double* t0 = stage[w][0];
    double* t1 = stage[w][1];
    const int32_t E = rd.n_events, K = rd.read_len - (int32_t)kmer_size + 1;
    float* ev = means + rd.ev_off;
    const uint8_t* s = seq + rd.seq_off;

    /* The lanes turn the loaded floats into finished double terms (conversion on the XU pipe, products on the FP64
     * pipe, all 32 lanes at once); lane 0 — and lane 1 for the second sum of the k-mer pass — only adds. */
    double acc0 = 0.0; /* lane 0: the sum of the pass; lane 1: the second sum of the k-mer pass */
    double shift = 0.0, kmer_level_sq_sum = 0.0;
    for (int pass = 0; pass < 3; pass++) {
        const int32_t n = (pass == 1) ? K : E;
        float v[SCL_CHUNK / 32];
        for (int u = 0; u < SCL_CHUNK / 32; u++) {
            const int32_t i = lane + 32 * u;
            v[u] = 0.f;
            if (i < n) v[u] = (pass == 1) ? model[scl_kmer_rank(s + i, kmer_size)].level_mean : ev[i];
        }
        acc0 = 0.0;
        for (int32_t base = 0; base < n; base += SCL_CHUNK) {
            for (int u = 0; u < SCL_CHUNK / 32; u++) {
                const double x = (double)v[u];
                if (pass == 0) {
                    t0[lane + 32 * u] = x;                               /* event_level_sum += mean (:69-71) */
                } else if (pass == 1) {
                    t0[lane + 32 * u] = x;                               /* kmer_level_sum += l (:79) */
                    t1[lane + 32 * u] = __dmul_rn(x, x);                 /* kmer_level_sq_sum += l * l (:80) */
                } else {
                    const double d = __dadd_rn(x, -shift);
                    t0[lane + 32 * u] = __dmul_rn(d, d);                 /* (mean - shift) * (mean - shift) (:88-91) */
                }
            }
            for (int u = 0; u < SCL_CHUNK / 32; u++) { /* next round, in flight while this one is summed */
                const int32_t i = base + SCL_CHUNK + lane + 32 * u;
                if (i < n) v[u] = (pass == 1) ? model[scl_kmer_rank(s + i, kmer_size)].level_mean : ev[i];
            }
            __syncwarp();
            const int32_t cnt = (n - base < SCL_CHUNK) ? n - base : SCL_CHUNK;
            if (lane < ((pass == 1) ? 2 : 1)) acc0 = scl_chain_d(lane ? t1 : t0, cnt, acc0);
            __syncwarp();
        }
        if (pass == 0) {
            shift = acc0; /* event_level_sum, held by lane 0 */
        } else if (pass == 1) {
            const double kmer_level_sum = __shfl_sync(ABEA_FULL, acc0, 0);
            kmer_level_sq_sum = __shfl_sync(ABEA_FULL, acc0, 1);
            const double event_level_sum = __shfl_sync(ABEA_FULL, shift, 0);
            /* event_level_sum / et.n - kmer_level_sum / n_kmers (:84): et.n is size_t, n_kmers int32 */
            shift = event_level_sum / (double)(size_t)E - kmer_level_sum / (double)K;
        }
    }
    if (lane == 0) {
        const double scale = (acc0 / (double)(size_t)E) / (kmer_level_sq_sum / (double)K);                 /* :93 */
        abea_scalings_t out;
        out.shift = (float)shift;
        out.scale = (float)scale;
        out.var = 0.f;
        out.log_var = 0.f;
        scalings[r] = out;
        if (rd.sched >= 0) { /* the alignment that follows reads its scalings from the schedule's descriptor */
            reads[rd.sched].scale = out.scale;
            reads[rd.sched].shift = out.shift;
        }
    }
    if (reverse_events) { /* RNA: events become 3'->5' AFTER the estimate (src/f5c.c:713-721); only their means live here */
        __syncwarp();
        for (int32_t i = lane; i < E / 2; i += 32) {
            const float a = ev[i], b = ev[E - 1 - i];
            ev[i] = b;
            ev[E - 1 - i] = a;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* N1. One warp per read of the batch.
 *
 * postalign's walk over the pair list (src/align.c:585-598) relies on nothing but list order; here it is done in
 * parallel using what a traceback path guarantees: ref_pos never decreases, the pairs of one k-mer are consecutive,
 * and inside such a run every pair but the first was reached by an "up" move (a new event). So the first pair of a
 * run whose event differs from its predecessor's gives `start`, else the second one does; the run's last pair gives
 * `stop`; a run that is a single pair reached by a skip (same event as before) leaves the k-mer without events.
 *
 * The alignment list (src/align.c:611-655) is never materialised: recalibrate_model only reads its 'M' entries —
 * the first event of every k-mer that has events and whose rank differs from the previous k-mer that has events
 * (:640) — and those are found 32 k-mers at a time with a ballot. Their five normal-equation terms (:703-724) are
 * computed by the lanes, compacted in k-mer order into shared memory, and added by lanes 0..4, one accumulator each.
 */
#define SCL_ROUND 4 /* 32-k-mer chunks per round of the recalibration sums */
struct scl_terms_t {
    double t[5][32 * SCL_ROUND];
};

/* The 'M' rows among 32 consecutive k-mers, given this lane's map entry and k-mer rank (-1 / {-1,-1} past the end):
 * returns whether this lane's k-mer is a row, its position among the chunk's rows and the row count, and carries the
 * rank of the last k-mer with events forward. */
__device__ __forceinline__ bool scl_m_rows(const abea_index_pair_t m, const int32_t rank, const int lane,
                                           int32_t& carry_rank, int32_t& span, int& pos, int& cnt) {
    const bool has = m.start != -1;
    span = has ? (m.stop - m.start + 1) : 0;
    const uint32_t hm = __ballot_sync(ABEA_FULL, has);
    const uint32_t below = hm & ((1u << lane) - 1u);
    const int src = below ? (31 - __clz((int)below)) : 0;
    const int32_t prev_in_chunk = __shfl_sync(ABEA_FULL, rank, src);
    const int32_t prev_rank = below ? prev_in_chunk : carry_rank;
    const bool isM = has && (rank != prev_rank);
    const uint32_t mm = __ballot_sync(ABEA_FULL, isM);
    pos = __popc(mm & ((1u << lane) - 1u));
    cnt = __popc(mm);
    const int last = hm ? (31 - __clz((int)hm)) : 0;
    const int32_t last_rank = __shfl_sync(ABEA_FULL, rank, last);
    if (hm) carry_rank = last_rank;
    return isM;
}

/* One round = SCL_ROUND chunks of 32 k-mers: all the loads first (map entries, sequence bytes), so that their
 * latencies overlap, then the ballots. Fills rank / start / pos / isM per chunk; returns the number of rows and adds
 * the events covered to n_ea. */
__device__ __forceinline__ int scl_round_rows(const abea_index_pair_t* map, const uint8_t* __restrict__ s,
                                              uint32_t kmer_size, int32_t base, int32_t K, int lane, int32_t& carry,
                                              int32_t* rank, int32_t* start, int* pos, bool* isM, int32_t& n_ea) {
    abea_index_pair_t m[SCL_ROUND];
#pragma unroll
    for (int u = 0; u < SCL_ROUND; u++) {
        const int32_t ki = base + 32 * u + lane;
        m[u].start = -1;
        m[u].stop = -1;
        rank[u] = -1;
        if (ki < K) {
            m[u] = map[ki];
            rank[u] = (int32_t)scl_kmer_rank(s + ki, kmer_size);
        }
    }
    int total = 0;
#pragma unroll
    for (int u = 0; u < SCL_ROUND; u++) {
        int32_t span;
        int cnt;
        isM[u] = scl_m_rows(m[u], rank[u], lane, carry, span, pos[u], cnt);
        start[u] = m[u].start;
        pos[u] += total;
        total += cnt;
        n_ea += span;
    }
    return total;
}

__global__ void __launch_bounds__(32 * SCL_WARPS)
abea_scaling_kernel(const abea_sread_t* __restrict__ sreads, int32_t n_reads, const uint8_t* __restrict__ seq,
                    const float* __restrict__ means, const abea_model_t* __restrict__ model,
                    uint32_t kmer_size, const abea_pair_t* __restrict__ pairs, const int32_t* __restrict__ n_pairs,
                    const abea_scalings_t* __restrict__ scalings_in, abea_index_pair_t* __restrict__ maps,
                    abea_scaling_result_t* __restrict__ results, int32_t min_num_events_to_rescale) {
    __shared__ scl_terms_t terms[SCL_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t r = blockIdx.x * SCL_WARPS + w;
    if (r >= n_reads) return;
    const abea_sread_t rd = sreads[r];
    const int32_t np = n_pairs[r];
    abea_scaling_result_t out;
    out.scalings = scalings_in[r];
    out.var_d = 0.0;
    out.events_per_base = 0.0;
    out.n_event_alignment = 0;
    out.num_m_state = 0;
    out.flags = 0;
    out.calibrated = 0;
    const int32_t K = rd.read_len - (int32_t)kmer_size + 1;
    abea_index_pair_t* map = maps + rd.map_off;
    if (np <= 0) { /* could not align (src/f5c.c:787-793): the reference allocates no map; here the read's region of
                    * d_maps is set to "no events" so that no consumer ever sees an earlier batch's entries */
        out.flags = ABEA_FAILED_ALIGNMENT;
        if (lane == 0) results[r] = out;
        for (int32_t ki = lane; ki < K; ki += 32) {
            abea_index_pair_t e;
            e.start = -1;
            e.stop = -1;
            map[ki] = e;
        }
        return;
    }
    const abea_pair_t* p = pairs + rd.pair_off;
    const float* ev = means + rd.ev_off;
    const uint8_t* s = seq + rd.seq_off;
    scl_terms_t& T = terms[w];

    /* ---- postalign, part 1: the k-mer -> event range map (src/align.c:573-598) ---- */
    for (int32_t ki = lane; ki < K; ki += 32) {
        abea_index_pair_t e;
        e.start = -1;
        e.stop = -1;
        map[ki] = e;
    }
    __syncwarp();
    int32_t max_event = 0, min_event = 0x7fffffff;
    for (int32_t base = 0; base < np; base += 32 * SCL_ROUND) {
        abea_pair_t c[SCL_ROUND], pm1[SCL_ROUND], pm2[SCL_ROUND], pp1[SCL_ROUND];
#pragma unroll
        for (int u = 0; u < SCL_ROUND; u++) { /* all the loads of the round first */
            const int32_t i = base + 32 * u + lane;
            c[u].ref_pos = -1; c[u].read_pos = -1; /* prev_event_idx starts at -1 (:583) */
            pm1[u] = c[u];
            pm2[u] = c[u];
            pp1[u] = c[u];
            if (i < np) {
                c[u] = p[i];
                if (i >= 1) pm1[u] = p[i - 1];
                if (i >= 2) pm2[u] = p[i - 2];
                if (i + 1 < np) pp1[u] = p[i + 1];
            }
        }
#pragma unroll
        for (int u = 0; u < SCL_ROUND; u++) {
            const int32_t i = base + 32 * u + lane;
            if (i < np) {
                const bool q = c[u].read_pos != pm1[u].read_pos;                 /* a new event (:591) */
                const bool first = (i == 0) || (pm1[u].ref_pos != c[u].ref_pos); /* first pair of its k-mer's run */
                const bool last = (i + 1 == np) || (pp1[u].ref_pos != c[u].ref_pos);
                if (q) {
                    /* second of its run, after a first pair that was reached by a skip (same event as its predecessor) */
                    const bool after_skip = !first && (i >= 2) && (pm2[u].ref_pos != pm1[u].ref_pos) &&
                                            (pm1[u].read_pos == pm2[u].read_pos);
                    if (first || after_skip) map[c[u].ref_pos].start = c[u].read_pos;
                    if (last) map[c[u].ref_pos].stop = c[u].read_pos;
                }
                max_event = max_event > c[u].read_pos ? max_event : c[u].read_pos;
                min_event = min_event < c[u].read_pos ? min_event : c[u].read_pos;
            }
        }
    }
    for (int d = 16; d >= 1; d >>= 1) {
        const int32_t a = __shfl_xor_sync(ABEA_FULL, max_event, d), b = __shfl_xor_sync(ABEA_FULL, min_event, d);
        max_event = max_event > a ? max_event : a;
        min_event = min_event < b ? min_event : b;
    }
    out.events_per_base = (double)(max_event - min_event) / (double)K; /* :604 */
    __syncwarp();

    /* ---- postalign, part 2 + recalibrate_model, first pass: the 'M' rows and the normal equations ---- */
    double acc = 0.0; /* lane c < 5: A00, A01, A11, b0, b1 */
    int32_t n_ea = 0, n_m = 0, carry = -1;
    for (int32_t base = 0; base < K; base += 32 * SCL_ROUND) {
        int32_t rank[SCL_ROUND], start[SCL_ROUND];
        int pos[SCL_ROUND];
        bool isM[SCL_ROUND];
        const int total = scl_round_rows(map, s, kmer_size, base, K, lane, carry, rank, start, pos, isM, n_ea);
        n_m += total;
        abea_model_t gm[SCL_ROUND];
        float ge[SCL_ROUND];
#pragma unroll
        for (int u = 0; u < SCL_ROUND; u++) { /* the gathers of the whole round in flight together */
            gm[u].level_mean = 0.f;
            gm[u].level_stdv = 1.f;
            gm[u].level_log_stdv = 0.f;
            ge[u] = 0.f;
            if (isM[u]) {
                gm[u] = model[rank[u]];
                ge[u] = ev[start[u]];                         /* :709 */
            }
        }
#pragma unroll
        for (int u = 0; u < SCL_ROUND; u++) {
            if (isM[u]) {
                const double e = (double)ge[u];
                const double mu = (double)gm[u].level_mean;
                const double sd = (double)gm[u].level_stdv;
                const double inv_var = 1. / __dmul_rn(sd, sd);     /* :713 */
                const int q = pos[u];
                T.t[0][q] = inv_var;                               /* A00 += inv_var */
                T.t[1][q] = __dmul_rn(mu, inv_var);                /* A01 += mu * inv_var */
                T.t[2][q] = __dmul_rn(__dmul_rn(mu, mu), inv_var); /* A11 += mu * mu * inv_var */
                T.t[3][q] = __dmul_rn(e, inv_var);                 /* b0 += e * inv_var */
                T.t[4][q] = __dmul_rn(__dmul_rn(mu, e), inv_var);  /* b1 += mu * e * inv_var */
            }
        }
        __syncwarp();
        if (lane < 5) acc = scl_chain_d(T.t[lane], total, acc);
        __syncwarp();
    }
    for (int d = 16; d >= 1; d >>= 1) n_ea += __shfl_xor_sync(ABEA_FULL, n_ea, d);
    out.n_event_alignment = n_ea;
    out.num_m_state = n_m;

    if (n_m >= min_num_events_to_rescale) { /* :696 */
        const double A00 = __shfl_sync(ABEA_FULL, acc, 0), A01 = __shfl_sync(ABEA_FULL, acc, 1);
        const double A11 = __shfl_sync(ABEA_FULL, acc, 2), b0 = __shfl_sync(ABEA_FULL, acc, 3);
        const double b1 = __shfl_sync(ABEA_FULL, acc, 4);
        const double A10 = A01;
        const double div = __dadd_rn(__dmul_rn(A00, A11), -__dmul_rn(A01, A10));                       /* :729 */
        const double shift = -(__dadd_rn(__dmul_rn(A01, b1), -__dmul_rn(A11, b0))) / div;              /* :730 */
        const double scale = __dadd_rn(__dmul_rn(A00, b1), -__dmul_rn(A10, b0)) / div;                 /* :731 */
        /* second pass: var = sqrt(sum(yi^2 / stdv^2) / num_M) over the same rows (:738-751) */
        double vacc = 0.0;
        carry = -1;
        for (int32_t base = 0; base < K; base += 32 * SCL_ROUND) {
            int32_t rank[SCL_ROUND], start[SCL_ROUND], unused = 0;
            int pos[SCL_ROUND];
            bool isM[SCL_ROUND];
            const int total = scl_round_rows(map, s, kmer_size, base, K, lane, carry, rank, start, pos, isM, unused);
            abea_model_t gm[SCL_ROUND];
            float ge[SCL_ROUND];
#pragma unroll
            for (int u = 0; u < SCL_ROUND; u++) {
                gm[u].level_mean = 0.f;
                gm[u].level_stdv = 1.f;
                gm[u].level_log_stdv = 0.f;
                ge[u] = 0.f;
                if (isM[u]) {
                    gm[u] = model[rank[u]];
                    ge[u] = ev[start[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < SCL_ROUND; u++) {
                if (isM[u]) {
                    const double e = (double)ge[u];
                    const double mu = (double)gm[u].level_mean;
                    const double sd = (double)gm[u].level_stdv;
                    const double yi = __dadd_rn(__dadd_rn(e, -shift), -__dmul_rn(scale, mu)); /* raw - shift - scale * level */
                    T.t[0][pos[u]] = __dmul_rn(yi, yi) / __dmul_rn(sd, sd);
                }
            }
            __syncwarp();
            if (lane == 0) vacc = scl_chain_d(T.t[0], total, vacc);
            __syncwarp();
        }
        if (lane == 0) {
            double var = vacc / (double)n_m; /* :749 */
            var = sqrt(var);                 /* :750 */
            out.scalings.shift = (float)shift;
            out.scalings.scale = (float)scale;
            out.scalings.var = (float)var;
            out.scalings.log_var = 0.f;      /* (float)log(var): glibc's log on the host, abea_scaling_download */
            out.var_d = var;
            out.calibrated = 1;
        }
    }
    if (lane == 0) {
        /* QC of scaling_single (src/f5c.c:776-803) */
        if (!out.calibrated || (double)out.scalings.var > ABEA_MIN_CALIBRATION_VAR) out.flags |= ABEA_FAILED_CALIBRATION;
        else if (out.events_per_base > ABEA_MAX_EVENTS_PER_BASE) out.flags |= ABEA_FAILED_QUALITY_CHK;
        results[r] = out;
    }
}
