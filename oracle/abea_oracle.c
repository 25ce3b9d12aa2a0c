/* abea_oracle.c — TEST INFRASTRUCTURE: CPU restatement of f5c's adaptive banded event alignment.
 *
 * Parity status: PINNED. This restatement is checked (tests/test_oracle.py) against
 *   (1) the reference's golden vectors test/ecoli_2kb_region/single_read/adaptive.exp and
 *       test/ecoli_2kb_region/adaptive.exp (n_aligned_events exact, sum_emission to print precision),
 *   (2) the unmodified reference align() compiled in place (oracle/_ref/libf5c_ref.so): identical pair
 *       lists on real reads and on seeded synthetic R9 / R10 / RNA004 batches,
 *   (3) committed fixtures under tests/golden/ generated from (2) by tests/golden/make_golden.py.
 *
 * It is written from the algorithm, not from the reference's text; every function cites the reference
 * lines whose behaviour it must reproduce (paths relative to /root/reference/). Arithmetic notes that
 * matter for bit-exactness: emissions are float; the three transition sums are evaluated in double and
 * rounded once to float (lp_* are double in the reference, src/align.c:207-216, 382-384); build with
 * -ffp-contract=off so no FMA is formed.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library. The product path (f5c_b200/) never does.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/abea_types.h"

#define W ABEA_BANDWIDTH

enum { FROM_D = 0, FROM_U = 1, FROM_L = 2 }; /* src/align.c:194-196 */

typedef struct {
    double sum_emission;   /* double sum of float emissions in traceback order (src/align.c:476) */
    float end_score;       /* best last-column score incl. trailing trim (src/align.c:438-443) */
    int32_t n_aligned;     /* pairs emitted before QC (src/align.c:480) */
    int32_t max_gap;       /* longest run of FROM_L (src/align.c:495-497) */
    int32_t end_event;     /* event index the traceback starts from */
    int32_t spanned;       /* first.ref_pos==0 && last.ref_pos==K-1 (src/align.c:529-530) */
    int64_t n_bands;
    int64_t n_fills;       /* cells filled by the inner loop (src/align.c:408) */
} abea_oracle_stats_t;

/* src/align.c:19-32: A,C,G,T -> 0..3, anything else -> 0 (the reference also prints a WARNING) */
static inline uint32_t base_rank(char b) {
    switch (b) {
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return 0;
    }
}

/* src/align.c:36-47: first base is the most significant 2-bit digit */
static inline uint32_t kmer_rank(const char* s, uint32_t k) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < k; i++) r = (r << 2) | base_rank(s[i]);
    return r;
}

/* src/align.c:108-115 + 117-154 (CACHED_LOG): all float, no FMA, left-to-right association */
static inline float emission(float x, float scale, float shift, const abea_model_t* m) {
    float gp_mean = scale * m->level_mean + shift;
    float a = (x - gp_mean) / m->level_stdv;
    float lead = -0.918938f - m->level_log_stdv;
    float quad = -0.5f * a;
    quad = quad * a;
    return lead + quad;
}

/* src/align.c:207-216. n_events and n_kmers are size_t in the reference; the quotient is double. */
void abea_oracle_transitions(int64_t n_events, int64_t n_kmers, double* lp_skip, double* lp_stay,
                             double* lp_step, double* lp_trim) {
    double events_per_kmer = (double)(size_t)n_events / (size_t)n_kmers;
    double p_stay = 1 - (1 / (events_per_kmer + 1));
    *lp_skip = log(1e-10);
    *lp_stay = log(p_stay);
    *lp_step = log(1.0 - exp(*lp_skip) - exp(*lp_stay));
    *lp_trim = log(0.01);
}

/* src/model.c:93,179: level_log_stdv = log(level_stdv). The reference is compiled as C++ (Makefile:6), where
 * log(float) binds to the float overload, i.e. glibc logf — NOT a double log rounded to float (probed: the two
 * differ in 250 of the 262144 R10 entries). */
void abea_oracle_fill_log_stdv(abea_model_t* model, int64_t n) {
    for (int64_t i = 0; i < n; i++) model[i].level_log_stdv = logf(model[i].level_stdv);
}

/* The whole of align() (src/align.c:180-559) for one read.
 * out must hold n_events + seq_len pairs (src/f5c.c:724-726). Returns the pair count after QC
 * (0 when QC fails, src/align.c:534-543); stats (optional) reports the pre-QC quantities. */
int32_t abea_oracle_align(abea_pair_t* out, const char* seq, int32_t seq_len, const abea_event_t* ev,
                          int64_t n_events_i, const abea_model_t* model, uint32_t k, float scale,
                          float shift, abea_oracle_stats_t* stats) {
    const int64_t E = n_events_i;
    const int64_t K = (int64_t)seq_len - (int64_t)k + 1;
    const int64_t NB = (E + 1) + (K + 1); /* src/align.c:219-221 */

    double lp_skip, lp_stay, lp_step, lp_trim;
    abea_oracle_transitions(E, K, &lp_skip, &lp_stay, &lp_step, &lp_trim);

    uint32_t* ranks = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(K > 0 ? K : 1));
    for (int64_t i = 0; i < K; i++) ranks[i] = kmer_rank(seq + i, k);

    /* full score and trace matrices, band-major (src/align.c:242-259) */
    float* score = (float*)malloc(sizeof(float) * (size_t)NB * W);
    uint8_t* trace = (uint8_t*)malloc((size_t)NB * W);
    int32_t* ll_e = (int32_t*)malloc(sizeof(int32_t) * (size_t)NB); /* event index of offset 0 */
    int32_t* ll_k = (int32_t*)malloc(sizeof(int32_t) * (size_t)NB); /* k-mer index of offset 0 */
    for (int64_t i = 0; i < NB * W; i++) score[i] = -INFINITY;
    memset(trace, 0, (size_t)NB * W);

    /* bands 0 and 1 (src/align.c:277-291): cell (event -1, kmer -1) scores 0; trimming event 0 costs lp_trim */
    ll_e[0] = W / 2 - 1;
    ll_k[0] = -1 - W / 2;
    ll_e[1] = ll_e[0] + 1;
    ll_k[1] = ll_k[0];
    score[0 * W + (-1 - ll_k[0])] = 0.0f;
    score[1 * W + (ll_e[1] - 0)] = (float)lp_trim;
    trace[1 * W + (ll_e[1] - 0)] = FROM_U;

    int64_t fills = 0;
    for (int64_t b = 2; b < NB; b++) {
        const float* p1 = score + (b - 1) * W;
        const float* p2 = score + (b - 2) * W;
        float* cur = score + b * W;
        uint8_t* tr = trace + b * W;

        /* Suzuki's rule (src/align.c:304-322) */
        float ll = p1[0], ur = p1[W - 1];
        int right;
        if (ll == -INFINITY && ur == -INFINITY) right = (b % 2 == 1);
        else right = ll < ur;
        ll_e[b] = ll_e[b - 1] + (right ? 0 : 1);
        ll_k[b] = ll_k[b - 1] + (right ? 1 : 0);
        const int32_t eb = ll_e[b], kb = ll_k[b];

        /* trim column, k-mer -1 (src/align.c:324-333) */
        int32_t to = -1 - kb;
        if (to >= 0 && to < W) {
            int64_t te = (int64_t)eb - to;
            if (te >= 0 && te < E) {
                cur[to] = (float)(lp_trim * (double)(te + 1));
                tr[to] = FROM_U;
            } else {
                cur[to] = -INFINITY;
            }
        }

        /* offsets whose event and k-mer both exist (src/align.c:337-346) */
        int64_t lo = -(int64_t)kb;                    /* kmer >= 0       */
        if ((int64_t)eb - (E - 1) > lo) lo = (int64_t)eb - (E - 1); /* event <= E-1 */
        if (lo < 0) lo = 0;
        int64_t hi = K - (int64_t)kb;                 /* kmer <= K-1     */
        if ((int64_t)eb + 1 < hi) hi = (int64_t)eb + 1; /* event >= 0    */
        if (hi > W) hi = W;

        /* neighbour offsets relative to this band's offset o (src/align.c:354-356) */
        const int32_t d_up = ll_e[b - 1] - eb + 1;   /* o_up   = o + d_up   */
        const int32_t d_left = -1 - ll_k[b - 1] + kb; /* o_left = o + d_left */
        const int32_t d_diag = -1 - ll_k[b - 2] + kb; /* o_diag = o + d_diag */

        for (int64_t o = lo; o < hi; o++) {
            int32_t e = eb - (int32_t)o, km = kb + (int32_t)o;
            int64_t ou = o + d_up, ol = o + d_left, od = o + d_diag;
            float up = (ou >= 0 && ou < W) ? p1[ou] : -INFINITY;
            float left = (ol >= 0 && ol < W) ? p1[ol] : -INFINITY;
            float diag = (od >= 0 && od < W) ? p2[od] : -INFINITY;
            float lp = emission(ev[e].mean, scale, shift, &model[ranks[km]]);
            /* double sums rounded once (src/align.c:382-384) */
            float sd = (float)(((double)diag + lp_step) + (double)lp);
            float su = (float)(((double)up + lp_stay) + (double)lp);
            float sl = (float)((double)left + lp_skip);
            /* ties resolve L over U over D (src/align.c:386-392) */
            float best = sd;
            uint8_t from = FROM_D;
            best = su > best ? su : best;
            from = best == su ? FROM_U : from;
            best = sl > best ? sl : best;
            from = best == sl ? FROM_L : from;
            cur[o] = best;
            tr[o] = from;
            fills++;
        }
    }

    /* best end cell on the last k-mer column, trailing events trimmed (src/align.c:424-445) */
    float best_score = -INFINITY;
    int32_t ce = 0, ck = (int32_t)(K - 1);
    for (int64_t e = 0; e < E; e++) {
        int64_t b = (e + 1) + ((int64_t)ck + 1);
        int64_t o = (int64_t)ll_e[b] - e;
        if (o >= 0 && o < W) {
            float s = (float)((double)score[b * W + o] + (double)(size_t)(E - e) * lp_trim);
            if (s > best_score) {
                best_score = s;
                ce = (int32_t)e;
            }
        }
    }
    const int32_t end_event = ce;

    /* traceback (src/align.c:452-499) */
    int32_t n = 0, gap = 0, max_gap = 0;
    double sum_emission = 0;
    while (ck >= 0 && ce >= 0) {
        out[n].ref_pos = ck;
        out[n].read_pos = ce;
        n++;
        sum_emission += emission(ev[ce].mean, scale, shift, &model[kmer_rank(seq + ck, k)]);
        int64_t b = ((int64_t)ce + 1) + ((int64_t)ck + 1);
        int64_t o = (int64_t)ll_e[b] - ce;
        uint8_t from = trace[b * W + o];
        if (from == FROM_D) { ck--; ce--; gap = 0; }
        else if (from == FROM_U) { ce--; gap = 0; }
        else { ck--; gap++; if (gap > max_gap) max_gap = gap; }
    }

    /* ascending order (src/align.c:503-513) */
    for (int32_t i = 0, j = n - 1; i < j; i++, j--) {
        abea_pair_t t = out[i];
        out[i] = out[j];
        out[j] = t;
    }

    /* QC (src/align.c:526-543) */
    double avg = sum_emission / (double)n;
    int spanned = n > 0 && out[0].ref_pos == 0 && out[n - 1].ref_pos == (int32_t)(K - 1);
    if (stats) {
        stats->sum_emission = sum_emission;
        stats->end_score = best_score;
        stats->n_aligned = n;
        stats->max_gap = max_gap;
        stats->end_event = end_event;
        stats->spanned = spanned;
        stats->n_bands = NB;
        stats->n_fills = fills;
    }
    int32_t ret = n;
    if (avg < -5.0 || !spanned || max_gap > 50) ret = 0;

    free(ranks);
    free(score);
    free(trace);
    free(ll_e);
    free(ll_k);
    return ret;
}

/* estimate_scalings_using_mom (src/align.c:58-106): method-of-moments shift/scale, double accumulators */
void abea_oracle_estimate_scalings(const char* seq, int32_t seq_len, const abea_model_t* model, uint32_t k,
                                   const abea_event_t* ev, int64_t n_events, abea_scalings_t* out) {
    int32_t n_kmers = seq_len - (int32_t)k + 1;
    double ev_sum = 0.0;
    for (int64_t i = 0; i < n_events; i++) ev_sum += ev[i].mean;
    double km_sum = 0.0, km_sq = 0.0;
    for (int32_t i = 0; i < n_kmers; i++) {
        double l = model[kmer_rank(seq + i, k)].level_mean;
        km_sum += l;
        km_sq += l * l;
    }
    double shift = ev_sum / (size_t)n_events - km_sum / n_kmers;
    double ev_sq = 0.0;
    for (int64_t i = 0; i < n_events; i++) ev_sq += (ev[i].mean - shift) * (ev[i].mean - shift);
    double scale = (ev_sq / (size_t)n_events) / (km_sq / n_kmers);
    out->shift = (float)shift;
    out->scale = (float)scale;
    out->var = 0;
    out->log_var = 0;
}

/* ---- the stage after ABEA: scaling_single = postalign + recalibrate_model + read flags ------------- */

/* postalign (src/align.c:561-660). Walk the pair list once: a pair whose event differs from the previous pair's
 * opens / extends the event range of its k-mer (:585-598). Then every k-mer with a range contributes its events
 * start..stop to the alignment list (:611-655); an entry is 'M' when its k-mer rank differs from the rank of the
 * entry before it, so only the first event of a k-mer can be 'M' (:640). recalibrate_model reads nothing but the
 * 'M' entries (their k-mer rank and event index), so the list is returned as two int arrays of those.
 * m_kmer / m_event need room for n_kmers entries (may be NULL). Returns the length of the full list. */
static int32_t oracle_postalign(abea_index_pair_t* map, double* events_per_base, const char* seq, int32_t n_kmers,
                                const abea_pair_t* pairs, int32_t n_pairs, uint32_t k, int32_t* m_rank,
                                int32_t* m_event, int32_t* n_m) {
    for (int32_t i = 0; i < n_kmers; i++) map[i].start = map[i].stop = -1;
    int32_t max_event = 0, min_event = INT32_MAX, prev_event = -1;
    for (int32_t i = 0; i < n_pairs; i++) {
        int32_t ki = pairs[i].ref_pos, e = pairs[i].read_pos;
        if (e != prev_event) {
            if (map[ki].start == -1) map[ki].start = e;
            map[ki].stop = e;
        }
        if (e > max_event) max_event = e;
        if (e < min_event) min_event = e;
        prev_event = e;
    }
    *events_per_base = (double)(max_event - min_event) / n_kmers;
    int32_t n = 0, nm = 0, prev_rank = -1;
    for (int32_t ki = 0; ki < n_kmers; ki++) {
        if (map[ki].start == -1) continue;
        int32_t rank = (int32_t)kmer_rank(seq + ki, k);
        for (int32_t e = map[ki].start; e <= map[ki].stop; e++) {
            if (prev_rank != rank) {
                if (m_rank) m_rank[nm] = rank;
                if (m_event) m_event[nm] = e;
                nm++;
            }
            n++;
            prev_rank = rank;
        }
    }
    *n_m = nm;
    return n;
}

/* recalibrate_model (src/align.c:665-773) over the 'M' entries: weighted least squares for (shift, scale) from
 * the 2x2 normal equations accumulated in list order in double (:703-724), closed-form solve (:729-731), then
 * var = sqrt(mean of squared residuals / stdv^2) (:738-751). Narrowing to float happens when the results are
 * stored in scalings_t (:753-758). Returns 1 if recalibrated. */
static int oracle_recalibrate(const abea_model_t* model, const abea_event_t* ev, const int32_t* m_rank,
                              const int32_t* m_event, int32_t n_m, int32_t min_events, abea_scalings_t* sc,
                              double* var_d) {
    *var_d = 0;
    if (n_m < min_events) return 0;
    double A00 = 0, A01 = 0, A11 = 0, b0 = 0, b1 = 0;
    for (int32_t j = 0; j < n_m; j++) {
        double e = ev[m_event[j]].mean;
        double mu = model[m_rank[j]].level_mean;
        double sd = model[m_rank[j]].level_stdv;
        double inv_var = 1. / (sd * sd);
        A00 += inv_var;
        A01 += mu * inv_var;
        A11 += mu * mu * inv_var;
        b0 += e * inv_var;
        b1 += mu * e * inv_var;
    }
    double A10 = A01;
    double div = A00 * A11 - A01 * A10;
    double shift = -(A01 * b1 - A11 * b0) / div;
    double scale = (A00 * b1 - A10 * b0) / div;
    double var = 0.;
    for (int32_t j = 0; j < n_m; j++) {
        double e = ev[m_event[j]].mean;
        double mu = model[m_rank[j]].level_mean;
        double sd = model[m_rank[j]].level_stdv;
        double yi = (e - shift - scale * mu);
        var += yi * yi / (sd * sd);
    }
    var /= n_m;
    var = sqrt(var);
    sc->shift = (float)shift;
    sc->scale = (float)scale;
    sc->var = (float)var;
    sc->log_var = (float)log(var);
    *var_d = var;
    return 1;
}

/* scaling_single (src/f5c.c:736-807): sc is db->scalings[i], in = method-of-moments, out = as the reference
 * leaves it; map needs n_kmers entries when n_pairs > 0. */
void abea_oracle_scaling_single(const abea_pair_t* pairs, int32_t n_pairs, const char* seq, int32_t seq_len,
                                const abea_event_t* ev, int64_t n_events, const abea_model_t* model, uint32_t k,
                                int32_t min_num_events_to_rescale, abea_scalings_t* sc, abea_index_pair_t* map,
                                abea_scaling_result_t* out) {
    (void)n_events;
    memset(out, 0, sizeof(*out));
    int32_t n_kmers = seq_len - (int32_t)k + 1;
    if (n_pairs > 0) {
        int32_t* m_rank = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_kmers > 0 ? n_kmers : 1));
        int32_t* m_event = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_kmers > 0 ? n_kmers : 1));
        int32_t n_m = 0;
        out->n_event_alignment = oracle_postalign(map, &out->events_per_base, seq, n_kmers, pairs, n_pairs, k,
                                                  m_rank, m_event, &n_m);
        out->num_m_state = n_m;
        out->calibrated = oracle_recalibrate(model, ev, m_rank, m_event, n_m, min_num_events_to_rescale, sc,
                                             &out->var_d);
        free(m_rank);
        free(m_event);
        if (!out->calibrated || sc->var > ABEA_MIN_CALIBRATION_VAR) out->flags |= ABEA_FAILED_CALIBRATION; /* :776-781 */
        else if (out->events_per_base > ABEA_MAX_EVENTS_PER_BASE) out->flags |= ABEA_FAILED_QUALITY_CHK;   /* :798-803 */
    } else {
        out->flags |= ABEA_FAILED_ALIGNMENT; /* :787-793 */
    }
    out->scalings = *sc;
}

/* ---- the stage before everything: getevents (src/events.c:562-582) ----------------------------------------- */

/* getevents ignores what trim_and_segment_raw returns (src/events.c:572), so event detection runs over the whole
 * signal: cumulative sums in double (:297-307; the square is a FLOAT product), two windowed t-statistics
 * (:320-372), the short/long peak detector (:379-448), and one event per gap between peaks (:463-515).
 * Mixed float/double expressions are restated with the C++ promotion rules the reference is compiled under
 * (fabs / sqrt on float arguments are the float overloads). Cases the reference leaves undefined — fewer than 100
 * samples (an assert in trim_raw_by_mad), no peak at all (create_events reads peaks[-1]) — yield 0 events / one
 * event over the whole signal here. */
typedef struct {
    int32_t w1, w2;
    float thr1, thr2, peak_height;
} det_param_t;
static const det_param_t DET_DNA = {3, 6, 1.4f, 9.0f, 0.2f};   /* src/events.c:52-56 */
static const det_param_t DET_RNA = {7, 14, 2.5f, 9.0f, 1.0f};  /* src/events.c:59-63 */

static float tstat_at(const double* sum, const double* sumsq, int64_t n, int64_t i, int32_t w) {
    if (n < 2 * (int64_t)w || w < 2) return 0.f;
    if (i < w || i > n - w) return 0.f;
    const float wf = (float)w;
    double sum1 = sum[i], sumsq1 = sumsq[i];
    if (i > w) {
        sum1 -= sum[i - w];
        sumsq1 -= sumsq[i - w];
    }
    float sum2 = (float)(sum[i + w] - sum[i]);
    float sumsq2 = (float)(sumsq[i + w] - sumsq[i]);
    float mean1 = (float)(sum1 / wf);
    float mean2 = sum2 / wf;
    float combined_var = (float)(sumsq1 / wf - (double)(mean1 * mean1) + (double)(sumsq2 / wf) - (double)(mean2 * mean2));
    combined_var = fmaxf(combined_var, FLT_MIN);
    const float delta_mean = mean2 - mean1;
    return fabsf(delta_mean) / sqrtf(combined_var / wf);
}

typedef struct {
    int64_t masked_to;
    int64_t peak_pos; /* -1 = none */
    float peak_value, threshold;
    int32_t window;
    int valid;
} det_t;

/* events must have room for n_samples / 2 + 2 entries. Returns the number of events. */
int64_t abea_oracle_getevents(int64_t n_samples, const float* raw, int8_t rna, abea_event_t* events) {
    const det_param_t P = rna ? DET_RNA : DET_DNA;
    const int64_t n = n_samples;
    if (n < 100) return 0;
    double* sum = (double*)malloc(sizeof(double) * (size_t)(n + 1));
    double* sumsq = (double*)malloc(sizeof(double) * (size_t)(n + 1));
    int64_t* peaks = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    sum[0] = 0.0;
    sumsq[0] = 0.0;
    for (int64_t i = 0; i < n; i++) {
        sum[i + 1] = sum[i] + raw[i];
        sumsq[i + 1] = sumsq[i] + raw[i] * raw[i];
    }
    det_t d[2] = {{0, -1, FLT_MAX, P.thr1, P.w1, 0}, {0, -1, FLT_MAX, P.thr2, P.w2, 0}};
    int64_t n_peaks = 0;
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < 2; k++) {
            det_t* t = &d[k];
            if (t->masked_to >= i) continue;
            const float v = tstat_at(sum, sumsq, n, i, t->window);
            if (t->peak_pos == -1) {
                if (v < t->peak_value) {
                    t->peak_value = v;
                } else if (v - t->peak_value > P.peak_height) {
                    t->peak_value = v;
                    t->peak_pos = i;
                }
            } else {
                if (v > t->peak_value) {
                    t->peak_value = v;
                    t->peak_pos = i;
                }
                if (k == 0 && t->peak_value > t->threshold) { /* the short detector silences the long one */
                    d[1].masked_to = t->peak_pos + t->window;
                    d[1].peak_pos = -1;
                    d[1].peak_value = FLT_MAX;
                    d[1].valid = 0;
                }
                if (t->peak_value - v > P.peak_height && t->peak_value > t->threshold) t->valid = 1;
                if (t->valid && (i - t->peak_pos) > t->window / 2) {
                    peaks[n_peaks++] = t->peak_pos;
                    t->peak_pos = -1;
                    t->peak_value = v;
                    t->valid = 0;
                }
            }
        }
    }
    /* create_events counts the entries of the zero-padded peak list that are > 0 and < nsample (:485-489) */
    int64_t n_ev = 1;
    for (int64_t j = 0; j < n_peaks; j++) n_ev += (peaks[j] > 0 && peaks[j] < n);
    for (int64_t e = 0; e < n_ev; e++) {
        const int64_t start = e == 0 ? 0 : peaks[e - 1];
        const int64_t end = (e == n_ev - 1) ? n : peaks[e];
        abea_event_t ev;
        ev.start = (uint64_t)start;
        ev.length = (float)(end - start);
        ev.mean = (float)(sum[end] - sum[start]) / ev.length;
        const float deltasqr = (float)(sumsq[end] - sumsq[start]);
        const float var = deltasqr / ev.length - ev.mean * ev.mean;
        ev.stdv = sqrtf(fmaxf(var, 0.0f));
        events[e] = ev;
    }
    free(sum);
    free(sumsq);
    free(peaks);
    return n_ev;
}

/* ---- batch driver: CPU branch of align_db (src/f5c.c:811-845) over a flat batch ------------------- */

typedef struct {
    const abea_batch_t* b;
    const abea_model_t* model;
    uint32_t k;
    abea_pair_t* pairs;
    const int64_t* pair_ptr;
    int32_t* n_pairs;
    abea_oracle_stats_t* stats;
    volatile int32_t* next;
} pool_t;

static void* worker(void* p) {
    pool_t* a = (pool_t*)p;
    const abea_batch_t* b = a->b;
    for (;;) {
        int32_t i = __sync_fetch_and_add(a->next, 1);
        if (i >= b->n_reads) break;
        int good = b->good ? b->good[i] : 1;
        if (a->stats) memset(&a->stats[i], 0, sizeof(abea_oracle_stats_t));
        /* align_single's filter: good read and events/base < 15 in float (src/f5c.c:813-814) */
        if (good && (size_t)b->n_events[i] / (float)b->read_len[i] < ABEA_AVG_EVENTS_PER_KMER_MAX) {
            a->n_pairs[i] = abea_oracle_align(a->pairs + a->pair_ptr[i], b->seq + b->seq_ptr[i], b->read_len[i],
                                              b->events + b->event_ptr[i], b->n_events[i], a->model, a->k,
                                              b->scalings[i].scale, b->scalings[i].shift,
                                              a->stats ? &a->stats[i] : NULL);
        } else {
            a->n_pairs[i] = 0;
        }
    }
    return NULL;
}

/* Returns wall-clock seconds spent aligning. stats may be NULL. */
double abea_oracle_align_batch(const abea_batch_t* b, const abea_model_t* model, uint32_t k, abea_pair_t* pairs,
                               const int64_t* pair_ptr, int32_t* n_pairs, abea_oracle_stats_t* stats,
                               int32_t n_threads) {
    volatile int32_t next = 0;
    pool_t a = {b, model, k, pairs, pair_ptr, n_pairs, stats, &next};
    if (n_threads < 1) n_threads = 1;
    pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 1; t < n_threads; t++) pthread_create(&tid[t], NULL, worker, &a);
    worker(&a);
    for (int t = 1; t < n_threads; t++) pthread_join(tid[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(tid);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
