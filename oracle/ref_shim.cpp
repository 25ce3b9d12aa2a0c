/* ref_shim.cpp — TEST INFRASTRUCTURE. extern "C" doorway into the UNMODIFIED reference object code.
 *
 * This file is ours; the algorithm it exposes is the reference's own: it is compiled together with
 * /root/reference/src/{align,model,events}.c where they lie (see oracle/Makefile) into
 * oracle/_ref/libf5c_ref.so. Nothing here restates ABEA; it only (1) gives C linkage to the reference's
 * C++-linkage functions, (2) stands in for the two CpG tables model.c declares extern, and (3) runs the
 * CPU branch of align_db (pthread_db(core, db, align_single), src/f5c.c:833-845) as a dynamic work
 * queue over reads, because f5c.c itself cannot link without htslib.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "f5c.h"
#include "f5cmisc.h"

#include "../include/abea_types.h"

/* model.c:17-18 wants these two symbols (defined in the 1.97M-line methmodel.c, which ABEA never reads) */
float r9_4_450bps_cpg_6mer_template_model_builtin_data[2] = {0, 0};
float r10_4_400bps_cpg_9mer_template_model_builtin_data[2] = {0, 0};

static_assert(sizeof(abea_event_t) == sizeof(event_t), "event_t layout");
static_assert(sizeof(abea_model_t) == sizeof(model_t), "model_t layout");
static_assert(sizeof(abea_scalings_t) == sizeof(scalings_t), "scalings_t layout");
static_assert(sizeof(abea_pair_t) == sizeof(AlignedPair), "AlignedPair layout");

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

extern "C" {

/* set_model (src/model.c:132): fills model[0..4^k) and returns k. Caller provides ABEA_MAX_NUM_KMER slots. */
uint32_t f5cref_set_model(abea_model_t* model, uint32_t model_id) {
    return set_model((model_t*)model, model_id);
}

/* read_model (src/model.c:39): load a text model file */
uint32_t f5cref_read_model(abea_model_t* model, const char* file) {
    return read_model((model_t*)model, file, MODEL_TYPE_NUCLEOTIDE);
}

/* align (src/align.c:180) for one read */
int32_t f5cref_align(abea_pair_t* out, const char* seq, int32_t seq_len, const abea_event_t* ev,
                     int64_t n_events, const abea_model_t* model, uint32_t kmer_size, float scale,
                     float shift) {
    event_table et;
    et.n = (size_t)n_events;
    et.start = 0;
    et.end = (size_t)n_events;
    et.event = (event_t*)ev;
    scalings_t sc;
    memset(&sc, 0, sizeof(sc));
    sc.scale = scale;
    sc.shift = shift;
    return align((AlignedPair*)out, (char*)seq, seq_len, et, (model_t*)model, kmer_size, sc, 4000.0f);
}

/* estimate_scalings_using_mom (src/align.c:58) */
void f5cref_estimate_scalings(const char* seq, int32_t seq_len, const abea_model_t* model,
                              uint32_t kmer_size, const abea_event_t* ev, int64_t n_events,
                              abea_scalings_t* out) {
    event_table et;
    et.n = (size_t)n_events;
    et.start = 0;
    et.end = (size_t)n_events;
    et.event = (event_t*)ev;
    scalings_t sc = estimate_scalings_using_mom((char*)seq, seq_len, (model_t*)model, kmer_size, et);
    out->scale = sc.scale;
    out->shift = sc.shift;
    out->var = 0;
    out->log_var = 0;
}

/* getevents (src/events.c:562). Returns the number of events; copies min(n, cap) into out. */
int64_t f5cref_getevents(int64_t nsample, float* raw_pa, int8_t rna, abea_event_t* out, int64_t cap) {
    event_table et = getevents((size_t)nsample, raw_pa, rna);
    int64_t n = (int64_t)et.n;
    int64_t m = n < cap ? n : cap;
    if (out && m > 0) memcpy(out, et.event, (size_t)m * sizeof(event_t));
    free(et.event);
    return n;
}

/* scaling_single (src/f5c.c:736-807) for one read: the reference's own postalign (src/align.c:561) and
 * recalibrate_model (src/align.c:665) called in the order and with the flag logic of scaling_single, which lives
 * in f5c.c and cannot link here. sc is db->scalings[i] (in: method-of-moments values; out: as the reference leaves
 * it). map must hold n_kmers entries when n_pairs > 0 (the reference mallocs it there, src/f5c.c:746). */
void f5cref_scaling_single(const abea_pair_t* pairs, int32_t n_pairs, const char* seq, int32_t seq_len,
                           const abea_event_t* ev, int64_t n_events, const abea_model_t* model, uint32_t kmer_size,
                           int32_t min_num_events_to_rescale, abea_scalings_t* sc, abea_index_pair_t* map,
                           abea_scaling_result_t* out) {
    static_assert(sizeof(abea_index_pair_t) == sizeof(index_pair_t), "index_pair_t layout");
    memset(out, 0, sizeof(*out));
    int32_t n_kmers = seq_len - (int32_t)kmer_size + 1;
    scalings_t s;
    memcpy(&s, sc, sizeof(s));
    if (n_pairs > 0) {
        event_alignment_t* ea = (event_alignment_t*)malloc(sizeof(event_alignment_t) * (size_t)(n_pairs + 1));
        double epb = 0;
        int32_t n_ea = postalign(ea, (index_pair_t*)map, &epb, (char*)seq, n_kmers, (AlignedPair*)pairs, n_pairs,
                                 kmer_size);
        out->events_per_base = epb;
        out->n_event_alignment = n_ea;
        for (int32_t j = 0; j < n_ea; j++) out->num_m_state += (ea[j].hmm_state == 'M');
        event_table et;
        et.n = (size_t)n_events;
        et.start = 0;
        et.end = (size_t)n_events;
        et.event = (event_t*)ev;
        bool calibrated = recalibrate_model((model_t*)model, kmer_size, et, &s, ea, n_ea, 1, min_num_events_to_rescale);
        free(ea);
        out->calibrated = calibrated ? 1 : 0;
        if (!calibrated || s.var > MIN_CALIBRATION_VAR) out->flags |= FAILED_CALIBRATION;
        else if (epb > 5.0) out->flags |= FAILED_QUALITY_CHK;
    } else {
        out->flags |= FAILED_ALIGNMENT;
    }
    memcpy(sc, &s, sizeof(s));
    memcpy(&out->scalings, &s, sizeof(s));
    out->var_d = 0; /* internal to recalibrate_model; only the restatement and the CUDA path report it */
}

typedef struct {
    const abea_batch_t* b;
    const abea_model_t* model;
    uint32_t kmer_size;
    abea_pair_t* pairs;
    const int64_t* pair_ptr;
    int32_t* n_pairs;
    volatile int32_t* next;
} pool_arg_t;

static void* pool_worker(void* p) {
    pool_arg_t* a = (pool_arg_t*)p;
    const abea_batch_t* b = a->b;
    for (;;) {
        int32_t i = __sync_fetch_and_add(a->next, 1);
        if (i >= b->n_reads) break;
        /* align_single's filter, src/f5c.c:811-830 */
        int good = b->good ? b->good[i] : 1;
        if (good && (b->n_events[i]) / (float)(b->read_len[i]) < AVG_EVENTS_PER_KMER_MAX) {
            a->n_pairs[i] = f5cref_align(a->pairs + a->pair_ptr[i], b->seq + b->seq_ptr[i], b->read_len[i],
                                         b->events + b->event_ptr[i], b->n_events[i], a->model,
                                         a->kmer_size, b->scalings[i].scale, b->scalings[i].shift);
        } else {
            a->n_pairs[i] = 0;
        }
    }
    return NULL;
}

/* CPU branch of align_db over a flat batch with n_threads workers pulling reads from a shared counter.
 * pairs[pair_ptr[i] ..) must have room for n_events[i]+read_len[i] pairs (src/f5c.c:724-726).
 * Returns wall-clock seconds spent inside the pool. */
double f5cref_align_batch(const abea_batch_t* b, const abea_model_t* model, uint32_t kmer_size,
                          abea_pair_t* pairs, const int64_t* pair_ptr, int32_t* n_pairs, int32_t n_threads) {
    volatile int32_t next = 0;
    pool_arg_t a = {b, model, kmer_size, pairs, pair_ptr, n_pairs, &next};
    if (n_threads < 1) n_threads = 1;
    pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    double t0 = now_s();
    for (int t = 1; t < n_threads; t++) pthread_create(&tid[t], NULL, pool_worker, &a);
    pool_worker(&a);
    for (int t = 1; t < n_threads; t++) pthread_join(tid[t], NULL);
    double t1 = now_s();
    free(tid);
    return t1 - t0;
}

} /* extern "C" */
