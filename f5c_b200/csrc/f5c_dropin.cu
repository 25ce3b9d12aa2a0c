/* f5c_dropin.cu — the reference-facing entry points, to be compiled INSIDE the f5c tree in place of
 * src/f5c.cu + src/align.cu:
 *
 *     void init_cuda(core_t* core);              reference src/f5c.h:575-581, src/f5c.cu:23-202
 *     void free_cuda(core_t* core);              reference src/f5c.cu:204-234
 *     void align_cuda(core_t* core, db_t* db);   reference src/f5cmisc.h:122-125, src/f5c.cu:647-1061
 *
 * plus two entry points the reference has no GPU counterpart for (it runs these stages per read on CPU threads):
 *
 *     void scaling_cuda(core_t* core, db_t* db);   replaces pthread_db(core, db, scaling_single) in process_db,
 *                                                  src/f5c.c:932 (scaling_single: src/f5c.c:736-807)
 *     void estimate_scalings_cuda(core_t*, db_t*); the estimate_scalings_using_mom call of event_single for a whole
 *                                                  batch, src/f5c.c:709-711 (src/align.c:58-106)
 *     void getevents_cuda(core_t*, db_t*);         the pA conversion + getevents() call of event_single for a whole
 *                                                  batch, src/f5c.c:692-703 (src/events.c:562-582)
 *
 * with the reference's own C++ linkage, structs (core_t/db_t from the reference's f5c.h, -DHAVE_CUDA=1) and error
 * convention (message on stderr + exit, src/f5cmisc.cuh:54-118). It is a thin packer over the C ABI in
 * include/abea_b200.h: ragged db_t -> flat pinned staging -> abea_align_batch -> db->event_align_pairs[i] /
 * db->n_event_align_pairs[i]. It needs the reference headers, so it is built only where the f5c tree is available
 * (__graft_entry__.build_dropin); nothing in it is copied from the reference.
 *
 * Differences from the reference's align_cuda, all deliberate: no read is diverted to CPU threads (src/f5c.cu:440-452,
 * 701-735 have no counterpart), no load/memory "advisor" messages (:457-644), device memory grows on demand instead of
 * being pre-sized from cuda_mem_frac (:121-146).
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "f5c.h"
#include "f5cmisc.h"

#include "../../include/abea_b200.h"

static_assert(sizeof(abea_event_t) == sizeof(event_t), "event_t layout differs from abea_event_t");
static_assert(sizeof(abea_model_t) == sizeof(model_t), "model_t layout differs (CACHED_LOG must be defined)");
static_assert(sizeof(abea_scalings_t) == sizeof(scalings_t), "scalings_t layout differs");
static_assert(sizeof(abea_pair_t) == sizeof(AlignedPair), "AlignedPair layout differs");
static_assert(offsetof(abea_scalings_t, scale) == offsetof(scalings_t, scale) && offsetof(abea_scalings_t, shift) == offsetof(scalings_t, shift), "scalings_t field order differs");
static_assert(offsetof(abea_event_t, mean) == offsetof(event_t, mean), "event_t.mean offset differs");
static_assert(ABEA_BANDWIDTH == ALN_BANDWIDTH, "band width differs");
static_assert(sizeof(abea_index_pair_t) == sizeof(index_pair_t), "index_pair_t layout differs");
static_assert(ABEA_FAILED_CALIBRATION == FAILED_CALIBRATION && ABEA_FAILED_ALIGNMENT == FAILED_ALIGNMENT &&
              ABEA_FAILED_QUALITY_CHK == FAILED_QUALITY_CHK, "read_stat_flag bits differ");

namespace {

/* what core->cuda points to: the reference's own struct first (so the pointer type is honoured), ours after it */
struct dropin_data {
    cuda_data_t base;
    abea_ctx_t* ctx;
    /* pinned staging, grown on demand */
    char* seq; size_t seq_cap;
    abea_event_t* events; size_t ev_cap;
    abea_pair_t* pairs; size_t pair_cap;
    int64_t* seq_ptr; int64_t* event_ptr; int64_t* pair_ptr;
    int32_t* read_len; int32_t* n_events; int32_t* n_pairs;
    abea_scalings_t* scalings; uint8_t* good;
    size_t read_cap;
    /* scaling_cuda staging */
    abea_scaling_result_t* sres; size_t sres_cap;
    abea_index_pair_t* maps; size_t map_cap;
    int64_t* map_ptr; size_t map_ptr_cap;
    const db_t* aligned_db; /* the batch whose pair lists are resident on the device */
    const abea_event_t** rag_events; int32_t* rag_nev; uint8_t* rag_good; size_t rag_cap; /* align_cuda's per-read arrays */
    /* getevents_cuda staging */
    float* raw; size_t raw_cap;
    int64_t* raw_ptr; int32_t* n_samples; float* cal_off; float* cal_range; float* cal_dig; size_t sig_cap;
    abea_event_t* ev_out; size_t ev_out_cap;
};

void die(const char* func, const char* what, abea_ctx_t* ctx) {
    fprintf(stderr, "[%s::ERROR]\033[1;31m %s: %s\033[0m\n", func, what, ctx ? abea_last_error(ctx) : "");
    exit(-1);
}

/* The per-read copies between db_t's ragged arrays and the flat staging buffers are the bulk of the host time of a
 * batch (cfg2: 389 MB of events in, 130 MB of pairs out), so they run on core->opt.num_thread threads — the threads
 * the reference uses for pthread_db (src/f5c.c:590-676) — pulling blocks of reads from a shared counter. */
template <typename F> void parallel_reads(int32_t n, int threads, F f) {
    if (threads <= 1 || n < 64) {
        for (int32_t i = 0; i < n; i++) f(i);
        return;
    }
    std::atomic<int32_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int32_t i0 = next.fetch_add(16);
            if (i0 >= n) break;
            const int32_t i1 = std::min(n, i0 + 16);
            for (int32_t i = i0; i < i1; i++) f(i);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (std::thread& t : pool) t.join();
}

/* db->read[i] / db->et[i].event -> flat seq / events at the given offsets */
void copy_in(const db_t* db, const int64_t* seq_ptr, const int64_t* event_ptr, char* seq, abea_event_t* events, int threads) {
    parallel_reads(db->n_bam_rec, threads, [&](int32_t i) {
        memcpy(seq + seq_ptr[i], db->read[i], (size_t)db->read_len[i]);
        seq[seq_ptr[i] + db->read_len[i]] = 0;
        if (db->et[i].n) memcpy(events + event_ptr[i], db->et[i].event, db->et[i].n * sizeof(event_t));
    });
}

/* flat pairs -> db->event_align_pairs[i] (src/f5c.cu:1005-1030): already ascending, no host-side reversal */
void copy_out(db_t* db, const int32_t* n_pairs, const int64_t* pair_ptr, const abea_pair_t* pairs, int threads) {
    parallel_reads(db->n_bam_rec, threads, [&](int32_t i) {
        db->n_event_align_pairs[i] = n_pairs[i];
        if (n_pairs[i] > 0) memcpy(db->event_align_pairs[i], pairs + pair_ptr[i], (size_t)n_pairs[i] * sizeof(AlignedPair));
    });
}

int host_threads(const core_t* core) { return core->opt.num_thread > 0 ? core->opt.num_thread : 1; }

template <typename T> void grow(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return;
    if (p) abea_host_free(p);
    cap = need + need / 4 + 64;
    p = (T*)abea_host_alloc(cap * sizeof(T));
    if (!p) { fprintf(stderr, "[align_cuda::ERROR] pinned allocation of %zu bytes failed\n", cap * sizeof(T)); exit(EXIT_FAILURE); }
}

} // namespace

void init_cuda(core_t* core) {
    dropin_data* d = (dropin_data*)calloc(1, sizeof(dropin_data));
    if (!d) { fprintf(stderr, "[init_cuda::ERROR] out of memory\n"); exit(EXIT_FAILURE); }
    int rc = abea_create(&d->ctx, core->opt.cuda_dev_id);
    if (rc == ABEA_ERR_NODEVICE) { fprintf(stderr, "[init_cuda::ERROR] no CUDA capable device %d\n", core->opt.cuda_dev_id); exit(1); }
    if (rc) die("init_cuda", "abea_create failed", NULL);
    if (abea_set_model(d->ctx, (const abea_model_t*)core->model, core->kmer_size)) die("init_cuda", "model upload", d->ctx);
    if (core->opt.verbosity > 1) {
        int sms = 0; char name[256];
        abea_device_info(d->ctx, &sms, name);
        fprintf(stderr, "[init_cuda] %s on %s (%d SMs), k=%u\n", abea_version(), name, sms, core->kmer_size);
    }
    core->cuda = &d->base;
    core->align_kernel_time = core->align_pre_kernel_time = core->align_core_kernel_time = 0;
    core->align_post_kernel_time = core->align_cuda_malloc = core->align_cuda_memcpy = 0;
    core->align_cuda_postprocess = core->align_cuda_preprocess = core->align_cuda_total_kernel = 0;
    core->extra_load_cpu = 0;
}

void free_cuda(core_t* core) {
    dropin_data* d = (dropin_data*)core->cuda;
    if (!d) return;
    abea_destroy(d->ctx);
    void* bufs[] = {d->seq, d->events, d->pairs, d->seq_ptr, d->event_ptr, d->pair_ptr, d->read_len, d->n_events,
                    d->n_pairs, d->scalings, d->good, d->sres, d->maps, d->map_ptr, d->raw, d->raw_ptr, d->n_samples,
                    d->cal_off, d->cal_range, d->cal_dig, d->ev_out};
    for (void* b : bufs) abea_host_free(b);
    free(d->rag_events); free(d->rag_nev); free(d->rag_good);
    free(d);
    core->cuda = NULL;
}

/* flatten the ragged batch (what the reference does at src/f5c.cu:744-800) into pinned staging */
static void pack_db(dropin_data* d, const db_t* db, abea_batch_t& b, bool with_scalings, int threads) {
    const int32_t n = db->n_bam_rec;
    if ((size_t)n > d->read_cap) {
        size_t c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0, c8 = 0;
        grow(d->seq_ptr, c1, (size_t)n); grow(d->event_ptr, c2, (size_t)n); grow(d->pair_ptr, c3, (size_t)n);
        grow(d->read_len, c4, (size_t)n); grow(d->n_events, c5, (size_t)n); grow(d->n_pairs, c6, (size_t)n);
        grow(d->scalings, c7, (size_t)n); grow(d->good, c8, (size_t)n);
        d->read_cap = c1;
    }
    int64_t sp = 0, ep = 0, pp = 0;
    for (int32_t i = 0; i < n; i++) {
        d->seq_ptr[i] = sp; d->event_ptr[i] = ep; d->pair_ptr[i] = pp;
        d->read_len[i] = db->read_len[i];
        d->n_events[i] = (int32_t)db->et[i].n;
        d->good[i] = (db->sig[i] && db->sig[i]->nsample > 0) ? 1 : 0;   /* align_single, src/f5c.c:811 */
        d->scalings[i].scale = db->scalings[i].scale; d->scalings[i].shift = db->scalings[i].shift;
        d->scalings[i].var = db->scalings[i].var; d->scalings[i].log_var = db->scalings[i].log_var;
        sp += db->read_len[i] + 1; ep += (int64_t)db->et[i].n; pp += (int64_t)db->et[i].n + db->read_len[i];
    }
    grow(d->seq, d->seq_cap, (size_t)sp + 1);
    grow(d->events, d->ev_cap, (size_t)ep + 1);
    grow(d->pairs, d->pair_cap, (size_t)pp + 1);
    copy_in(db, d->seq_ptr, d->event_ptr, d->seq, d->events, threads);
    b.n_reads = n; b.seq = d->seq; b.seq_ptr = d->seq_ptr; b.read_len = d->read_len; b.events = d->events;
    b.event_ptr = d->event_ptr; b.n_events = d->n_events; b.scalings = with_scalings ? d->scalings : NULL;
    b.good = d->good;
    b.event_means = NULL;
}

void align_cuda(core_t* core, db_t* db) {
    dropin_data* d = (dropin_data*)core->cuda;
    double t0 = realtime();
    /* db_t is ragged (one malloc per read); the library takes it as it is and does the flattening on
     * core->opt.num_thread threads while its kernels run (abea_align_ragged). Only the per-read pointer / count
     * arrays are built here. */
    const int32_t n = db->n_bam_rec;
    if ((size_t)n > d->rag_cap) {
        free(d->rag_events); free(d->rag_nev); free(d->rag_good);
        d->rag_cap = (size_t)n + (size_t)n / 4 + 64;
        d->rag_events = (const abea_event_t**)malloc(d->rag_cap * sizeof(abea_event_t*));
        d->rag_nev = (int32_t*)malloc(d->rag_cap * sizeof(int32_t));
        d->rag_good = (uint8_t*)malloc(d->rag_cap);
        MALLOC_CHK(d->rag_events); MALLOC_CHK(d->rag_nev); MALLOC_CHK(d->rag_good);
    }
    for (int32_t i = 0; i < n; i++) {
        d->rag_events[i] = (const abea_event_t*)db->et[i].event;
        d->rag_nev[i] = (int32_t)db->et[i].n;
        d->rag_good[i] = (db->sig[i] && db->sig[i]->nsample > 0) ? 1 : 0;   /* align_single, src/f5c.c:811 */
    }
    abea_ragged_t rg;
    rg.n_reads = n;
    rg.seq = (const char* const*)db->read;
    rg.read_len = db->read_len;
    rg.events = d->rag_events;
    rg.n_events = d->rag_nev;
    rg.scalings = (const abea_scalings_t*)db->scalings;
    rg.good = d->rag_good;
    rg.pairs = (abea_pair_t* const*)db->event_align_pairs;
    rg.n_pairs = db->n_event_align_pairs;
    double t1 = realtime();

    abea_timing_t tm;
    if (abea_align_ragged(d->ctx, &rg, host_threads(core), &tm)) die("align_cuda", "Cuda error", d->ctx);
    d->aligned_db = db;
    double t2 = realtime();

    /* the reference's timer split (src/f5c.h:457-466), printed by meth_main (src/meth_main.c:767-788); packing and
     * unpacking overlap the kernels here, so only what is not hidden behind them is charged to pre / postprocess */
    core->align_cuda_preprocess += (t1 - t0) + tm.pack_ms * 1e-3;
    core->align_cuda_memcpy += (tm.h2d_ms + tm.d2h_ms) * 1e-3;
    core->align_kernel_time += tm.kernel_ms * 1e-3;
    core->align_pre_kernel_time += tm.kmer_ms * 1e-3;
    core->align_core_kernel_time += tm.fill_ms * 1e-3;
    core->align_post_kernel_time += tm.trace_ms * 1e-3;
    core->align_cuda_total_kernel += tm.kernel_ms * 1e-3;
    const double rest = (t2 - t1) - (tm.pack_ms + tm.h2d_ms + tm.kernel_ms) * 1e-3;
    core->align_cuda_postprocess += rest > 0 ? rest : 0;
    if (core->opt.verbosity > 1)
        fprintf(stderr, "[align_cuda] Load : GPU %d entries (%.1fM events), CPU 0 entries; kernels %.3f ms\n",
                tm.n_scheduled, tm.n_events / 1e6, tm.kernel_ms);
}

/* The estimate_scalings_using_mom call of event_single (src/f5c.c:709-711) for the whole batch: db->scalings[i].shift
 * and .scale of every read with events. RNA batches must call it BEFORE event_single's reversal (src/f5c.c:713-721);
 * the reversal itself stays where it is. */
void estimate_scalings_cuda(core_t* core, db_t* db) {
    dropin_data* d = (dropin_data*)core->cuda;
    abea_batch_t b;
    pack_db(d, db, b, false, host_threads(core));
    if (abea_upload_batch(d->ctx, &b, NULL)) die("estimate_scalings_cuda", "Cuda error", d->ctx);
    if (abea_estimate_scalings(d->ctx, 0, d->scalings, NULL)) die("estimate_scalings_cuda", "Cuda error", d->ctx);
    d->aligned_db = NULL;
    for (int32_t i = 0; i < db->n_bam_rec; i++)
        if (d->good[i] && db->et[i].n >= 1 && db->read_len[i] >= (int32_t)core->kmer_size) {
            db->scalings[i].shift = d->scalings[i].shift;
            db->scalings[i].scale = d->scalings[i].scale;
        }
}

/* The first half of event_single (src/f5c.c:684-703) for the whole batch: db->sig[i]->rawptr is converted to pA in
 * place exactly as the reference does (consumers downstream read it), and db->et[i] receives the event table of
 * getevents(nsample, rawptr, rna) — allocated here with the reference's allocator so that free_db_tmp releases it.
 * event_single then only has to do what follows (scalings, RNA reversal, the pair buffer). */
void getevents_cuda(core_t* core, db_t* db) {
    dropin_data* d = (dropin_data*)core->cuda;
    const int32_t n = db->n_bam_rec;
    if ((size_t)n > d->sig_cap) {
        size_t c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
        grow(d->raw_ptr, c1, (size_t)n); grow(d->n_samples, c2, (size_t)n); grow(d->cal_off, c3, (size_t)n);
        grow(d->cal_range, c4, (size_t)n); grow(d->cal_dig, c5, (size_t)n);
        d->sig_cap = c1;
    }
    int64_t total = 0;
    for (int32_t i = 0; i < n; i++) {
        const int32_t ns = (db->sig[i] && db->sig[i]->nsample > 0) ? (int32_t)db->sig[i]->nsample : 0;
        d->raw_ptr[i] = total;
        d->n_samples[i] = ns;
        d->cal_off[i] = ns ? db->sig[i]->offset : 0.f;
        d->cal_range[i] = ns ? db->sig[i]->range : 1.f;
        d->cal_dig[i] = ns ? db->sig[i]->digitisation : 1.f;
        total += ns;
    }
    grow(d->raw, d->raw_cap, (size_t)total + 1);
    parallel_reads(n, host_threads(core), [&](int32_t i) {
        if (d->n_samples[i]) memcpy(d->raw + d->raw_ptr[i], db->sig[i]->rawptr, (size_t)d->n_samples[i] * sizeof(float));
    });
    abea_signals_t sg;
    memset(&sg, 0, sizeof(sg));
    sg.n_reads = n; sg.raw = d->raw; sg.raw_ptr = d->raw_ptr; sg.n_samples = d->n_samples;
    sg.offset = d->cal_off; sg.range = d->cal_range; sg.digitisation = d->cal_dig;
    if ((size_t)n > d->read_cap) { /* n_events / event_ptr staging is shared with align_cuda */
        size_t c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0, c8 = 0;
        grow(d->seq_ptr, c1, (size_t)n); grow(d->event_ptr, c2, (size_t)n); grow(d->pair_ptr, c3, (size_t)n);
        grow(d->read_len, c4, (size_t)n); grow(d->n_events, c5, (size_t)n); grow(d->n_pairs, c6, (size_t)n);
        grow(d->scalings, c7, (size_t)n); grow(d->good, c8, (size_t)n);
        d->read_cap = c1;
    }
    if (abea_getevents(d->ctx, &sg, (core->opt.flag & F5C_RNA) ? 1 : 0, d->n_events, NULL)) die("getevents_cuda", "Cuda error", d->ctx);
    d->aligned_db = NULL;
    int64_t ne = 0;
    for (int32_t i = 0; i < n; i++) {
        if (d->n_events[i] < 0) { fprintf(stderr, "[getevents_cuda::ERROR] read %d: event table overflow\n", i); exit(EXIT_FAILURE); }
        d->event_ptr[i] = ne;
        ne += d->n_events[i];
    }
    grow(d->ev_out, d->ev_out_cap, (size_t)ne + 1);
    if (abea_getevents_download(d->ctx, d->ev_out, d->event_ptr)) die("getevents_cuda", "Cuda error", d->ctx);
    parallel_reads(n, host_threads(core), [&](int32_t i) {
        if (d->n_samples[i]) { /* convert to pA in place, src/f5c.c:692-696 */
            float* rawptr = db->sig[i]->rawptr;
            const float raw_unit = db->sig[i]->range / db->sig[i]->digitisation, offset = db->sig[i]->offset;
            for (int32_t j = 0; j < d->n_samples[i]; j++) rawptr[j] = (rawptr[j] + offset) * raw_unit;
        }
        const size_t m = (size_t)d->n_events[i];
        db->et[i].n = m;
        db->et[i].start = 0;
        db->et[i].end = m;
        db->et[i].event = NULL;
        if (m) {
            db->et[i].event = (event_t*)calloc(m, sizeof(event_t)); /* as create_events does, src/events.c:492 */
            MALLOC_CHK(db->et[i].event);
            memcpy(db->et[i].event, d->ev_out + d->event_ptr[i], m * sizeof(event_t));
        }
    });
}

/* scaling_single (src/f5c.c:736-807) for every read of the batch align_cuda has just aligned: fills
 * db->base_to_event_map[i] (malloc'd here like the reference does, freed by free_db_tmp), db->events_per_base[i],
 * db->scalings[i], db->n_event_alignment[i] and ORs the FAILED_* bits into db->read_stat_flag[i].
 * db->event_alignment[i] is left NULL: the reference frees it before scaling_single returns. */
void scaling_cuda(core_t* core, db_t* db) {
    dropin_data* d = (dropin_data*)core->cuda;
    const int32_t n = db->n_bam_rec;
    if (d->aligned_db != db) { fprintf(stderr, "[scaling_cuda::ERROR] align_cuda has not been called on this batch\n"); exit(EXIT_FAILURE); }
    double t0 = realtime();
    if (abea_scaling_stage(d->ctx, core->opt.min_num_events_to_rescale, NULL)) die("scaling_cuda", "Cuda error", d->ctx);
    grow(d->sres, d->sres_cap, (size_t)n + 1);
    grow(d->map_ptr, d->map_ptr_cap, (size_t)n + 1);
    int64_t mp = 0;
    for (int32_t i = 0; i < n; i++) {
        d->map_ptr[i] = mp;
        const int32_t K = db->read_len[i] - (int32_t)core->kmer_size + 1;
        mp += K > 0 ? K : 0;
    }
    grow(d->maps, d->map_cap, (size_t)mp + 1);
    if (abea_scaling_download(d->ctx, d->sres, d->maps, d->map_ptr)) die("scaling_cuda", "Cuda error", d->ctx);
    for (int32_t i = 0; i < n; i++) {
        const abea_scaling_result_t& r = d->sres[i];
        db->event_alignment[i] = NULL;
        db->n_event_alignment[i] = 0;
        db->events_per_base[i] = 0;
        if (db->n_event_align_pairs[i] > 0) {
            const int32_t K = db->read_len[i] - (int32_t)core->kmer_size + 1;
            db->base_to_event_map[i] = (index_pair_t*)malloc(sizeof(index_pair_t) * (size_t)K);
            MALLOC_CHK(db->base_to_event_map[i]);
            memcpy(db->base_to_event_map[i], d->maps + d->map_ptr[i], sizeof(index_pair_t) * (size_t)K);
            db->n_event_alignment[i] = r.n_event_alignment;
            db->events_per_base[i] = r.events_per_base;
            if (r.calibrated) memcpy(&db->scalings[i], &r.scalings, sizeof(scalings_t));
        } else {
            db->base_to_event_map[i] = NULL;
        }
        db->read_stat_flag[i] |= (int32_t)r.flags;
    }
    core->est_scale_time += realtime() - t0;
}

/* ---- self-test door (used by tests/test_dropin.py): builds core_t/db_t from a flat batch, calls the three entry
 * points exactly as init_core/align_db/free_core would, and hands the per-read outputs back. ------------------- */
extern "C" int f5c_dropin_selftest(const abea_batch_t* b, const abea_model_t* model, uint32_t kmer_size, int device,
                                   abea_pair_t* pairs, const int64_t* pair_ptr, int32_t* n_pairs) {
    core_t* core = (core_t*)calloc(1, sizeof(core_t));
    db_t* db = (db_t*)calloc(1, sizeof(db_t));
    core->model = (model_t*)model;
    core->kmer_size = kmer_size;
    core->opt.cuda_dev_id = device;
    core->opt.verbosity = 0;
    const int32_t n = b->n_reads;
    db->n_bam_rec = n;
    db->capacity_bam_rec = n;
    db->read = (char**)calloc(n, sizeof(char*));
    db->read_len = (int32_t*)calloc(n, sizeof(int32_t));
    db->et = (event_table*)calloc(n, sizeof(event_table));
    db->scalings = (scalings_t*)calloc(n, sizeof(scalings_t));
    db->sig = (signal_t**)calloc(n, sizeof(signal_t*));
    db->event_align_pairs = (AlignedPair**)calloc(n, sizeof(AlignedPair*));
    db->n_event_align_pairs = (int32_t*)calloc(n, sizeof(int32_t));
    for (int32_t i = 0; i < n; i++) {
        db->read[i] = (char*)(b->seq + b->seq_ptr[i]);
        db->read_len[i] = b->read_len[i];
        db->et[i].n = (size_t)b->n_events[i];
        db->et[i].end = (size_t)b->n_events[i];
        db->et[i].event = (event_t*)(b->events + b->event_ptr[i]);
        memcpy(&db->scalings[i], &b->scalings[i], sizeof(scalings_t));
        db->sig[i] = (signal_t*)calloc(1, sizeof(signal_t));
        db->sig[i]->nsample = (b->good && !b->good[i]) ? 0 : 1;
        /* event_single allocates E+L pairs per good read (src/f5c.c:724-731) */
        db->event_align_pairs[i] = db->sig[i]->nsample ? (AlignedPair*)malloc(sizeof(AlignedPair) * ((size_t)b->n_events[i] + b->read_len[i])) : NULL;
        db->sum_bases += b->read_len[i];
    }
    init_cuda(core);
    align_cuda(core, db);
    align_cuda(core, db); /* a second batch through the same core: buffers are reused */
    for (int32_t i = 0; i < n; i++) {
        n_pairs[i] = db->n_event_align_pairs[i];
        if (n_pairs[i] > 0) memcpy(pairs + pair_ptr[i], db->event_align_pairs[i], (size_t)n_pairs[i] * sizeof(AlignedPair));
        free(db->event_align_pairs[i]);
        free(db->sig[i]);
    }
    free_cuda(core);
    free(db->read); free(db->read_len); free(db->et); free(db->scalings); free(db->sig);
    free(db->event_align_pairs); free(db->n_event_align_pairs);
    free(db); free(core);
    return 0;
}

/* Second door: event_single's estimate -> align_db -> scaling_db through the drop-in, on real core_t / db_t. */
extern "C" int f5c_dropin_selftest_scaling(const abea_batch_t* b, const abea_model_t* model, uint32_t kmer_size, int device,
                                           int32_t min_num_events_to_rescale, int32_t* n_pairs,
                                           abea_scaling_result_t* results, abea_index_pair_t* maps, const int64_t* map_ptr) {
    core_t* core = (core_t*)calloc(1, sizeof(core_t));
    db_t* db = (db_t*)calloc(1, sizeof(db_t));
    core->model = (model_t*)model;
    core->kmer_size = kmer_size;
    core->opt.cuda_dev_id = device;
    core->opt.min_num_events_to_rescale = min_num_events_to_rescale;
    const int32_t n = b->n_reads;
    db->n_bam_rec = n;
    db->capacity_bam_rec = n;
    db->read = (char**)calloc(n, sizeof(char*));
    db->read_len = (int32_t*)calloc(n, sizeof(int32_t));
    db->et = (event_table*)calloc(n, sizeof(event_table));
    db->scalings = (scalings_t*)calloc(n, sizeof(scalings_t));
    db->sig = (signal_t**)calloc(n, sizeof(signal_t*));
    db->event_align_pairs = (AlignedPair**)calloc(n, sizeof(AlignedPair*));
    db->n_event_align_pairs = (int32_t*)calloc(n, sizeof(int32_t));
    db->event_alignment = (event_alignment_t**)calloc(n, sizeof(event_alignment_t*));
    db->n_event_alignment = (int32_t*)calloc(n, sizeof(int32_t));
    db->events_per_base = (double*)calloc(n, sizeof(double));
    db->base_to_event_map = (index_pair_t**)calloc(n, sizeof(index_pair_t*));
    db->read_stat_flag = (int32_t*)calloc(n, sizeof(int32_t));
    for (int32_t i = 0; i < n; i++) {
        db->read[i] = (char*)(b->seq + b->seq_ptr[i]);
        db->read_len[i] = b->read_len[i];
        db->et[i].n = (size_t)b->n_events[i];
        db->et[i].end = (size_t)b->n_events[i];
        db->et[i].event = (event_t*)(b->events + b->event_ptr[i]);
        db->sig[i] = (signal_t*)calloc(1, sizeof(signal_t));
        db->sig[i]->nsample = (b->good && !b->good[i]) ? 0 : 1;
        db->event_align_pairs[i] = db->sig[i]->nsample ? (AlignedPair*)malloc(sizeof(AlignedPair) * ((size_t)b->n_events[i] + b->read_len[i])) : NULL;
    }
    init_cuda(core);
    estimate_scalings_cuda(core, db);
    align_cuda(core, db);
    scaling_cuda(core, db);
    for (int32_t i = 0; i < n; i++) {
        n_pairs[i] = db->n_event_align_pairs[i];
        memset(&results[i], 0, sizeof(results[i]));
        memcpy(&results[i].scalings, &db->scalings[i], sizeof(scalings_t));
        results[i].events_per_base = db->events_per_base[i];
        results[i].n_event_alignment = db->n_event_alignment[i];
        results[i].flags = (uint32_t)db->read_stat_flag[i];
        if (db->base_to_event_map[i]) {
            memcpy(maps + map_ptr[i], db->base_to_event_map[i], sizeof(index_pair_t) * (size_t)(b->read_len[i] - (int32_t)kmer_size + 1));
            free(db->base_to_event_map[i]);
        }
        free(db->event_align_pairs[i]);
        free(db->sig[i]);
    }
    free_cuda(core);
    free(db->read); free(db->read_len); free(db->et); free(db->scalings); free(db->sig);
    free(db->event_align_pairs); free(db->n_event_align_pairs); free(db->event_alignment); free(db->n_event_alignment);
    free(db->events_per_base); free(db->base_to_event_map); free(db->read_stat_flag);
    free(db); free(core);
    return 0;
}

/* Third door: getevents_cuda on real core_t / db_t; hands the event tables back flat. */
extern "C" int f5c_dropin_selftest_events(const abea_signals_t* sg, int device, int rna, int32_t* n_events,
                                          abea_event_t* events, const int64_t* event_cap_ptr, float* pa_out) {
    core_t* core = (core_t*)calloc(1, sizeof(core_t));
    db_t* db = (db_t*)calloc(1, sizeof(db_t));
    abea_model_t* model = (abea_model_t*)calloc(4096, sizeof(abea_model_t)); /* any 6-mer table: not used here */
    for (int i = 0; i < 4096; i++) { model[i].level_mean = 90.f; model[i].level_stdv = 2.f; }
    core->model = (model_t*)model;
    core->kmer_size = 6;
    core->opt.cuda_dev_id = device;
    if (rna) core->opt.flag |= F5C_RNA;
    const int32_t n = sg->n_reads;
    db->n_bam_rec = n;
    db->capacity_bam_rec = n;
    db->sig = (signal_t**)calloc(n, sizeof(signal_t*));
    db->et = (event_table*)calloc(n, sizeof(event_table));
    for (int32_t i = 0; i < n; i++) {
        db->sig[i] = (signal_t*)calloc(1, sizeof(signal_t));
        db->sig[i]->nsample = (uint64_t)sg->n_samples[i];
        db->sig[i]->rawptr = (float*)malloc(sizeof(float) * (size_t)(sg->n_samples[i] + 1));
        memcpy(db->sig[i]->rawptr, sg->raw + sg->raw_ptr[i], sizeof(float) * (size_t)sg->n_samples[i]);
        db->sig[i]->offset = sg->offset[i];
        db->sig[i]->range = sg->range[i];
        db->sig[i]->digitisation = sg->digitisation[i];
    }
    init_cuda(core);
    getevents_cuda(core, db);
    for (int32_t i = 0; i < n; i++) {
        n_events[i] = (int32_t)db->et[i].n;
        if (db->et[i].n) memcpy(events + event_cap_ptr[i], db->et[i].event, db->et[i].n * sizeof(event_t));
        memcpy(pa_out + sg->raw_ptr[i], db->sig[i]->rawptr, sizeof(float) * (size_t)sg->n_samples[i]);
        free(db->et[i].event);
        free(db->sig[i]->rawptr);
        free(db->sig[i]);
    }
    free_cuda(core);
    free(db->sig); free(db->et); free(db); free(model); free(core);
    return 0;
}

/* Fourth door, runnable without a GPU: the threaded copies of the packer / unpacker on a db_t built over a flat
 * batch. seq_out / events_out receive the flattened copy (same offsets as the batch), pairs_rt the pairs copied out
 * of pairs_in into per-read buffers and back. Returns the milliseconds the two copies took. */
extern "C" double f5c_dropin_selftest_pack(const abea_batch_t* b, int threads, char* seq_out, abea_event_t* events_out,
                                           const abea_pair_t* pairs_in, const int64_t* pair_ptr, const int32_t* n_pairs,
                                           abea_pair_t* pairs_rt) {
    const int32_t n = b->n_reads;
    db_t* db = (db_t*)calloc(1, sizeof(db_t));
    db->n_bam_rec = n;
    db->read = (char**)calloc(n, sizeof(char*));
    db->read_len = (int32_t*)calloc(n, sizeof(int32_t));
    db->et = (event_table*)calloc(n, sizeof(event_table));
    db->event_align_pairs = (AlignedPair**)calloc(n, sizeof(AlignedPair*));
    db->n_event_align_pairs = (int32_t*)calloc(n, sizeof(int32_t));
    for (int32_t i = 0; i < n; i++) {
        db->read[i] = (char*)(b->seq + b->seq_ptr[i]);
        db->read_len[i] = b->read_len[i];
        db->et[i].n = (size_t)b->n_events[i];
        db->et[i].event = (event_t*)(b->events + b->event_ptr[i]);
        db->event_align_pairs[i] = (AlignedPair*)malloc(sizeof(AlignedPair) * ((size_t)b->n_events[i] + b->read_len[i] + 1));
    }
    const double t0 = realtime();
    copy_in(db, b->seq_ptr, b->event_ptr, seq_out, events_out, threads);
    copy_out(db, n_pairs, pair_ptr, pairs_in, threads);
    const double t1 = realtime();
    for (int32_t i = 0; i < n; i++) {
        if (n_pairs[i] > 0) memcpy(pairs_rt + pair_ptr[i], db->event_align_pairs[i], (size_t)n_pairs[i] * sizeof(AlignedPair));
        free(db->event_align_pairs[i]);
    }
    free(db->read); free(db->read_len); free(db->et); free(db->event_align_pairs); free(db->n_event_align_pairs); free(db);
    return (t1 - t0) * 1e3;
}

/* Fifth door: what one call of align_cuda(core, db) costs on a real, ragged db_t (every read's sequence, event table
 * and pair buffer its own allocation, as load_db / event_single leave them). The batch is built once; align_cuda is
 * then called warmup + steps times on it and the wall time of each of the last `steps` calls is returned in ms_out —
 * the `e2e_dropin` figure of bench.py. The last call's results come back flat for a parity check. */
extern "C" int f5c_dropin_bench(const abea_batch_t* b, const abea_model_t* model, uint32_t kmer_size, int device,
                                int num_thread, int warmup, int steps, double* ms_out, abea_pair_t* pairs,
                                const int64_t* pair_ptr, int32_t* n_pairs) {
    core_t* core = (core_t*)calloc(1, sizeof(core_t));
    db_t* db = (db_t*)calloc(1, sizeof(db_t));
    core->model = (model_t*)model;
    core->kmer_size = kmer_size;
    core->opt.cuda_dev_id = device;
    core->opt.num_thread = num_thread;
    const int32_t n = b->n_reads;
    db->n_bam_rec = n;
    db->capacity_bam_rec = n;
    db->read = (char**)calloc(n, sizeof(char*));
    db->read_len = (int32_t*)calloc(n, sizeof(int32_t));
    db->et = (event_table*)calloc(n, sizeof(event_table));
    db->scalings = (scalings_t*)calloc(n, sizeof(scalings_t));
    db->sig = (signal_t**)calloc(n, sizeof(signal_t*));
    db->event_align_pairs = (AlignedPair**)calloc(n, sizeof(AlignedPair*));
    db->n_event_align_pairs = (int32_t*)calloc(n, sizeof(int32_t));
    for (int32_t i = 0; i < n; i++) {
        const int32_t L = b->read_len[i], E = b->n_events[i];
        db->read[i] = (char*)malloc((size_t)L + 1);
        memcpy(db->read[i], b->seq + b->seq_ptr[i], (size_t)L);
        db->read[i][L] = 0;
        db->read_len[i] = L;
        db->et[i].n = (size_t)E;
        db->et[i].end = (size_t)E;
        db->et[i].event = (event_t*)malloc(sizeof(event_t) * (size_t)(E > 0 ? E : 1));
        if (E > 0) memcpy(db->et[i].event, b->events + b->event_ptr[i], sizeof(event_t) * (size_t)E);
        memcpy(&db->scalings[i], &b->scalings[i], sizeof(scalings_t));
        db->sig[i] = (signal_t*)calloc(1, sizeof(signal_t));
        db->sig[i]->nsample = (b->good && !b->good[i]) ? 0 : 1;
        db->event_align_pairs[i] = db->sig[i]->nsample ? (AlignedPair*)malloc(sizeof(AlignedPair) * ((size_t)E + L)) : NULL;
        db->sum_bases += L;
    }
    init_cuda(core);
    for (int s = 0; s < warmup + steps; s++) {
        for (int32_t i = 0; i < n; i++) db->n_event_align_pairs[i] = -3;
        const double t0 = realtime();
        align_cuda(core, db);
        const double t1 = realtime();
        if (s >= warmup) ms_out[s - warmup] = (t1 - t0) * 1e3;
    }
    for (int32_t i = 0; i < n; i++) {
        n_pairs[i] = db->n_event_align_pairs[i];
        if (n_pairs[i] > 0) memcpy(pairs + pair_ptr[i], db->event_align_pairs[i], (size_t)n_pairs[i] * sizeof(AlignedPair));
        free(db->event_align_pairs[i]);
        free(db->sig[i]);
        free(db->read[i]);
        free(db->et[i].event);
    }
    free_cuda(core);
    free(db->read); free(db->read_len); free(db->et); free(db->scalings); free(db->sig);
    free(db->event_align_pairs); free(db->n_event_align_pairs);
    free(db); free(core);
    return 0;
}
