/* abea_b200.h — C ABI of libabea_b200.so: f5c's adaptive banded event alignment (ABEA) on B200 (sm_100a).
 *
 * This is the boundary a host program binds. Plain pointers and sizes only; no C++ or torch types.
 * Each entry point names the reference interface it replaces (paths relative to the f5c tree):
 *
 *   abea_create / abea_destroy      init_cuda / free_cuda            src/f5c.h:575-581, src/f5c.cu:23-234
 *   abea_set_model                  the model H2D inside init_cuda   src/f5c.cu:96-103 (core->model, src/f5c.c:286-337)
 *   abea_align_batch                align_cuda(core_t*, db_t*)       src/f5cmisc.h:122-125, src/f5c.cu:647-1061
 *                                   == the GPU branch of align_db    src/f5c.c:833-845
 *   abea_upload_batch / abea_run /  the three phases of align_cuda   src/f5c.cu:744-899 (pack + H2D),
 *   abea_download                   kept separable for measurement   :910-960 (kernels), :979-1030 (D2H + unpack)
 *   abea_model_fill_log_stdv        set_model's CACHED_LOG fill      src/model.c:179
 *   abea_getevents_blow5            read_slow5_single's slow5lib     src/f5cio.c:421-470; slow5lib/src/slow5_press.c:921-1010,
 *                                   calls + event_single's first half 1118-1170, 1262-1842; slow5.c:2840-2930
 *   abea_getevents /                getevents (event detection) per  src/events.c:562-582, called by event_single
 *   abea_getevents_download         read + the pA conversion         src/f5c.c:692-703
 *   abea_estimate_scalings          estimate_scalings_using_mom      src/align.c:58-106, called per read by event_single
 *                                   (+ the RNA event reversal)       src/f5c.c:709-721
 *   abea_scaling_stage /            scaling_db = scaling_single per  src/f5c.c:736-807: postalign src/align.c:561-660,
 *   abea_scaling_download           read                             recalibrate_model src/align.c:665-773, read flags
 *
 * The drop-in with the reference's own C++-linkage symbols (align_cuda/init_cuda/free_cuda over core_t/db_t) is
 * f5c_b200/csrc/f5c_dropin.cu, a thin packer over this ABI that is compiled inside the f5c tree.
 *
 * Semantics of abea_align_batch (identical to the reference CPU path, align_single src/f5c.c:811-830 + align
 * src/align.c:180-559): for every read i, n_pairs[i] is the number of aligned pairs after QC (0 for bad reads,
 * over-segmented reads with events/base >= 15, and QC failures) and pairs[pair_ptr[i] .. +n_pairs[i]) holds them in
 * ascending order, bit-identical to the reference. The caller sizes read i's region to n_events[i]+read_len[i]
 * pairs (src/f5c.c:724-726); a region's slots past n_pairs[i] are left as they were. No read is ever sent to a CPU fallback (the reference's if_on_gpu heuristic,
 * src/f5c.cu:440-452, has no counterpart).
 *
 * Errors: functions return 0 on success or a negative abea_status; abea_last_error() gives the message. The
 * reference convention (print and exit(-1), src/f5cmisc.cuh:54-97) is applied by the drop-in shim, not here.
 * Threading: one context per host thread / GPU; a context is not re-entrant.
 */
#ifndef ABEA_B200_H
#define ABEA_B200_H

#include "abea_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct abea_ctx abea_ctx_t;

enum abea_status {
    ABEA_OK = 0,
    ABEA_ERR_CUDA = -1,     /* a CUDA runtime call failed */
    ABEA_ERR_ARG = -2,      /* invalid argument */
    ABEA_ERR_NOMODEL = -3,  /* abea_set_model not called */
    ABEA_ERR_NODEVICE = -4, /* no usable CUDA device (the library has no CPU path) */
    ABEA_ERR_STATE = -5     /* call order violated (e.g. abea_run before abea_upload_batch) */
};

/* Device-side time of each phase in milliseconds (CUDA events on the context's stream), plus host-side packing.
 * Mirrors the reference's timer split (core->align_cuda_preprocess/memcpy/kernel/postprocess, src/f5c.h:457-466). */
typedef struct {
    double pack_ms;        /* host: descriptors, scheduling order */
    double h2d_ms;         /* device: sequence + event + descriptor copies */
    double kmer_ms;        /* device: abea_prepare_kernel (k-mer parameter cache + input validation) */
    double fill_ms;        /* device: band fill + fused traceback / QC (narrow + wide kernels, concurrent) */
    double trace_ms;       /* device: always ~0 — traceback + QC are fused into the fill kernels (kept for the reference's timer split) */
    double kernel_ms;      /* device: first kernel start to last kernel end */
    double d2h_ms;         /* device: result copies */
    double unpack_ms;      /* host: scatter into the caller's buffers */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int32_t kernel_launches; /* kernels of this library launched by the call */
    int32_t n_scheduled;     /* reads that passed the eligibility filter */
    int32_t n_wide;          /* of those, reads filled by the wide (4 warps per read) kernel */
    int32_t streamed;        /* bit 0: events streamed in by abea_load_kernel; bit 1: pair lists written whole to the caller's mapped
                              * buffer; bit 2: pair lists out as path codes expanded by host threads */
    int64_t n_bands;         /* sum of NB over scheduled reads */
    int64_t n_events;        /* sum of E over scheduled reads (the metric's numerator) */
    double load_ms;          /* device: abea_load_kernel, first CTA start to last piece landed (streaming only) */
    double mom_ms;           /* device: abea_mom_kernel (abea_estimate_scalings) */
    double scaling_ms;       /* device: abea_scaling_kernel (abea_scaling_stage) */
    double events_ms;        /* device: abea_events_kernel (abea_getevents) */
    int64_t n_samples;       /* raw samples of the last abea_getevents */
    double ragged_ms;        /* host: wall time of the last abea_align_ragged call, everything included */
    double blow5_ms;         /* device: BLOW5 record inflate + parse + signal decode (abea_getevents_blow5) */
} abea_timing_t;

/* Create a context on CUDA device `device` (cudaSetDevice is applied on every call). */
int abea_create(abea_ctx_t** ctx, int device);
void abea_destroy(abea_ctx_t* ctx);
const char* abea_last_error(const abea_ctx_t* ctx);

/* Upload the pore-model table: 4^kmer_size entries with level_log_stdv already filled. */
int abea_set_model(abea_ctx_t* ctx, const abea_model_t* model, uint32_t kmer_size);

/* level_log_stdv = logf(level_stdv) on the host, exactly as the reference's set_model/read_model do. */
void abea_model_fill_log_stdv(abea_model_t* model, int64_t n);

/* The whole path with HOST buffers in and out (pack + H2D + kernels + D2H + unpack). timing may be NULL. */
int abea_align_batch(abea_ctx_t* ctx, const abea_batch_t* batch, abea_pair_t* pairs, const int64_t* pair_ptr,
                     int32_t* n_pairs, abea_timing_t* timing);

/* The whole path on the caller's RAGGED per-read arrays — the form db_t holds a batch in (db->read[i], db->et[i].event,
 * db->event_align_pairs[i]; reference src/f5c.h:290-352) — so that the flattening the reference's align_cuda does on
 * its calling thread before and after its kernels (src/f5c.cu:744-800, 1005-1030) overlaps the kernels instead:
 * `threads` host threads extract the event means piece by piece into pinned staging in the order the loader kernel is
 * going to ship them, and expand every read's pair list from its path codes (8 bytes per 32 pairs over PCIe, see
 * abea_host_threads) into pairs[i] as soon as its count appears in the pinned count array (the traceback publishes it
 * behind a system-scope fence). Same results as abea_align_batch. */
int abea_align_ragged(abea_ctx_t* ctx, const abea_ragged_t* batch, int threads, abea_timing_t* timing);

/* The same path in three separable phases. abea_run may be repeated on a resident batch (it re-zeroes its queues). */
int abea_upload_batch(abea_ctx_t* ctx, const abea_batch_t* batch, abea_timing_t* timing);
int abea_run(abea_ctx_t* ctx, abea_timing_t* timing);
int abea_download(abea_ctx_t* ctx, abea_pair_t* pairs, const int64_t* pair_ptr, int32_t* n_pairs,
                  abea_timing_t* timing);

/* ---- event detection; bit-identical to the reference's CPU code ----
 *
 * abea_getevents: getevents(nsample, rawptr, rna) (src/events.c:562-582) for every read of a batch of raw signals,
 * including event_single's conversion to picoamperes when the calibration arrays are given. n_events_out[i]
 * receives the number of events of read i (0 for signals shorter than 100 samples; the reference itself aborts on
 * any signal under 200 samples — nchunk == 1 in trim_raw_by_mad — and on zero-MAD signals, where this library
 * returns the events the detector finds; -1 if a signal produced more than n_samples/2 + 1 boundaries, which no real signal does). The event tables stay
 * on the device until abea_getevents_download copies them to events[event_ptr[i] ..] — the caller sizes and lays
 * out that array from the counts (e.g. a prefix sum), which is the abea_batch_t.events / event_ptr of the alignment.
 * rna != 0 selects the RNA detector parameters (src/events.c:59-63).
 * A following abea_upload_batch whose batch->events is NULL aligns those device-resident tables directly (they stay
 * valid until the next abea_getevents). */
int abea_getevents(abea_ctx_t* ctx, const abea_signals_t* signals, int rna, int32_t* n_events_out, abea_timing_t* timing);
int abea_getevents_download(abea_ctx_t* ctx, abea_event_t* events, const int64_t* event_ptr);

/* ---- BLOW5 records decoded on the device ----
 * abea_getevents_blow5: abea_getevents for reads that are still BLOW5 records — the file's own bytes cross PCIe and
 * slow5lib's reader side runs on the GPU: record decompression (zlib inflate, slow5lib/src/slow5_press.c:921-1010),
 * record parsing (slow5.c:2840-2930), signal decompression (svb-zd, slow5_press.c:1118-1170; ex-zd, :1262-1842), the widening to float
 * and read_slow5_single's narrowing of the calibration to float (src/f5cio.c:455-461). n_samples_out (may be NULL)
 * receives each read's sample count. Everything after that is abea_getevents: the event tables stay on the device for
 * abea_getevents_download / abea_upload_batch(events == NULL). zstd records return ABEA_ERR_ARG.
 * abea_raw_download: the float samples of the last abea_getevents / abea_getevents_blow5, for tests of the decoders. */
int abea_getevents_blow5(abea_ctx_t* ctx, const abea_blow5_t* records, int rna, int32_t* n_events_out,
                         int32_t* n_samples_out, abea_timing_t* timing);
int abea_raw_download(abea_ctx_t* ctx, float* raw, const int64_t* raw_ptr);

/* ---- the stages either side of the alignment; bit-identical to the reference's CPU code ----
 *
 * abea_estimate_scalings: estimate_scalings_using_mom (src/align.c:58-106) for every read of the RESIDENT batch
 * (abea_upload_batch; batch->scalings may then be NULL) that the reference would have run event_single on (good,
 * at least one event, at least k bases). The estimates replace the batch's scalings on the device — the next
 * abea_run aligns with them — and are copied to scalings_out (caller's order, var = log_var = 0) when it is not
 * NULL. reverse_events != 0 reverses every read's event array in place afterwards, as event_single does for RNA
 * (src/f5c.c:713-721): upload RNA events in signal order and let this call turn them 3'->5'.
 * abea_align_batch on a batch without scalings runs this stage itself (copy-engine path, no streaming). */
int abea_estimate_scalings(abea_ctx_t* ctx, int reverse_events, abea_scalings_t* scalings_out, abea_timing_t* timing);

/* abea_scaling_stage: scaling_single (src/f5c.c:736-807) for every read, on the pair lists the last abea_run left
 * on the device: the k-mer -> event-range map and events_per_base of postalign, the recalibrated shift / scale / var
 * of recalibrate_model (when at least min_num_events_to_rescale 'M' rows exist; the reference default is
 * ABEA_MIN_NUM_EVENTS_TO_RESCALE), and the ABEA_FAILED_* flags. Results stay on the device. */
int abea_scaling_stage(abea_ctx_t* ctx, int32_t min_num_events_to_rescale, abea_timing_t* timing);

/* Copy the stage's results to the host: results[n_reads] in the caller's order (log_var filled here with the host's
 * log(), as the reference does), and — when maps is not NULL — read i's base_to_event_map (K_i = read_len-k+1
 * entries) at maps[map_ptr[i]]. Only reads with n_event_alignment > 0 have a map (the reference allocates none for
 * the others, src/f5c.c:787). */
int abea_scaling_download(abea_ctx_t* ctx, abea_scaling_result_t* results, abea_index_pair_t* maps,
                          const int64_t* map_ptr);

/* Device-resident results of the stage for consumers that stay on the GPU: results in the caller's order, maps in
 * the canonical layout (read i at the prefix sum of max(K, 0)). Valid until the next upload / destroy. */
int abea_scaling_device_results(abea_ctx_t* ctx, const abea_scaling_result_t** d_results,
                                const abea_index_pair_t** d_maps, int64_t* total_map_entries);

/* Per-read diagnostics of the last run, indexed like the batch (any pointer may be NULL):
 * sum of emissions along the traceback (the quantity in the reference's adaptive.exp debug dumps), pairs before
 * QC, the event the traceback started from and the longest skip run. Reads that were not scheduled report 0. */
int abea_read_stats(abea_ctx_t* ctx, double* sum_emission, int32_t* n_aligned, int32_t* end_event,
                    int32_t* max_gap);

/* Per-read latency of the last run in SM clock cycles (band fill; traceback + QC) and whether the wide kernel took
 * the read; indexed like the batch, any pointer may be NULL. A profiling aid: it is how profiles/ shows what the
 * longest reads cost. */
int abea_read_cycles(abea_ctx_t* ctx, int64_t* fill_cycles, int64_t* trace_cycles, int32_t* wide);
/* Segment-parallel traceback: per read, how many of its (up to 32) segments were entered in another cell than the
 * speculative walk had assumed and were walked again. A profiling aid (the choice of the margin, ABEA_TB_MARGIN). */
int abea_read_respec(abea_ctx_t* ctx, int32_t* respec);
/* When the fill of each read began: %globaltimer in microseconds (low 31 bits), -1 for reads that were not scheduled.
 * A profiling aid (how far the streaming loader is ahead of the fill). */
int abea_read_starts(abea_ctx_t* ctx, int32_t* start_us);

/* The scheduler's model of the kernels: cycles per band of a wide CTA, of a narrow warp sharing its sub-partition and
 * of a narrow warp alone on it, and cycles per traceback step. Starts from values measured on B200 and is re-derived
 * from the per-read cycle counts of resident runs (ABEA_CALIBRATE=0 keeps the starting values). */
int abea_scheduler_model(abea_ctx_t* ctx, double* cycles4);

/* Device-resident results of the last abea_run, for consumers that stay on the GPU (e.g. the NCCL gather of a
 * multi-GPU driver): *d_pairs points at the pairs in the canonical capacity layout (read i of the batch at the prefix
 * sum of n_events+read_len, *total_pairs_capacity entries in all), *d_n_pairs at n_reads int32 counts in batch order.
 * The pointers stay valid until the next abea_upload_batch / abea_destroy on this context. */
int abea_device_results(abea_ctx_t* ctx, const abea_pair_t** d_pairs, const int32_t** d_n_pairs,
                        int64_t* total_pairs_capacity, int32_t* n_reads);

/* The pair lists of the last run packed back to back in DEVICE memory the caller provides (e.g. the send buffer of an
 * NCCL exchange): d_dst[0 .. *total_pairs) = read 0's pairs, read 1's pairs, ... in batch order; the per-read counts
 * are *d_n_pairs of abea_device_results. dst_capacity is in pairs (the batch's sum of n_events+read_len always
 * suffices). */
int abea_compact_results(abea_ctx_t* ctx, abea_pair_t* d_dst, int64_t dst_capacity, int64_t* total_pairs);

/* The same results as PATH CODES, 1/32 of the size — what a multi-GPU driver should move between GPUs. A pair list is a
 * monotone lattice path, so its first pair and one bit per coordinate per step describe it: read i (batch order) owns the
 * words [(P_i >> 5) + 2 i, ...) of *d_codes, P_i = the prefix sum of n_events+read_len (the capacity layout of
 * abea_device_results); word 0 = {first ref_pos, first read_pos}, word 1 + j = steps 32 j .. 32 j + 31 as two bit planes
 * (bit t of `a`: ref_pos advances at that step; bit t of `b`: read_pos does). Only the first 1 + ceil((n_pairs[i] - 1) / 32)
 * words of a read with n_pairs[i] > 0 are defined. *n_words = the buffer's size, a function of the batch shape alone
 * ((sum of n_events+read_len >> 5) + 2 n_reads + 2), so a peer can be sent the whole buffer without a size exchange.
 * The codes are made by the first call after a run (one kernel over the pair lists, one warp per read; the call
 * synchronises the context's stream). Valid until the next abea_run / abea_upload_batch / abea_destroy on this context. */
int abea_device_codes(abea_ctx_t* ctx, const abea_code_word_t** d_codes, int64_t* n_words);

/* Path codes back to dense pair lists, on this context's device: d_codes / d_n_pairs (device memory, e.g. received from
 * a peer) describe n_reads reads whose capacity prefix sums are cap_ptr[0 .. n_reads] (HOST memory); d_dst[0 .. total)
 * receives read 0's pairs, read 1's pairs, ... back to back. dst_capacity (in pairs) must be at least cap_ptr[n_reads].
 * The total is written to *d_total (device memory, may be NULL) and, if total_pairs is not NULL, returned to the host
 * (which synchronises the context's stream; with NULL the call only enqueues work). */
int abea_expand_codes(abea_ctx_t* ctx, const abea_code_word_t* d_codes, const int32_t* d_n_pairs, const int64_t* cap_ptr,
                      int32_t n_reads, abea_pair_t* d_dst, int64_t dst_capacity, int64_t* d_total, int64_t* total_pairs);

/* The reference's --print-banded-aln dump (src/f5c.c:989-1006) of a batch's pair lists, byte for byte the text f5c
 * prints, so that this path can be diffed against an f5c run: reads whose read_stat_flag has ABEA_FAILED_ALIGNMENT are
 * skipped (read_stat_flag may be NULL: none is). path "-" = stdout. Host-side formatting only. */
int abea_write_pairs(const char* path, int append, int32_t n_reads, const char* const* names, const int32_t* n_pairs,
                     const abea_pair_t* pairs, const int64_t* pair_ptr, const uint32_t* read_stat_flag);

/* f5c resquiggle's output (src/resquiggle.c:322-447) for a batch that has been through the whole chain — event detection,
 * alignment, abea_scaling_stage — byte for byte the text f5c prints: fmt 0 = TSV rows "read_id kmer_idx start_raw_idx
 * end_raw_idx" (one per k-mer, "." for k-mers without events; the header line is written when header != 0), fmt 1 = PAF
 * with the ss:Z: tag of per-base sample counts. Reads whose flags are not 0 are skipped, as the reference does. For RNA
 * (rna != 0) the k-mer map is reversed first, as output_db_rsq does. Inputs: names, read_len, n_samples (db->sig[i]->nsample)
 * per read; the event tables (events / event_ptr: only start and length are read); results / maps / map_ptr as
 * abea_scaling_download returns them. The reference prints the FIRST read's scale and shift in every PAF line
 * (db->scalings->scale, src/resquiggle.c:442-443); so does this. path "-" = stdout. Host-side formatting only. */
int abea_write_resquiggle(const char* path, int append, int fmt, int header, int rna, uint32_t kmer_size, int32_t n_reads,
                          const char* const* names, const int32_t* read_len, const int64_t* n_samples,
                          const abea_event_t* events, const int64_t* event_ptr, const abea_scaling_result_t* results,
                          const abea_index_pair_t* maps, const int64_t* map_ptr);

/* Host threads abea_align_batch uses to expand the pair lists, which leave the device as path codes (the first pair of
 * a list and two bits per step: 8 bytes per 32 pairs over PCIe instead of 256) while the kernels are still running.
 * threads >= 0 sets the number (0: no path codes — the lists are written whole into a pinned caller buffer, or copied
 * back by the copy engine); threads < 0 only queries. Returns the value in force. Default: min(8, CPUs of the calling
 * process); ABEA_HOST_THREADS overrides the default. abea_align_ragged has its own `threads` argument. */
int abea_host_threads(abea_ctx_t* ctx, int threads);

/* Pinned host memory for callers that want the H2D/D2H copies to run at full PCIe rate. */
void* abea_host_alloc(size_t bytes);
void abea_host_free(void* p);

/* Device properties of the context's GPU (for reports): SM count and name (buffer of >= 256 bytes). */
int abea_device_info(abea_ctx_t* ctx, int* sm_count, char* name);

/* Library version string. */
const char* abea_version(void);

#ifdef __cplusplus
}
#endif
#endif
