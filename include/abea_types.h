/* abea_types.h — plain-C mirror of the POD types on f5c's ABEA boundary.
 *
 * Each struct is layout-identical to the reference type it names (sizes probed in SURVEY.md §8:
 * event_t 24 B, model_t 12 B with CACHED_LOG, scalings_t 16 B, AlignedPair 8 B), so a pointer to
 * the reference's array can be passed straight through the C ABI in abea_b200.h without repacking.
 * The drop-in shim (f5c_b200/csrc/f5c_dropin.cu) static_asserts the equivalence against the
 * reference's own f5c.h when it is compiled inside the f5c tree.
 */
#ifndef ABEA_TYPES_H
#define ABEA_TYPES_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference: event_t, src/f5c.h:129-136 (only .mean is read by ABEA, src/align.c:131) */
typedef struct {
    uint64_t start;
    float length;
    float mean;
    float stdv;
} abea_event_t;

/* reference: model_t with CACHED_LOG, src/f5c.h:147-155 */
typedef struct {
    float level_mean;
    float level_stdv;
    float level_log_stdv; /* log(level_stdv) evaluated on the HOST (glibc), src/model.c:179 */
} abea_model_t;

/* reference: scalings_t with CACHED_LOG, src/f5c.h:158-172 (ABEA reads .scale and .shift only) */
typedef struct {
    float scale;
    float shift;
    float var;
    float log_var;
} abea_scalings_t;

/* reference: AlignedPair, src/f5c.h:181-184. ref_pos = k-mer index, read_pos = event index */
typedef struct {
    int32_t ref_pos;
    int32_t read_pos;
} abea_pair_t;

/* reference: index_pair_t, src/f5c.h:187-190 — events [start, stop] (inclusive) mapped to one k-mer, -1/-1 = none */
typedef struct {
    int32_t start;
    int32_t stop;
} abea_index_pair_t;

/* reference read_stat_flag bits, src/f5c.h:66-68 */
#define ABEA_FAILED_CALIBRATION 0x001u
#define ABEA_FAILED_ALIGNMENT 0x002u
#define ABEA_FAILED_QUALITY_CHK 0x004u
/* reference: MIN_CALIBRATION_VAR src/f5cmisc.h:16; opt.min_num_events_to_rescale default src/f5c.c (200); the
 * events-per-base ceiling of scaling_single src/f5c.c:798 */
#define ABEA_MIN_CALIBRATION_VAR 2.5
#define ABEA_MIN_NUM_EVENTS_TO_RESCALE 200
#define ABEA_MAX_EVENTS_PER_BASE 5.0

/* What scaling_single (src/f5c.c:736-807) leaves behind for one read: the outputs of postalign
 * (src/align.c:561-660) and recalibrate_model (src/align.c:665-773) and the read_stat_flag bits it sets.
 *   scalings          db->scalings[i] after the call: recalibrated {shift, scale, var, log_var} when num_m_state >=
 *                     min_num_events_to_rescale, otherwise the input (method-of-moments) values unchanged
 *   var_d             recalibrate_model's `var` before it is narrowed to float (log_var = (float)log(var_d))
 *   events_per_base   postalign: (max_event - min_event) / n_kmers, 0 when the read has no pairs
 *   n_event_alignment entries postalign writes to its event_alignment_t list (every event of every k-mer's range)
 *   num_m_state       of those, entries with hmm_state 'M' (first event of a k-mer whose rank differs from the
 *                     previous k-mer that has events) — the rows of recalibrate_model's least-squares system
 *   flags             ABEA_FAILED_* bits OR-ed into read_stat_flag
 *   calibrated        recalibrate_model's return value
 */
typedef struct {
    abea_scalings_t scalings;
    double var_d;
    double events_per_base;
    int32_t n_event_alignment;
    int32_t num_m_state;
    uint32_t flags;
    int32_t calibrated;
} abea_scaling_result_t;

/* reference constants: src/f5c.h:30-34, src/f5cmisc.h:18 */
#define ABEA_BANDWIDTH 100
#define ABEA_MAX_KMER_SIZE 9
#define ABEA_MAX_NUM_KMER 262144
#define ABEA_AVG_EVENTS_PER_KMER_MAX 15.0f

/* reference model ids: src/f5cmisc.h:24-30 */
#define ABEA_MODEL_ID_DNA_R9 1
#define ABEA_MODEL_ID_RNA_R9 3
#define ABEA_MODEL_ID_DNA_R10 4
#define ABEA_MODEL_ID_RNA_RNA004 6

/* A batch of raw nanopore signals in flat form, the input of event detection (reference: signal_t / db->sig[i],
 * src/f5c.h:262-287, as event_single reads it, src/f5c.c:684-696).
 *   raw          : concatenated samples as float — the ADC counts the reader hands over (slow5/fast5 int16 widened
 *                  to float), or picoamperes when offset == NULL
 *   raw_ptr      : first sample of read i; n_samples: its length
 *   offset, range, digitisation : per read; pA = (raw + offset) * (range / digitisation), all in float
 */
typedef struct {
    int32_t n_reads;
    const float* raw;
    const int64_t* raw_ptr;
    const int32_t* n_samples;
    const float* offset;
    const float* range;
    const float* digitisation;
    const int16_t* raw_i16; /* the ADC counts as the file holds them, used when raw == NULL: 2 bytes per sample cross PCIe
                             * and the widening to float (src/f5cio.c:461) is done on the device */
} abea_signals_t;

/* BLOW5 records as they lie in the file (slow5lib v1.x binary format, slow5lib/src/slow5.c): the input of the
 * device-side record decode. A host reader only walks the framing (uint64 size + payload per record).
 *   bytes         : the records' payloads back to back (what follows each record's size field)
 *   rec_ptr/len   : first byte and stored size of record i
 *   record_method : 0 none, 1 zlib (the record-compression byte of the file header; zstd = 2 is not supported)
 *   signal_method : 0 none, 1 svb-zd, 2 ex-zd (the signal-compression byte, files >= v0.2.0) */
typedef struct {
    int32_t n_reads;
    const uint8_t* bytes;
    const int64_t* rec_ptr;
    const int32_t* rec_len;
    int32_t record_method;
    int32_t signal_method;
} abea_blow5_t;

/* A ragged batch of reads in flat (CSR-style) form: what the reference's align_cuda packs db_t into
 * (src/f5c.cu:744-800) and what every implementation in this repo (CUDA path, oracle, _ref shim)
 * consumes. All pointers are host pointers unless a function says otherwise.
 *
 *   seq      : concatenated read sequences, read i at seq[seq_ptr[i] .. seq_ptr[i]+read_len[i]) followed
 *              by one NUL (so seq_ptr advances by read_len+1, like the reference's read_ptr)
 *   events   : concatenated event tables, read i at events[event_ptr[i] .. +n_events[i]); NULL (with event_means NULL
 *              as well) = the event tables the last abea_getevents left on the device (same reads, same order; n_events must repeat its counts,
 *              event_ptr is ignored) — raw signal in, alignment out, no event table crosses PCIe
 *   scalings : per read (from estimate_scalings_using_mom); NULL = estimate them on the device (abea_estimate_scalings)
 *   good     : per read, non-zero iff db->sig[i]->nsample > 0 (src/f5c.c:811); NULL = all good
 *   event_means : optional flat array of the events' means alone, read i at event_means[event_ptr[i] .. +n_events[i]).
 *              ABEA and the stages either side of it read nothing of an event but .mean (reference src/align.c:131,
 *              src/align.cu:415), so a caller that holds (or extracts) the means ships 4 bytes per event instead of
 *              the 24 of event_t. When it is not NULL it is used and `events` is ignored (and may be NULL).
 */
typedef struct {
    int32_t n_reads;
    const char* seq;
    const int64_t* seq_ptr;
    const int32_t* read_len;
    const abea_event_t* events;
    const int64_t* event_ptr;
    const int32_t* n_events;
    const abea_scalings_t* scalings;
    const uint8_t* good;
    const float* event_means;
} abea_batch_t;

/* The same batch in the RAGGED form db_t holds it in (reference src/f5c.h:290-352): one pointer per read.
 *   seq[i]     db->read[i] (read_len[i] bases)          events[i]  db->et[i].event (n_events[i] entries)
 *   pairs[i]   db->event_align_pairs[i], caller-allocated with n_events[i] + read_len[i] entries (src/f5c.c:724-726),
 *              may be NULL for reads that are not good
 *   n_pairs    db->n_event_align_pairs, written for every read */
typedef struct {
    int32_t n_reads;
    const char* const* seq;
    const int32_t* read_len;
    const abea_event_t* const* events;
    const int32_t* n_events;
    const abea_scalings_t* scalings;
    const uint8_t* good;
    abea_pair_t* const* pairs;
    int32_t* n_pairs;
} abea_ragged_t;

/* One word of a pair list's path codes (abea_device_codes in abea_b200.h): the first word of a read holds its first
 * pair, every further word 32 steps as two bit planes. */
typedef struct {
    uint32_t a, b;
} abea_code_word_t;

#ifdef __cplusplus
}
#endif
#endif
