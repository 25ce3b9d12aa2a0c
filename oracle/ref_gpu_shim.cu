/* ref_gpu_shim.cu — TEST / MEASUREMENT INFRASTRUCTURE. extern "C" doorway into the reference's own CUDA kernels.
 *
 * This file is ours; the kernels it launches are the reference's: /root/reference/src/align.cu is compiled where it
 * lies, for sm_100a, with the reference's flags (Makefile:43-51: -O2 -std=c++11, nvcc's default -fmad=true) into
 * oracle/_ref/align_cu.o (see oracle/Makefile, target refgpu), and linked with this shim into
 * oracle/_ref/libf5c_refgpu.so. The shim does what the reference's align_cuda does around them (src/f5c.cu:744-1030):
 * the flat "SoA" layout (read / read_ptr / read_len, event_table / event_ptr / n_events, scalings), device arrays sized
 * as init_cuda sizes them per batch (bands 400 B + trace 100 B + lower-left 8 B per band, src/f5c.cu:121-146), the
 * cudaMemset of the trace (:832), seven blocking H2D copies (:875-899), the three launches with the reference's grid
 * and block shapes and a device synchronisation after each (:910-960), two D2H copies (:979-985) and the reversal of
 * the pair lists on the host (REVERSAL_ON_CPU, :1011-1020).
 *
 * What it does NOT do: the reference diverts reads longer than 3x the batch mean or with >= 5 events per base to CPU
 * threads (if_on_gpu, src/f5c.cu:440-452). Here EVERY eligible read goes through the kernels — the comparison is
 * kernel against kernel, on the same box — which is the reference's own "GPU only" variant (src/f5c_gpuonly.cu).
 *
 * It is the same-box GPU baseline of bench.py ("gpu_reference"). Its pairs are NOT expected to be bit-equal to the
 * CPU align(): the reference kernels compute the transition constants in float with logf/expf and let nvcc contract
 * multiply-adds; bench.py reports the fraction of reads whose pair lists agree.
 * Only tests/ and bench.py may load it.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "f5c.h"
#include "f5cmisc.h"
#include "f5cmisc.cuh"

#include "../include/abea_types.h"

static_assert(sizeof(abea_event_t) == sizeof(event_t), "event_t layout");
static_assert(sizeof(abea_model_t) == sizeof(model_t), "model_t layout");
static_assert(sizeof(abea_scalings_t) == sizeof(scalings_t), "scalings_t layout");
static_assert(sizeof(abea_pair_t) == sizeof(AlignedPair), "AlignedPair layout");

#define RG_CHK(call)                                                                                        \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            fprintf(stderr, "[ref_gpu_shim] %s failed: %s (line %d)\n", #call, cudaGetErrorString(e_), __LINE__); \
            return -1;                                                                                      \
        }                                                                                                   \
    } while (0)

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + 1e-6 * ts.tv_nsec;
}

extern "C" {

/* Bytes of device memory the reference layout needs for this batch (so a caller can skip a batch that does not fit). */
int64_t f5cref_gpu_bytes(const abea_batch_t* b) {
    int64_t sum_read_len = 0, sum_n_events = 0;
    for (int32_t i = 0; i < b->n_reads; i++) {
        sum_read_len += b->read_len[i] + 1;
        sum_n_events += b->n_events[i];
    }
    const int64_t bands = sum_read_len + sum_n_events;
    return bands * (ALN_BANDWIDTH * 5 + 8) + sum_n_events * (24 + 16) + sum_read_len * 13;
}

/* The reference's three kernels over a whole batch, warmup + steps times. Reads that fail align_single's filter
 * (src/f5c.c:811-814) are left out, as the reference's packer leaves them out (src/f5c.cu:709-726).
 * kernel_ms[s]: CUDA-event time from the first to the last kernel of step s (pre + core + post, as
 * core->align_kernel_time sums them); e2e_ms[s]: wall time of memset + H2D + kernels + D2H + host reversal.
 * pairs / n_pairs receive the last step's lists in the caller's capacity layout (pair_ptr). */
int f5cref_gpu_align_batch(const abea_batch_t* b, const abea_model_t* model_host, uint32_t kmer_size, int device,
                           abea_pair_t* pairs, const int64_t* pair_ptr, int32_t* n_pairs, int warmup, int steps,
                           double* kernel_ms, double* e2e_ms) {
    RG_CHK(cudaSetDevice(device));
    const int32_t n_all = b->n_reads;
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_all + 1));
    int32_t n = 0;
    for (int32_t i = 0; i < n_all; i++) {
        n_pairs[i] = 0;
        const bool good = b->good ? b->good[i] != 0 : true;
        if (good && b->n_events[i] > 0 && b->read_len[i] >= (int32_t)kmer_size &&
            (b->n_events[i] / (float)b->read_len[i]) < AVG_EVENTS_PER_KMER_MAX)
            idx[n++] = i;
    }
    /* flat host arrays in the reference's layout */
    ptr_t* read_ptr_host = (ptr_t*)malloc(sizeof(ptr_t) * (size_t)(n + 1));
    ptr_t* event_ptr_host = (ptr_t*)malloc(sizeof(ptr_t) * (size_t)(n + 1));
    int32_t* read_len_host = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t* n_events_host = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
    scalings_t* scalings_host = (scalings_t*)malloc(sizeof(scalings_t) * (size_t)(n + 1));
    int32_t* n_event_align_pairs_host = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int64_t sum_read_len = 0, sum_n_events = 0;
    for (int32_t j = 0; j < n; j++) {
        const int32_t i = idx[j];
        read_ptr_host[j] = sum_read_len;
        event_ptr_host[j] = sum_n_events;
        read_len_host[j] = b->read_len[i];
        n_events_host[j] = b->n_events[i];
        memcpy(&scalings_host[j], &b->scalings[i], sizeof(scalings_t));
        sum_read_len += b->read_len[i] + 1;
        sum_n_events += b->n_events[i];
    }
    char* read_host = (char*)malloc((size_t)sum_read_len + 1);
    event_t* event_table_host = (event_t*)malloc(sizeof(event_t) * (size_t)(sum_n_events + 1));
    AlignedPair* event_align_pairs_host = (AlignedPair*)malloc(sizeof(AlignedPair) * 2 * (size_t)(sum_n_events + 1));
    for (int32_t j = 0; j < n; j++) {
        const int32_t i = idx[j];
        memcpy(read_host + read_ptr_host[j], b->seq + b->seq_ptr[i], (size_t)b->read_len[i]);
        read_host[read_ptr_host[j] + b->read_len[i]] = 0;
        memcpy(event_table_host + event_ptr_host[j], b->events + b->event_ptr[i], sizeof(event_t) * (size_t)b->n_events[i]);
    }
    /* device arrays (src/f5c.cu:62-146) */
    char* read; ptr_t *read_ptr, *event_ptr; int32_t *read_len, *n_events, *n_event_align_pairs;
    event_t* event_table; scalings_t* scalings; model_t *model, *model_kmer_cache; AlignedPair* event_align_pairs;
    float* bands; uint8_t* trace; EventKmerPair* band_lower_left;
    const size_t sum_n_bands = (size_t)(sum_n_events + sum_read_len);
    const size_t n_model = (size_t)1 << (2 * kmer_size);
    RG_CHK(cudaMalloc((void**)&read, (size_t)sum_read_len + 1));
    RG_CHK(cudaMalloc((void**)&read_ptr, sizeof(ptr_t) * (size_t)(n + 1)));
    RG_CHK(cudaMalloc((void**)&event_ptr, sizeof(ptr_t) * (size_t)(n + 1)));
    RG_CHK(cudaMalloc((void**)&read_len, sizeof(int32_t) * (size_t)(n + 1)));
    RG_CHK(cudaMalloc((void**)&n_events, sizeof(int32_t) * (size_t)(n + 1)));
    RG_CHK(cudaMalloc((void**)&n_event_align_pairs, sizeof(int32_t) * (size_t)(n + 1)));
    RG_CHK(cudaMalloc((void**)&event_table, sizeof(event_t) * (size_t)(sum_n_events + 1)));
    RG_CHK(cudaMalloc((void**)&scalings, sizeof(scalings_t) * (size_t)(n + 1)));
    RG_CHK(cudaMalloc((void**)&model, sizeof(model_t) * n_model));
    RG_CHK(cudaMalloc((void**)&model_kmer_cache, sizeof(model_t) * (size_t)(sum_read_len + 1)));
    RG_CHK(cudaMalloc((void**)&event_align_pairs, sizeof(AlignedPair) * 2 * (size_t)(sum_n_events + 1)));
    RG_CHK(cudaMalloc((void**)&bands, sizeof(float) * sum_n_bands * ALN_BANDWIDTH));
    RG_CHK(cudaMalloc((void**)&trace, sizeof(uint8_t) * sum_n_bands * ALN_BANDWIDTH));
    RG_CHK(cudaMalloc((void**)&band_lower_left, sizeof(EventKmerPair) * sum_n_bands));
    RG_CHK(cudaMemcpy(model, model_host, sizeof(model_t) * n_model, cudaMemcpyHostToDevice)); /* init_cuda, src/f5c.cu:96-103 */
    cudaEvent_t ev0, ev1;
    RG_CHK(cudaEventCreate(&ev0));
    RG_CHK(cudaEventCreate(&ev1));

    for (int s = 0; s < warmup + steps; s++) {
        const double t0 = now_ms();
        RG_CHK(cudaMemset(trace, 0, sizeof(uint8_t) * sum_n_bands * ALN_BANDWIDTH));                       /* :832 */
        RG_CHK(cudaMemcpy(read_ptr, read_ptr_host, (size_t)n * sizeof(ptr_t), cudaMemcpyHostToDevice));   /* :875-899 */
        RG_CHK(cudaMemcpy(read, read_host, (size_t)sum_read_len, cudaMemcpyHostToDevice));
        RG_CHK(cudaMemcpy(read_len, read_len_host, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
        RG_CHK(cudaMemcpy(n_events, n_events_host, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
        RG_CHK(cudaMemcpy(event_ptr, event_ptr_host, (size_t)n * sizeof(ptr_t), cudaMemcpyHostToDevice));
        RG_CHK(cudaMemcpy(event_table, event_table_host, sizeof(event_t) * (size_t)sum_n_events, cudaMemcpyHostToDevice));
        RG_CHK(cudaMemcpy(scalings, scalings_host, sizeof(scalings_t) * (size_t)n, cudaMemcpyHostToDevice));
        RG_CHK(cudaEventRecord(ev0, 0));
        if (n > 0) {
            dim3 grid1(1, (n + BLOCK_LEN_READS - 1) / BLOCK_LEN_READS);                                    /* :910-933 */
            dim3 block1(BLOCK_LEN_BANDWIDTH, BLOCK_LEN_READS);
            align_kernel_pre_2d<<<grid1, block1>>>(read, read_len, read_ptr, n_events, event_ptr, model, kmer_size, n,
                                                   model_kmer_cache, bands, trace, band_lower_left);
            RG_CHK(cudaDeviceSynchronize());
            align_kernel_core_2d_shm<<<grid1, block1>>>(read_len, read_ptr, event_table, n_events, event_ptr, scalings, n,
                                                        model_kmer_cache, kmer_size, bands, trace, band_lower_left);
            RG_CHK(cudaDeviceSynchronize());
            const int32_t BLOCK_LEN = 64;                           /* opt.cuda_block_size default, src/f5c.c:1199 */
            dim3 blockpost(BLOCK_LEN);
            dim3 grid1post((n + (BLOCK_LEN / 32) - 1) / (BLOCK_LEN / 32));                                 /* WARP_HACK, :944-960 */
            align_kernel_post<<<grid1post, blockpost>>>(event_align_pairs, n_event_align_pairs, read_len, read_ptr,
                                                        event_table, n_events, event_ptr, scalings, n, model_kmer_cache,
                                                        kmer_size, bands, trace, band_lower_left);
            RG_CHK(cudaDeviceSynchronize());
        }
        RG_CHK(cudaEventRecord(ev1, 0));
        RG_CHK(cudaGetLastError());
        RG_CHK(cudaMemcpy(n_event_align_pairs_host, n_event_align_pairs, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost)); /* :979-985 */
        RG_CHK(cudaMemcpy(event_align_pairs_host, event_align_pairs, 2 * (size_t)sum_n_events * sizeof(AlignedPair), cudaMemcpyDeviceToHost));
        for (int32_t j = 0; j < n; j++) {                                                                  /* :1005-1030 */
            const int32_t i = idx[j];
            const int32_t np = n_event_align_pairs_host[j];
            n_pairs[i] = np;
            const AlignedPair* in_2 = &event_align_pairs_host[event_ptr_host[j] * 2];
            abea_pair_t* out_2 = pairs + pair_ptr[i];
            int32_t end = np - 1;
            for (int32_t c = 0; c < np; c++, end--) {
                out_2[c].ref_pos = in_2[end].ref_pos;
                out_2[c].read_pos = in_2[end].read_pos;
            }
        }
        const double t1 = now_ms();
        float ms = 0.f;
        RG_CHK(cudaEventElapsedTime(&ms, ev0, ev1));
        if (s >= warmup) {
            kernel_ms[s - warmup] = ms;
            e2e_ms[s - warmup] = t1 - t0;
        }
    }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    cudaFree(read); cudaFree(read_ptr); cudaFree(event_ptr); cudaFree(read_len); cudaFree(n_events);
    cudaFree(n_event_align_pairs); cudaFree(event_table); cudaFree(scalings); cudaFree(model); cudaFree(model_kmer_cache);
    cudaFree(event_align_pairs); cudaFree(bands); cudaFree(trace); cudaFree(band_lower_left);
    free(idx); free(read_ptr_host); free(event_ptr_host); free(read_len_host); free(n_events_host); free(scalings_host);
    free(n_event_align_pairs_host); free(read_host); free(event_table_host); free(event_align_pairs_host);
    return 0;
}

} /* extern "C" */
