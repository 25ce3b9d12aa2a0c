/* simt_emu.h — TEST INFRASTRUCTURE: a tiny lock-step SIMT emulator for the CPU.
 *
 * Lets tests/ compile f5c_b200/csrc/*.cu{,h} UNCHANGED with g++ (-DABEA_SIMT_EMU) and run the kernels' exact
 * control flow, lane mapping, shuffles, trace packing and host packer on a box without a GPU. Every CUDA thread of a
 * block is a ucontext fiber; a warp collective (__shfl_*_sync, __ballot_sync, __syncwarp) parks the fiber until all
 * live lanes of its warp have reached a collective, then the exchange is resolved; __syncthreads parks until all live
 * threads of the block arrive. Blocks run one after another. CUDA runtime calls used by the host code are mapped to
 * malloc/memcpy. Float intrinsics map to plain IEEE operations (the build uses -ffp-contract=off), so results are
 * bit-identical to what the GPU must produce.
 *
 * This is NOT a CPU fallback of the product: it lives under tests/, builds tests/simt/libabea_emu.so, and is loaded
 * only by the "not gpu" unit tests. It is far too slow for anything else (every shuffle costs ~100 context switches).
 */
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>

#include <functional>
#include <vector>

/* ---- CUDA keywords / vector types ------------------------------------------------------------------------- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ static
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n)

struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }

namespace simt {

struct Dim3 {
    unsigned x, y, z;
    Dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

enum { OP_SHFL_IDX, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_SYNCWARP };
enum { ST_READY, ST_WAIT_WARP, ST_WAIT_BLOCK, ST_DONE };

struct Fiber {
    ucontext_t ctx;
    char* stack;
    int state;
    Dim3 tid;
    int warp, lane;
    int op;
    uint64_t payload;
    int arg;
    uint64_t result;
};

extern Fiber* cur;
extern Dim3 g_blockIdx, g_blockDim, g_gridDim;
extern void* g_dynsmem; /* dynamic shared memory of the running block */

void launch(Dim3 grid, Dim3 block, const std::function<void()>& body, size_t dyn_smem = 0);
uint64_t collective(int op, uint64_t payload, int arg);
void block_barrier();

template <typename T> static inline uint64_t pack(T v) {
    uint64_t u = 0;
    static_assert(sizeof(T) <= 8, "shuffle payload");
    memcpy(&u, &v, sizeof(T));
    return u;
}
template <typename T> static inline T unpack(uint64_t u) {
    T v;
    memcpy(&v, &u, sizeof(T));
    return v;
}

} // namespace simt

typedef simt::Dim3 dim3;
#define threadIdx (simt::cur->tid)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)

/* ---- warp / block collectives ------------------------------------------------------------------------------ */
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    return simt::unpack<T>(simt::collective(simt::OP_SHFL_IDX, simt::pack(v), src));
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    return simt::unpack<T>(simt::collective(simt::OP_SHFL_UP, simt::pack(v), (int)d));
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    return simt::unpack<T>(simt::collective(simt::OP_SHFL_DOWN, simt::pack(v), (int)d));
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return simt::unpack<T>(simt::collective(simt::OP_SHFL_XOR, simt::pack(v), m));
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    return (unsigned)simt::collective(simt::OP_BALLOT, pred ? 1 : 0, 0);
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::collective(simt::OP_SYNCWARP, 0, 0); }
static inline void __syncthreads() { simt::block_barrier(); }
static inline void __threadfence() {}

/* ---- arithmetic intrinsics: plain IEEE (built with -ffp-contract=off) -------------------------------------- */
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __double2float_rn(double a) { return (float)a; }
static inline float __int_as_float(int i) { return simt::unpack<float>((uint64_t)(uint32_t)i); }
static inline int __float_as_int(float f) { return (int)(uint32_t)simt::pack(f); }
static inline float __uint_as_float(unsigned i) { return simt::unpack<float>((uint64_t)i); }
static inline unsigned __float_as_uint(float f) { return (unsigned)simt::pack(f); }
static inline double __hiloint2double(int hi, int lo) {
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    return simt::unpack<double>(u);
}
static inline int __double2hiint(double d) { return (int)(simt::pack(d) >> 32); }
static inline int __double2loint(double d) { return (int)(simt::pack(d) & 0xffffffffu); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __brev(unsigned x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned atomicAnd(unsigned* p, unsigned v) { unsigned o = *p; *p = o & v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }

/* ---- the slice of the CUDA runtime the host code uses ------------------------------------------------------- */
typedef int cudaError_t;
typedef int cudaStream_t_;
typedef cudaStream_t_* cudaStream_t;
struct cudaEvent_st { double t; };
typedef cudaEvent_st* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; };

static inline double simt_now_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static inline const char* cudaGetErrorString(cudaError_t) { return "simt-emu error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    snprintf(p->name, sizeof(p->name), "SIMT-EMU (CPU, tests only)");
    p->multiProcessorCount = 8;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)calloc(1, sizeof(cudaStream_t_)); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)calloc(1, sizeof(cudaStream_t_)); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)calloc(1, sizeof(cudaEvent_st)); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = simt_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

/* kernel launch: the product's ABEA_LAUNCH macro expands to this under ABEA_SIMT_EMU */
#define SIMT_LAUNCH(kern, grid, block, ...) simt::launch(simt::Dim3(grid), simt::Dim3(block), [&]() { kern(__VA_ARGS__); })
#define SIMT_LAUNCH_SMEM(kern, grid, block, smem, ...) simt::launch(simt::Dim3(grid), simt::Dim3(block), [&]() { kern(__VA_ARGS__); }, smem)
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
