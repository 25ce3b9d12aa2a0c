// microbench.cu — per-op latency / throughput on the target GPU for the instructions the ABEA fill kernel leans on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o microbench microbench.cu
// Output: cycles per dependent op (latency, 1 warp/SM) and warp-instructions per cycle per SM (throughput).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { FADD, FMUL, DADD, CVT_F2D2F, CVT_F2D, CVT_D2F, FDIV, SHFL, SHFL_BCAST, RCP, FSEL, IADD, LOP, DSETP, DMUL, I2F, FMNMX, F2D_INT, NUM_OPS };
const char* names[] = {"FADD", "FMUL", "DADD", "F2F.F64.F32+DADD+F2F.F32.F64", "F2F f32->f64", "F2F f64->f32", "fdiv_rn", "SHFL.UP", "SHFL.IDX", "MUFU.RCP", "FSETP+FSEL", "IADD", "LOP3", "DSETP+SEL", "DMUL", "I2F", "FMNMX", "f32->f64 via int ops"};

template <int OP, int ILP>
__global__ void k(float* out, int iters, float seed, long long* cyc) {
    float f[ILP]; double d[ILP]; int n[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) { f[j] = seed + j + threadIdx.x * 1e-3f; d[j] = seed + j * 0.5; n[j] = threadIdx.x + j; }
    double dc = seed * 0.37; float fc = seed * 1.0001f;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int j = 0; j < ILP; j++) {
                if (OP == FADD) f[j] = __fadd_rn(f[j], fc);
                if (OP == FMUL) f[j] = __fmul_rn(f[j], fc);
                if (OP == DADD) d[j] = __dadd_rn(d[j], dc);
                if (OP == DMUL) d[j] = __dmul_rn(d[j], dc);
                if (OP == CVT_F2D2F) f[j] = __double2float_rn(__dadd_rn((double)f[j], dc));
                if (OP == CVT_F2D) { d[j] = (double)__int_as_float(__double2loint(d[j]) | 0x3f800000); }
                if (OP == CVT_D2F) { f[j] = __double2float_rn(__hiloint2double(__float_as_int(f[j]), 0x12345678)); }
                if (OP == FDIV) f[j] = __fdiv_rn(f[j], fc);
                if (OP == SHFL) f[j] = __shfl_up_sync(0xffffffffu, f[j], 1);
                if (OP == SHFL_BCAST) f[j] = __shfl_sync(0xffffffffu, f[j], 3);
                if (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
                if (OP == FSEL) f[j] = f[j] > fc ? f[j] : __fadd_rn(fc, 0.0f) ;
                if (OP == FMNMX) f[j] = fmaxf(f[j], fc);
                if (OP == IADD) n[j] = n[j] + iters;
                if (OP == LOP) n[j] = (n[j] ^ iters) & 0x7fffffff;
                if (OP == DSETP) d[j] = d[j] > dc ? d[j] : dc * 0.5;
                if (OP == I2F) { f[j] = (float)__float_as_int(f[j]); }
                if (OP == F2D_INT) { int b = __float_as_int(f[j]); int hi = (b & 0x80000000) | (((b & 0x7fffffff) >> 3) + 0x38000000); d[j] = __hiloint2double(hi, b << 29); f[j] = __int_as_float(__double2hiint(d[j])); }
            }
        }
    }
    long long t1 = clock64();
    float acc = 0; double dacc = 0; int nacc = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) { acc += f[j]; dacc += d[j]; nacc += n[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)dacc + nacc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(float* out, long long* cyc) {
    const int iters = 256;
    long long h;
    // latency: one warp, ILP 1
    k<OP, 1><<<1, 32>>>(out, iters, 1.5f, cyc); cudaDeviceSynchronize();
    k<OP, 1><<<1, 32>>>(out, iters, 1.5f, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double lat = (double)h / (iters * 16.0);
    // throughput: 1 block/SM x 16 warps (4 per SMSP), ILP 4
    k<OP, 4><<<148, 512>>>(out, iters, 1.5f, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double tp16 = (iters * 16.0 * 4 * 16) / (double)h; // warp-instr per cycle per SM
    // throughput with 1 warp per SMSP (4 warps), ILP 4
    k<OP, 4><<<148, 128>>>(out, iters, 1.5f, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double tp4 = (iters * 16.0 * 4 * 4) / (double)h;
    printf("%-28s latency %6.1f cyc | warp-ops/clk/SM: 16 warps %5.2f, 4 warps(ILP4) %5.2f\n", names[OP], lat, tp16, tp4);
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    run<FADD>(out, cyc); run<FMUL>(out, cyc); run<DADD>(out, cyc); run<DMUL>(out, cyc);
    run<CVT_F2D2F>(out, cyc); run<CVT_F2D>(out, cyc); run<CVT_D2F>(out, cyc); run<F2D_INT>(out, cyc);
    run<FDIV>(out, cyc); run<RCP>(out, cyc); run<SHFL>(out, cyc); run<SHFL_BCAST>(out, cyc);
    run<FSEL>(out, cyc); run<FMNMX>(out, cyc); run<DSETP>(out, cyc); run<IADD>(out, cyc); run<LOP>(out, cyc); run<I2F>(out, cyc);
    return 0;
}
