// Stand-alone micro-benchmark: what does a plain FP64 instruction (DFMA) cost when it is interleaved
// with FP64 tensor MMAs (DMMA) in the same warp?  The Gram kernel generates its A operand
// p_t = conj(psi_n) psi_n' with 4 DMUL/DFMA per complex value, one scalar FP64 instruction per DMMA.8x8x4.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_mix dmma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

// per iteration: 4 x m16n8k8 (= 16 DMMA.8x8x4) and NF dependent-free DFMAs whose results feed the next A fragment
template <int NF>
__global__ void k_mix(double* out, int iters) {
    double a[4], b[2], f[16];
    for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 2; ++i) b[i] = 1.0 + threadIdx.x * 1e-6 + i;
    for (int i = 0; i < 16; ++i) f[i] = 1.0 + i * 1e-3;
    double c[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) c[u][v] = u + v;
    const double m = 1.0000001, k = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < NF; ++q) f[q & 15] = fma(f[q & 15], m, k);
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = (NF > 0) ? f[q] : a[q];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    }
    double r = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) r += c[u][v];
    for (int i = 0; i < 16; ++i) r += f[i];
    if (r == 123.456) out[0] = r;
}

template <typename F>
static double timeit(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 8);
    const int iters = 4096, wpb = 4, blocks = sms * 4, threads = wpb * 32;   // 16 warps per SM, as the Gram kernel
    const double mmaflop = (double)blocks * wpb * iters * 4 * 2048.0;
    double t0 = timeit([&] { k_mix<0><<<blocks, threads>>>(out, iters); });
    printf("DMMA only           : %.3f ms  %.1f TFLOP/s (MMA flops)\n", t0 * 1e3, mmaflop / t0 / 1e12);
#define RUN(NF) { double t = timeit([&] { k_mix<NF><<<blocks, threads>>>(out, iters); }); \
    printf("16 DMMA + %2d DFMA   : %.3f ms  %.1f TFLOP/s (MMA flops)  -> %.1f SMSP-cycles per extra DFMA\n", NF, t * 1e3, \
           mmaflop / t / 1e12, (t - t0) * 1.965e9 / ((double)iters * NF * 4)); }
    RUN(4) RUN(8) RUN(16) RUN(32)
    return 0;
}
