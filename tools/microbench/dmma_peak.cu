// Stand-alone micro-benchmark: FP64 throughput of DFMA vs mma.sync (DMMA) shapes on sm_100a.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) out[0] = r;
}

__global__ void k_dmma884(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c0[8], c1[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { c0[u] = u; c1[u] = -u; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[u]), "+d"(c1[u]) : "d"(a), "d"(b));
    }
    double r = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) r += c0[u] + c1[u];
    if (r == 123.456) out[0] = r;
}

__global__ void k_dmma1688(double* out, int iters) {
    double a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 2; ++i) b[i] = 1.0 + threadIdx.x * 1e-6 + i;
    double c[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) c[u][v] = u + v;
    for (int i = 0; i < iters; ++i) {
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
    const int iters = 4096;
    for (int wpb = 4; wpb <= 16; wpb *= 2) {
        const int blocks = sms * 4, threads = wpb * 32;
        double t = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
        printf("DFMA      warps/block %2d: %.2f TFLOP/s\n", wpb, 2.0 * 64 * iters * (double)blocks * threads / t / 1e12);
        t = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
        printf("DMMA 8x8x4   warps/block %2d: %.2f TFLOP/s\n", wpb, 2.0 * 256 * 8 * iters * (double)blocks * wpb / t / 1e12);
        t = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters); });
        printf("DMMA 16x8x8  warps/block %2d: %.2f TFLOP/s\n", wpb, 2.0 * 1024 * 4 * iters * (double)blocks * wpb / t / 1e12);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
