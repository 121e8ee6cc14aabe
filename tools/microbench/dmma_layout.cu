// Verifies the assumed fragment layout of mma.sync.aligned.m16n8k8.row.col.f64 on sm_100a.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double* A, const double* B, double* C) {  // A 16x8 row-major, B 8x8 (k x n) row-major, C 16x8
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a0 = A[g * 8 + t], a1 = A[(g + 8) * 8 + t], a2 = A[g * 8 + t + 4], a3 = A[(g + 8) * 8 + t + 4];
    double b0 = B[t * 8 + g], b1 = B[(t + 4) * 8 + g];
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1; C[(g + 8) * 8 + 2 * t] = c2; C[(g + 8) * 8 + 2 * t + 1] = c3;
}
int main() {
    double hA[128], hB[64], hC[128], ref[128];
    for (int i = 0; i < 128; ++i) hA[i] = sin(i * 0.37) + 0.01 * i;
    for (int i = 0; i < 64; ++i) hB[i] = cos(i * 0.91) - 0.02 * i;
    for (int r = 0; r < 16; ++r) for (int c = 0; c < 8; ++c) { double s = 0; for (int q = 0; q < 8; ++q) s += hA[r * 8 + q] * hB[q * 8 + c]; ref[r * 8 + c] = s; }
    double *dA, *dB, *dC; cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(dA, dB, dC); cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
    double err = 0; for (int i = 0; i < 128; ++i) err = fmax(err, fabs(hC[i] - ref[i]));
    printf("m16n8k8 layout max abs err %.3e (%s) status %s\n", err, err < 1e-12 ? "OK" : "MISMATCH", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
