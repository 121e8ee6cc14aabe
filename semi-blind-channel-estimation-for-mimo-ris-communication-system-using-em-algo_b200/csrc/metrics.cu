// Per-iteration bookkeeping and the metrics the reference drivers compute:
//   NMSE  ||theta^ - h||^2 / ||h||^2          /root/reference/Proposed_method_NMSEvsTp.py:138
//   LLF   exactly as coded (un-squared norms) /root/reference/Proposed method/ML_detecctor.py:55-57,84
//   genie stop | ||theta|| - ||h|| | < 1, l!=0 /root/reference/Proposed method/PM.py:110-112
#include <math.h>

#include "common.cuh"

namespace sbce {

__global__ void k_init_state(Dims d, const cplx* __restrict__ theta0, cplx* __restrict__ theta,
                             int32_t* __restrict__ active, int32_t* __restrict__ stat, int32_t* __restrict__ iters,
                             double* __restrict__ llf, double* __restrict__ lse) {
    const int b = blockIdx.x;
    const size_t n = (size_t)d.L * d.n_rx;
    const bool zero = (d.flags & SBCE_FLAG_ZERO_START) || theta0 == nullptr;
    for (size_t e = threadIdx.x; e < n; e += blockDim.x)
        theta[b * n + e] = zero ? mk(0.0, 0.0) : theta0[b * n + e];
    if (threadIdx.x == 0) {
        active[b] = 1;
        stat[b] = 0;
        if (iters) iters[b] = 0;
    }
    const double nanv = nan("");
    for (int l = threadIdx.x; l < d.itera; l += blockDim.x) {
        if (llf) llf[(size_t)b * d.itera + l] = nanv;
        if (lse) lse[(size_t)b * d.itera + l] = nanv;
    }
}

cudaError_t launch_init_state(const Dims& d, int nb, const double* theta0, double* theta, int32_t* active,
                              int32_t* stat, int32_t* iters, double* llf, double* lse, cudaStream_t s) {
    k_init_state<<<nb, 128, 0, s>>>(d, (const cplx*)theta0, (cplx*)theta, active, stat, iters, llf, lse);
    count_launch();
    return cudaGetLastError();
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

// sum_t || y_t - Theta^T (psi~_t (x) x_t) ||^2 over a block of symbols (thread per symbol)
__device__ double residual_sq(const Dims& d, int T, const cplx* __restrict__ Y, const cplx* __restrict__ Psi,
                              const cplx* __restrict__ X, const cplx* __restrict__ th) {
    double acc = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        cplx pred[8];
        for (int r = 0; r < d.n_rx; ++r) pred[r] = mk(0.0, 0.0);
        for (int n = 0; n < d.N1; ++n) {
            const cplx p = Psi[(size_t)t * d.N1 + n];
            for (int j = 0; j < d.n_tx; ++j) {
                const cplx w = cmul(p, X[(size_t)t * d.n_tx + j]);
                const cplx* row = th + (size_t)(n * d.n_tx + j) * d.n_rx;
                for (int r = 0; r < d.n_rx; ++r) cfma(pred[r], w, row[r]);
            }
        }
        for (int r = 0; r < d.n_rx; ++r) acc += cnorm2(csub(Y[(size_t)t * d.n_rx + r], pred[r]));
    }
    return acc;
}

__global__ void __launch_bounds__(256) k_after_iteration(Dims d, int l, const cplx* __restrict__ theta,
                                                         const cplx* __restrict__ h_true, const cplx* __restrict__ Yp,
                                                         const cplx* __restrict__ Yd, const cplx* __restrict__ PsiP,
                                                         const cplx* __restrict__ PsiD, const cplx* __restrict__ Xp,
                                                         const cplx* __restrict__ Xd_true,
                                                         const double* __restrict__ varn,
                                                         const double* __restrict__ lse_sym,
                                                         int32_t* __restrict__ active, int32_t* __restrict__ iters,
                                                         double* __restrict__ llf, double* __restrict__ lse) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    if (active[b] == 0) return;
    const size_t n = (size_t)d.L * d.n_rx;
    const cplx* th = theta + b * n;
    if (threadIdx.x == 0 && iters) iters[b] = l + 1;
    if (lse && lse_sym) {
        double a = 0.0;
        for (int t = threadIdx.x; t < d.T_d; t += blockDim.x) a += lse_sym[(size_t)b * d.T_d + t];
        a = block_sum(a, red);
        if (threadIdx.x == 0) lse[(size_t)b * d.itera + l] = a;
    }
    if (llf && Xd_true) {
        const size_t pb = d.psi_shared ? 0 : b, ppb = d.psiP_shared ? 0 : b;
        double rp = residual_sq(d, d.T_p, Yp + (size_t)b * d.T_p * d.n_rx, PsiP + ppb * d.T_p * d.N1,
                                Xp + (size_t)b * d.T_p * d.n_tx, th);
        rp = block_sum(rp, red);
        double rd = residual_sq(d, d.T_d, Yd + (size_t)b * d.T_d * d.n_rx, PsiD + pb * d.T_d * d.N1,
                                Xd_true + (size_t)b * d.T_d * d.n_tx, th);
        rd = block_sum(rd, red);
        if (threadIdx.x == 0) {
            const double v2 = varn[b] * varn[b];
            const double e1 = (double)d.T_d * d.n_tx * log((double)d.M);
            const double e2 = (double)(d.T_d + d.T_p) * log(M_PI * v2);
            llf[(size_t)b * d.itera + l] = -e1 - e2 - sqrt(rp) / v2 - sqrt(rd) / v2;
        }
    }
    if ((d.flags & SBCE_FLAG_GENIE_STOP) && h_true != nullptr) {
        double a = 0.0, c = 0.0;
        for (size_t e = threadIdx.x; e < n; e += blockDim.x) {
            a += cnorm2(th[e]);
            c += cnorm2(h_true[b * n + e]);
        }
        a = block_sum(a, red);
        c = block_sum(c, red);
        // em_zf's stop has no `l != 0` guard (PMvsMLvsZFvsMMSE.py:128); every other estimator's has
        const bool guard_ok = (l != 0) || (d.mode == SBCE_MODE_ZF && !(d.flags & SBCE_FLAG_ZF_STOP_GUARD));
        if (threadIdx.x == 0 && guard_ok && fabs(sqrt(a) - sqrt(c)) < 1.0) active[b] = 0;
    }
}

cudaError_t launch_after_iteration(const Dims& d, int nb, int l, const double* theta, const double* h_true,
                                   const double* Yp, const double* Yd, const double* PsiP, const double* PsiD,
                                   const double* Xp, const double* Xd_true, const double* varn, const double* lse_sym,
                                   int32_t* active, int32_t* iters, double* llf, double* lse, cudaStream_t s) {
    k_after_iteration<<<nb, 256, 0, s>>>(d, l, (const cplx*)theta, (const cplx*)h_true, (const cplx*)Yp,
                                         (const cplx*)Yd, (const cplx*)PsiP, (const cplx*)PsiD, (const cplx*)Xp,
                                         (const cplx*)Xd_true, varn, lse_sym, active, iters, llf, lse);
    count_launch();
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_final_metrics(Dims d, const cplx* __restrict__ theta,
                                                       const cplx* __restrict__ h_true,
                                                       const int32_t* __restrict__ stat, double* __restrict__ nmse,
                                                       int32_t* __restrict__ status) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    const size_t n = (size_t)d.L * d.n_rx;
    if (nmse && h_true) {
        double a = 0.0, c = 0.0;
        for (size_t e = threadIdx.x; e < n; e += blockDim.x) {
            const cplx h = h_true[b * n + e];
            a += cnorm2(csub(theta[b * n + e], h));
            c += cnorm2(h);
        }
        a = block_sum(a, red);
        c = block_sum(c, red);
        if (threadIdx.x == 0) nmse[b] = a / c;
    }
    if (status && threadIdx.x == 0) status[b] = stat[b];
}

cudaError_t launch_final_metrics(const Dims& d, int nb, const double* theta, const double* h_true, const int32_t* stat,
                                 double* nmse, int32_t* status, cudaStream_t s) {
    k_final_metrics<<<nb, 256, 0, s>>>(d, (const cplx*)theta, (const cplx*)h_true, stat, nmse, status);
    count_launch();
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_accumulate_nmse(const double* __restrict__ nmse,
                                                         const int32_t* __restrict__ status, int batch,
                                                         double* __restrict__ acc) {
    __shared__ double red[8];
    double s = 0.0, cnt = 0.0, bad = 0.0;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        const bool ok = (status == nullptr) || status[b] == 0;
        if (ok) { s += nmse[b]; cnt += 1.0; } else bad += 1.0;
    }
    s = block_sum(s, red);
    cnt = block_sum(cnt, red);
    bad = block_sum(bad, red);
    if (threadIdx.x == 0) { acc[0] += s; acc[1] += cnt; acc[2] += bad; }
}

cudaError_t launch_accumulate_nmse(const double* nmse, const int32_t* status, int batch, double* acc, cudaStream_t s) {
    k_accumulate_nmse<<<1, 256, 0, s>>>(nmse, status, batch, acc);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// FP64 FMA peak: 8 independent dependent-chains per thread, enough warps to fill every SMSP
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
           a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) out[0] = r;
}

cudaError_t run_fp64_peak(double* tflops, double* seconds) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* out = nullptr;
    e = cudaMalloc(&out, 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, blocks = sms * 8, threads = 256;
    k_dfma_peak<<<blocks, threads>>>(out, 64, 1.0);  // warm-up
    double best = 0.0, bests = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak<<<blocks, threads>>>(out, iters, 1.0);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) { best = tf; bests = ms * 1e-3; }
    }
    count_launch(4);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (tflops) *tflops = best;
    if (seconds) *seconds = bests;
    return e;
}

}  // namespace sbce
