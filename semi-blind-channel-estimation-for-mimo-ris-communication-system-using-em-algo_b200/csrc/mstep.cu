// M-step of the semi-blind EM estimator: weighted least squares for the channel.
//
// Reference semantics (/root/reference/Proposed_method_NMSEvsTp.py:61-66):
//   denom = sum_p Z_p^H Z_p + sum_{t,k} beta Z^H Z      (D x D, D = L*n_rx)
//   numer = sum_p Z_p^H y_p + sum_{t,k} beta Z^H y_t    (D)
//   theta = np.linalg.solve(denom, numer)
// With Z = (psi~^T (x) x^T) (x) I_nrx every Z^H Z is (w^* w^T) (x) I_nrx, so the
// system is  (G (x) I) theta = vec(B)  with the L x L Hermitian
//   G = sum_t (psi~_t^* psi~_t^T) (x) R_t ,   B = sum_t (psi~_t^* (x) m_t) y_t^T   (L x n_rx)
// (pilots: R = conj(x) x^T, m = conj(x)).  SURVEY.md section 8a-6.
//
// B200 design:
//  * k_gram: the Kronecker / Khatri-Rao structure is applied by index
//    arithmetic.  One thread owns one (n >= n') pair of RIS indices and
//    accumulates the whole n_tx x n_tx block  sum_t conj(psi[t,n]) psi[t,n'] R_t
//    in registers; psi and R_t chunks are staged in shared memory (R_t reads are
//    warp-wide broadcasts).  Only the lower triangle is produced.  k_rhs writes
//    B^H as RP extra rows below the matrix.
//  * k_chol: one CTA per trial, right-looking blocked complex Cholesky (panel
//    width 16) on the augmented lower trapezoid [G ; B^H]: the triangular solve
//    of the panel rows turns the B^H rows into (C^-1 B)^H for free, so only the
//    back substitution C^H theta = z remains.  Panels live in shared memory in a
//    micro-tile-friendly layout, the trailing update uses 4x4 complex register
//    tiles.  A non-positive pivot flags the trial (status bit) instead of
//    poisoning the batch.
#include <math.h>

#include "common.cuh"

namespace sbce {

// ---------------------------------------------------------------------------
// Gram + right-hand side
// ---------------------------------------------------------------------------
constexpr int GR_THREADS = 256;
constexpr int GR_TC = 16;  // symbols per shared-memory chunk

// One thread per lower-triangular (n >= n') pair of RIS indices.
template <int NTX>
__global__ void __launch_bounds__(GR_THREADS) k_gram(Dims d, int T, const cplx* __restrict__ Psi,
                                                     const cplx* __restrict__ sR, const cplx* __restrict__ Ginit,
                                                     cplx* __restrict__ Gout, const int32_t* __restrict__ active) {
    extern __shared__ double2 gsm[];
    const int b = blockIdx.y;
    if (active != nullptr && active[b] == 0) return;
    const int N1 = d.N1;
    const int P = N1 * (N1 + 1) / 2;
    const int item = blockIdx.x * GR_THREADS + threadIdx.x;
    const bool is_pair = item < P;
    int n = 0, np = 0;
    if (is_pair) {
        n = (int)((sqrt(8.0 * item + 1.0) - 1.0) * 0.5);
        while ((n + 1) * (n + 2) / 2 <= item) ++n;
        while (n * (n + 1) / 2 > item) --n;
        np = item - n * (n + 1) / 2;
    }

    cplx* sPsi = gsm;               // [GR_TC][N1]
    cplx* sRr = sPsi + GR_TC * N1;  // [GR_TC][NTX*NTX]

    cplx acc[NTX][NTX];
#pragma unroll
    for (int i = 0; i < NTX; ++i)
#pragma unroll
        for (int j = 0; j < NTX; ++j) acc[i][j] = mk(0.0, 0.0);

    const cplx* psi_b = Psi + (size_t)(d.psi_shared ? 0 : b) * T * N1;
    const cplx* R_b = sR + (size_t)b * T * NTX * NTX;

    for (int t0 = 0; t0 < T; t0 += GR_TC) {
        const int tc = min(GR_TC, T - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < tc * N1; e += GR_THREADS) sPsi[e] = psi_b[(size_t)t0 * N1 + e];
        for (int e = threadIdx.x; e < tc * NTX * NTX; e += GR_THREADS) sRr[e] = R_b[(size_t)t0 * NTX * NTX + e];
        __syncthreads();
        if (is_pair) {
            for (int tt = 0; tt < tc; ++tt) {
                const cplx a = sPsi[tt * N1 + n];
                const cplx c = sPsi[tt * N1 + np];
                const cplx p = cmulc(c, a);  // conj(psi[t,n]) psi[t,n']
                const cplx* Rt = sRr + tt * NTX * NTX;
#pragma unroll
                for (int i = 0; i < NTX; ++i)
#pragma unroll
                    for (int j = 0; j < NTX; ++j) cfma(acc[i][j], p, Rt[i * NTX + j]);
            }
        }
    }
    if (is_pair) {
        const size_t gstride = (size_t)d.Ltot * d.Lp;
        cplx* Gb = Gout + (size_t)b * gstride;
        const cplx* Gi = Ginit ? Ginit + (size_t)b * gstride : nullptr;
#pragma unroll
        for (int i = 0; i < NTX; ++i)
#pragma unroll
            for (int j = 0; j < NTX; ++j) {
                const size_t o = (size_t)(n * NTX + i) * d.Lp + (np * NTX + j);
                cplx v = acc[i][j];
                if (Gi) v = cadd(v, Gi[o]);
                Gb[o] = v;
            }
    }
}

// Right-hand side rows (general n_rx): thread per (n, i, r)
__global__ void k_rhs(Dims d, int T, const cplx* __restrict__ Psi, const cplx* __restrict__ Y,
                      const cplx* __restrict__ sm, const cplx* __restrict__ Ginit, cplx* __restrict__ Gout,
                      const int32_t* __restrict__ active) {
    const int b = blockIdx.y;
    if (active != nullptr && active[b] == 0) return;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = d.N1 * d.n_tx * d.n_rx;
    if (e >= total) return;
    const int r = e % d.n_rx, l = e / d.n_rx, i = l % d.n_tx, n = l / d.n_tx;
    const cplx* psi_b = Psi + (size_t)(d.psi_shared ? 0 : b) * T * d.N1;
    const cplx* m_b = sm + (size_t)b * T * d.n_tx;
    const cplx* y_b = Y + (size_t)b * T * d.n_rx;
    cplx acc = mk(0.0, 0.0);
    for (int t = 0; t < T; ++t) {
        const cplx z = cmul(m_b[(size_t)t * d.n_tx + i], y_b[(size_t)t * d.n_rx + r]);  // m_i y_r
        cfmac(acc, psi_b[(size_t)t * d.N1 + n], z);                                     // psi * conj(m y)
    }
    const size_t gstride = (size_t)d.Ltot * d.Lp;
    const size_t o = (size_t)(d.Lp + r) * d.Lp + l;
    if (Ginit) acc = cadd(acc, Ginit[(size_t)b * gstride + o]);
    Gout[(size_t)b * gstride + o] = acc;
}

// padding rows/cols (identity on the padded diagonal, zero B^H padding rows)
__global__ void k_pad(Dims d, cplx* __restrict__ Gout, const int32_t* __restrict__ active) {
    const int b = blockIdx.y;
    if (active != nullptr && active[b] == 0) return;
    cplx* Gb = Gout + (size_t)b * d.Ltot * d.Lp;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    // padded matrix rows L..Lp-1 (all columns), padded columns L..Lp-1 of the B^H rows, padded B^H rows
    const int padrows = (d.Lp - d.L) + d.RP;
    if (e >= padrows * d.Lp) return;
    const int pr = e / d.Lp, c = e % d.Lp;
    int row;
    if (pr < d.Lp - d.L) {
        row = d.L + pr;
        Gb[(size_t)row * d.Lp + c] = (c == row) ? mk(1.0, 0.0) : mk(0.0, 0.0);
    } else {
        const int r = pr - (d.Lp - d.L);
        row = d.Lp + r;
        if (r >= d.n_rx || c >= d.L) Gb[(size_t)row * d.Lp + c] = mk(0.0, 0.0);
    }
}

template <int NTX>
static cudaError_t run_gram(const Dims& d, int nb, const double* Psi, int T, const double* sR, const double* Ginit,
                            double* Gout, const int32_t* active, cudaStream_t s) {
    const int P = d.N1 * (d.N1 + 1) / 2;
    dim3 grid((P + GR_THREADS - 1) / GR_THREADS, nb);
    size_t smem = sizeof(cplx) * (size_t)(GR_TC * d.N1 + GR_TC * NTX * NTX);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_gram<NTX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_gram<NTX><<<grid, GR_THREADS, smem, s>>>(d, T, (const cplx*)Psi, (const cplx*)sR, (const cplx*)Ginit,
                                               (cplx*)Gout, active);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_gram(const Dims& d, int nb, const double* Psi, int T, const double* sR, const double* Ginit,
                        double* Gout, const int32_t* active, cudaStream_t s) {
    switch (d.n_tx) {
        case 1: return run_gram<1>(d, nb, Psi, T, sR, Ginit, Gout, active, s);
        case 2: return run_gram<2>(d, nb, Psi, T, sR, Ginit, Gout, active, s);
        case 3: return run_gram<3>(d, nb, Psi, T, sR, Ginit, Gout, active, s);
        case 4: return run_gram<4>(d, nb, Psi, T, sR, Ginit, Gout, active, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_rhs_pad(const Dims& d, int nb, const double* Psi, int T, const double* Y, const double* sm,
                           const double* Ginit, double* Gout, const int32_t* active, cudaStream_t s) {
    cudaError_t e;
    {
        const int total = d.N1 * d.n_tx * d.n_rx;
        dim3 grid((total + 127) / 128, nb);
        k_rhs<<<grid, 128, 0, s>>>(d, T, (const cplx*)Psi, (const cplx*)Y, (const cplx*)sm, (const cplx*)Ginit,
                                   (cplx*)Gout, active);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    {
        const int padrows = (d.Lp - d.L) + d.RP;
        dim3 grid((padrows * d.Lp + 127) / 128, nb);
        k_pad<<<grid, 128, 0, s>>>(d, (cplx*)Gout, active);
        count_launch();
        e = cudaGetLastError();
    }
    return e;
}

// ---------------------------------------------------------------------------
// Blocked Cholesky of the augmented trapezoid + back substitution
// ---------------------------------------------------------------------------
constexpr int CH_THREADS = 256;
constexpr int CH_WARPS = CH_THREADS / 32;
constexpr int CH_NB = 16;            // panel width
constexpr int CH_SP = CH_NB + 1;     // row stride (complex) of the shared panel: odd -> conflict-free column walks

__device__ __forceinline__ void dmma16x8x8(double (&c)[4], const double (&a)[4], double b0, double b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}

// LEFT-looking blocked complex Cholesky on the FP64 tensor path (mma.sync m16n8k8.f64, "DMMA"):
// panel k (16 columns, all rows below its diagonal block) is brought up to date with ALL previous
// panels in one register-accumulated product
//   S[r][c] = A[r][c] - sum_{q<k0} C[r][q] conj(C[k0+c][q])
// Each warp owns 16-row tiles of the panel; the operand fragments are loaded straight from the
// factor in global memory (L2 / L1 resident, 16-byte complex loads in fragment order, no shared
// staging), 8 DMMAs per 8 previous columns.  The updated panel goes to shared memory, its diagonal
// block is factored by one warp, the rows below are solved by forward substitution (thread per
// row) and the panel is written once; the trailing matrix is never rewritten.
__global__ void __launch_bounds__(CH_THREADS, 2) k_chol(Dims d, cplx* __restrict__ Gall, cplx* __restrict__ theta,
                                                        const int32_t* __restrict__ active, int32_t* __restrict__ stat) {
    extern __shared__ double2 csm[];
    const int b = blockIdx.x;
    if (active != nullptr && active[b] == 0) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int Lp = d.Lp, Ltot = d.Ltot, ld = d.Lp;
    cplx* A = Gall + (size_t)b * Ltot * Lp;

    cplx* sD = csm;                            // [CH_NB][CH_NB+1] diagonal block / its factor (also reduction scratch)
    cplx* buf = sD + 2 * CH_NB * (CH_NB + 1);  // updated panel [rows][CH_SP]; later theta
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;

    for (int k0 = 0; k0 < Lp; k0 += CH_NB) {
        const int nb = min(CH_NB, Lp - k0);   // multiple of 4
        const int rows = Ltot - k0;           // rows of the panel including its diagonal block
        const int nrt = (rows + 15) >> 4;
        __syncthreads();                      // previous panel fully written (global) and buf free
        for (int rt = warp; rt < nrt; rt += CH_WARPS) {
            const int r0 = rt << 4;
            // accumulators: [n-tile][c0..c3], real and imaginary parts
            double cr[2][4], ci[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { cr[j][e] = 0.0; ci[j][e] = 0.0; }
            const int ra = min(r0 + g, rows - 1), rb8 = min(r0 + g + 8, rows - 1);
            const cplx* pa0 = A + (size_t)(k0 + ra) * ld + tig;
            const cplx* pa1 = A + (size_t)(k0 + rb8) * ld + tig;
            const cplx* pb0 = A + (size_t)min(k0 + g, Ltot - 1) * ld + tig;
            const cplx* pb1 = A + (size_t)min(k0 + 8 + g, Ltot - 1) * ld + tig;
#pragma unroll 2
            for (int q0 = 0; q0 < k0; q0 += 8) {
                const cplx a0 = pa0[q0], a1 = pa1[q0], a2 = pa0[q0 + 4], a3 = pa1[q0 + 4];
                const cplx b00 = pb0[q0], b01 = pb0[q0 + 4], b10 = pb1[q0], b11 = pb1[q0 + 4];
                const double ar[4] = {a0.x, a1.x, a2.x, a3.x};
                const double ai[4] = {a0.y, a1.y, a2.y, a3.y};
                // sum_q a conj(b):  re += ar br + ai bi ;  im += ai br - ar bi
                dmma16x8x8(cr[0], ar, b00.x, b01.x);
                dmma16x8x8(cr[0], ai, b00.y, b01.y);
                dmma16x8x8(ci[0], ai, b00.x, b01.x);
                dmma16x8x8(ci[0], ar, -b00.y, -b01.y);
                dmma16x8x8(cr[1], ar, b10.x, b11.x);
                dmma16x8x8(cr[1], ai, b10.y, b11.y);
                dmma16x8x8(ci[1], ai, b10.x, b11.x);
                dmma16x8x8(ci[1], ar, -b10.y, -b11.y);
            }
            // S = A - acc  -> shared panel (fragment: rows g / g+8, columns 8j + 2 tig + {0,1})
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int lr = r0 + g + 8 * h;
                    const int c = 8 * j + 2 * tig;
                    if (lr < rows && c < nb) {
                        const cplx* src = A + (size_t)(k0 + lr) * ld + k0 + c;
                        const cplx v0 = src[0], v1 = src[1];
                        buf[lr * CH_SP + c] = mk(v0.x - cr[j][2 * h], v0.y - ci[j][2 * h]);
                        buf[lr * CH_SP + c + 1] = mk(v1.x - cr[j][2 * h + 1], v1.y - ci[j][2 * h + 1]);
                    }
                }
        }
        __syncthreads();
        // ---- diagonal block (local rows 0..nb) -> sD, unblocked Cholesky by warp 0 (lane = row)
        if (tid < 32) {
            const int r = tid;
            if (r < nb)
                for (int c = 0; c < nb; ++c) sD[r * (CH_NB + 1) + c] = (c <= r) ? buf[r * CH_SP + c] : mk(0.0, 0.0);
            __syncwarp();
            for (int c = 0; c < nb; ++c) {
                double piv = sD[c * (CH_NB + 1) + c].x;
                if (!(piv > 0.0)) {
                    if (r == 0) s_bad = 1;
                    piv = 1.0;
                }
                const double dg = sqrt(piv);
                const double inv = 1.0 / dg;
                __syncwarp();
                if (r == c) sD[c * (CH_NB + 1) + c] = mk(dg, 0.0);
                if (r > c && r < nb) sD[r * (CH_NB + 1) + c] = cscale(sD[r * (CH_NB + 1) + c], inv);
                __syncwarp();
                if (r > c && r < nb) {
                    const cplx lrc = sD[r * (CH_NB + 1) + c];
                    for (int q = c + 1; q <= r; ++q) cfmsc(sD[r * (CH_NB + 1) + q], lrc, sD[q * (CH_NB + 1) + c]);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- write the factored diagonal block; rows below: forward substitution X D^H = S (thread per row)
        for (int e = tid; e < nb * nb; e += CH_THREADS) {
            const int r = e / nb, c = e % nb;
            if (c <= r) A[(size_t)(k0 + r) * ld + k0 + c] = sD[r * (CH_NB + 1) + c];
        }
        for (int lr = nb + tid; lr < rows; lr += CH_THREADS) {
            cplx* xr = buf + lr * CH_SP;  // the row is solved in place in shared memory
            cplx* dst = A + (size_t)(k0 + lr) * ld + k0;
            for (int c = 0; c < nb; ++c) {
                cplx v = xr[c];
                const cplx* dc = sD + c * (CH_NB + 1);
                for (int q = 0; q < c; ++q) cfmsc(v, xr[q], dc[q]);  // v -= x_q conj(D[c][q])
                v = cscale(v, 1.0 / dc[c].x);
                xr[c] = v;
                dst[c] = v;
            }
        }
    }
    cplx* Ps = buf;
    __syncthreads();
    if (tid == 0 && s_bad && stat) atomicOr(&stat[b], SBCE_ST_NOT_PD);

    // ---- back substitution  C^H theta = z,  z[l][r] = conj(A[Lp + r][l])
    // theta kept in shared (reuse Ps): th[l * RPp + r]
    const int nrx = d.n_rx;
    cplx* th = Ps;
    for (int e = tid; e < Lp * nrx; e += CH_THREADS) {
        const int l = e / nrx, r = e % nrx;
        th[e] = cconj(A[(size_t)(Lp + r) * ld + l]);
    }
    __syncthreads();
    for (int k0 = ((Lp - 1) / CH_NB) * CH_NB; k0 >= 0; k0 -= CH_NB) {
        const int nb = min(CH_NB, Lp - k0);
        const int c1 = k0 + nb;
        // th[k0+c][r] -= sum_{row >= c1} conj(A[row][k0+c]) th[row][r]
        const int nout = nb * nrx;
        // split rows among CH_THREADS/nout groups (nout <= 128)
        const int groups = max(1, CH_THREADS / nout);
        cplx part = mk(0.0, 0.0);
        const int o = tid % nout, gidx = tid / nout;
        if (gidx < groups) {
            const int c = o / nrx, r = o % nrx;
            for (int row = c1 + gidx; row < Lp; row += groups) cfmac(part, th[row * nrx + r], A[(size_t)row * ld + k0 + c]);
        }
        // reduce partial sums through shared sD/sW area (>= 2*16*17 cplx = 544)
        cplx* red = sD;
        __syncthreads();
        if (gidx < groups) red[gidx * nout + o] = part;
        __syncthreads();
        if (tid < nout) {
            cplx sum = mk(0.0, 0.0);
            for (int gq = 0; gq < groups; ++gq) sum = cadd(sum, red[gq * nout + tid]);
            th[(k0 + tid / nrx) * nrx + (tid % nrx)] = csub(th[(k0 + tid / nrx) * nrx + (tid % nrx)], sum);
        }
        __syncthreads();
        // solve the nb x nb upper-triangular system D^H x = rhs sequentially (one thread per rhs column)
        if (tid < nrx) {
            const int r = tid;
            for (int c = nb - 1; c >= 0; --c) {
                cplx v = th[(k0 + c) * nrx + r];
                for (int q = c + 1; q < nb; ++q) cfmsc(v, th[(k0 + q) * nrx + r], A[(size_t)(k0 + q) * ld + k0 + c]);
                const double invd = 1.0 / A[(size_t)(k0 + c) * ld + k0 + c].x;
                th[(k0 + c) * nrx + r] = cscale(v, invd);
            }
        }
        __syncthreads();
    }
    cplx* out = theta + (size_t)b * d.L * nrx;
    bool bad = false;
    for (int e = tid; e < d.L * nrx; e += CH_THREADS) {
        const cplx v = th[e];
        out[e] = v;
        if (!isfinite(v.x) || !isfinite(v.y)) bad = true;
    }
    if (bad && stat) atomicOr(&stat[b], SBCE_ST_NONFINITE);
}

cudaError_t launch_chol_solve(const Dims& d, int nb, double* G, double* theta, const int32_t* active, int32_t* stat,
                              cudaStream_t s) {
    size_t panel = (size_t)d.Ltot * CH_SP;
    size_t thsz = (size_t)d.Lp * d.n_rx;
    size_t smem = sizeof(cplx) * (2 * CH_NB * (CH_NB + 1) + (panel > thsz ? panel : thsz));
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_chol<<<nb, CH_THREADS, smem, s>>>(d, (cplx*)G, (cplx*)theta, active, stat);
    count_launch();
    return cudaGetLastError();
}

}  // namespace sbce
