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
constexpr int CH_NB = 16;

// shared-memory panel layout: element (local row lr, panel column q) lives at
//   Ps[(q*4 + (lr&3)) * nbr + (lr>>2)]      nbr = block rows of the panel (padded odd)
// so that the 4 rows of a micro-tile sit in 4 planes and consecutive micro-tile
// indices are consecutive 16-byte words (conflict-free, broadcast-friendly).
__device__ __forceinline__ int ps_index(int lr, int q, int nbr) { return (q * 4 + (lr & 3)) * nbr + (lr >> 2); }

__global__ void __launch_bounds__(CH_THREADS) k_chol(Dims d, cplx* __restrict__ Gall, cplx* __restrict__ theta,
                                                     const int32_t* __restrict__ active, int32_t* __restrict__ stat) {
    extern __shared__ double2 csm[];
    const int b = blockIdx.x;
    if (active != nullptr && active[b] == 0) return;
    const int tid = threadIdx.x;
    const int Lp = d.Lp, Ltot = d.Ltot, ld = d.Lp;
    cplx* A = Gall + (size_t)b * Ltot * Lp;

    const int nbr_max = (Ltot / 4) | 1;
    cplx* sD = csm;                        // [CH_NB][CH_NB+1] diagonal block / its factor
    cplx* sW = sD + CH_NB * (CH_NB + 1);   // [CH_NB][CH_NB+1] inverse of the factor
    cplx* Ps = sW + CH_NB * (CH_NB + 1);   // panel, CH_NB*4*nbr_max
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;

    for (int k0 = 0; k0 < Lp; k0 += CH_NB) {
        const int nb = min(CH_NB, Lp - k0);
        const int c1 = k0 + nb;  // first trailing row/col
        __syncthreads();
        // ---- diagonal block -> shared
        for (int e = tid; e < nb * nb; e += CH_THREADS) {
            const int r = e / nb, c = e % nb;
            sD[r * (CH_NB + 1) + c] = (c <= r) ? A[(size_t)(k0 + r) * ld + k0 + c] : mk(0.0, 0.0);
        }
        __syncthreads();
        // ---- unblocked Cholesky of the nb x nb block by warp 0 (lane = row), then its inverse
        if (tid < 32) {
            const int r = tid;
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
            // W = inverse of lower-triangular factor: lane = column of W
            if (r < nb) {
                const int c = r;
                for (int i = 0; i < nb; ++i) {
                    cplx v = mk(0.0, 0.0);
                    if (i == c) v = mk(1.0 / sD[i * (CH_NB + 1) + i].x, 0.0);
                    else if (i > c) {
                        cplx acc = mk(0.0, 0.0);
                        for (int q = c; q < i; ++q) cfma(acc, sD[i * (CH_NB + 1) + q], sW[q * (CH_NB + 1) + c]);
                        const double invd = 1.0 / sD[i * (CH_NB + 1) + i].x;
                        v = mk(-acc.x * invd, -acc.y * invd);
                    }
                    sW[i * (CH_NB + 1) + c] = v;
                }
            }
        }
        __syncthreads();
        // ---- write the factored diagonal block back
        for (int e = tid; e < nb * nb; e += CH_THREADS) {
            const int r = e / nb, c = e % nb;
            if (c <= r) A[(size_t)(k0 + r) * ld + k0 + c] = sD[r * (CH_NB + 1) + c];
        }
        // ---- panel rows: X = A21 * W^H   (thread per row), to global and to the shared panel
        const int nrows = Ltot - c1;
        const int nbr = (nrows / 4) | 1;
        for (int lr = tid; lr < nrows; lr += CH_THREADS) {
            cplx* rowp = A + (size_t)(c1 + lr) * ld + k0;
            cplx a[CH_NB];
#pragma unroll
            for (int q = 0; q < CH_NB; ++q) a[q] = (q < nb) ? rowp[q] : mk(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < CH_NB; ++c) {
                if (c < nb) {
                    cplx x = mk(0.0, 0.0);
#pragma unroll
                    for (int q = 0; q < CH_NB; ++q)
                        if (q <= c) cfmac(x, a[q], sW[c * (CH_NB + 1) + q]);  // a_q * conj(W[c][q])
                    rowp[c] = x;
                    Ps[ps_index(lr, c, nbr)] = x;
                }
            }
        }
        __syncthreads();
        // ---- trailing update with 4x4 micro-tiles: A[r][c] -= sum_q X[r][q] conj(X[c][q]),  c in [c1, Lp)
        const int nr = nrows / 4;         // block rows (Ltot, Lp, c1 are multiples of 4)
        const int nc = (Lp - c1) / 4;     // block cols
        const int ntri = nc * (nc + 1) / 2;
        const int nblk = ntri + (nr - nc) * nc;
        for (int p = tid; p < nblk; p += CH_THREADS) {
            int br, bc;
            if (p < ntri) {
                br = (int)((sqrtf(8.0f * p + 1.0f) - 1.0f) * 0.5f);
                while ((br + 1) * (br + 2) / 2 <= p) ++br;
                while (br * (br + 1) / 2 > p) --br;
                bc = p - br * (br + 1) / 2;
            } else {
                const int q = p - ntri;
                br = nc + q / nc;
                bc = q % nc;
            }
            cplx acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = mk(0.0, 0.0);
            for (int q = 0; q < nb; ++q) {
                cplx xa[4], xb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) xa[i] = Ps[(q * 4 + i) * nbr + br];
#pragma unroll
                for (int j = 0; j < 4; ++j) xb[j] = Ps[(q * 4 + j) * nbr + bc];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) cfmac(acc[i][j], xa[i], xb[j]);
            }
            cplx* base = A + (size_t)(c1 + 4 * br) * ld + (c1 + 4 * bc);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    cplx v = base[(size_t)i * ld + j];
                    base[(size_t)i * ld + j] = csub(v, acc[i][j]);
                }
        }
    }
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
    const int nbr_max = (d.Ltot / 4) | 1;
    size_t panel = (size_t)CH_NB * 4 * nbr_max;
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
