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
//  * Gram: the Kronecker / Khatri-Rao structure is applied by index arithmetic.  One (n >= n') pair of
//    RIS indices owns the whole n_tx x n_tx block  sum_t conj(psi[t,n]) psi[t,n'] R_t ; because R_t is
//    Hermitian its 2 n_tx^2 real accumulators are [p_r ; p_i] (2 x T) times a real (T x n_tx^2) matrix
//    built from diag(R_t) and the upper entries, i.e. the whole Gram is ONE real GEMM that runs on the
//    FP64 tensor path (k_gram_tma4 for n_tx = 4, k_gram_mma<NTX> for n_tx = 5..8 and long rows; scalar
//    k_gram<NTX> below 4).  Only the lower triangle is produced.  The last CTA of each trial writes B^H as extra
//    rows below the matrix, and the padding.
//  * k_chol: one CTA per trial, LEFT-looking blocked complex Cholesky (panel width 16) on the augmented
//    lower trapezoid [G ; B^H], panel updates and the triangular solve as DMMA products; the B^H rows
//    come out as (C^-1 B)^H for free, so only the back substitution C^H theta = z remains.  A
//    non-positive pivot flags the trial (status bit) instead of poisoning the batch.
#include <math.h>

#include "common.cuh"

namespace sbce {

// ---------------------------------------------------------------------------
// Gram + right-hand side
// ---------------------------------------------------------------------------
constexpr int GR_THREADS = 256;
constexpr int GR_TC = 16;  // symbols per shared-memory chunk (default; shrunk at launch for very long psi rows)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// FP64 tensor-path MMA (SASS: 4 x DMMA.8x8x4).  Fragment layout verified by tools/microbench/dmma_layout.cu.
__device__ __forceinline__ void dmma16x8x8(double (&c)[4], const double (&a)[4], double b0, double b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}


constexpr int RH_MAXR = 8;

// The last CTA of every trial: right-hand side rows B^H stored under the matrix,
//   row Lp + r, column l = n*n_tx + i :  conj(B[l][r]) = sum_t psi[t,n] conj(m_t[i] y_t[r]),
// one thread per column l with n_rx accumulators, plus the identity padding of the trapezoid.
template <int NTX>
__device__ __forceinline__ void gram_rhs_cta(const Dims& d, int T, int b, cplx* sPsi, cplx* /*unused*/, const cplx* psi_b,
                                             const cplx* __restrict__ Y, const cplx* __restrict__ sm, const cplx* Gi,
                                             cplx* Gb) {
    const int N1 = d.N1;
    // Chunk length: as many symbols as the CTA's dynamic shared memory holds (the pair CTAs of the same launch
    // size it for their operand ring).  profiles/r01m: with 16-symbol chunks this CTA -- one per trial, two
    // barriers and a global-load round trip per chunk -- ran 2.6x longer than a pair CTA and held 23 % of the
    // resident warp time while feeding the tensor pipe nothing; long chunks make it latency-cheap.
    unsigned dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    const int per_sym = N1 + NTX * (d.n_rx > NTX ? d.n_rx : NTX);
    int TCR = (int)(dyn_bytes / (sizeof(cplx) * per_sym));
    TCR = max(1, min(TCR, 128));
    cplx* sRr = sPsi + (size_t)TCR * N1;
    // ---------------- right-hand side rows + padding
    const int n_rx = d.n_rx, L = d.L;
    const cplx* m_b = sm + (size_t)b * T * NTX;
    const cplx* y_b = Y + (size_t)b * T * n_rx;
    // each thread owns columns l and l + GR_THREADS of a 2*GR_THREADS-wide pass (one pass for L <= 512)
    for (int l0 = 0; l0 < L; l0 += 2 * GR_THREADS) {
        int ls[2], ns[2], is[2];
        bool have[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            ls[u] = l0 + u * GR_THREADS + threadIdx.x;
            have[u] = ls[u] < L;
            ns[u] = have[u] ? ls[u] / NTX : 0;
            is[u] = have[u] ? ls[u] % NTX : 0;
        }
        cplx acc[2][RH_MAXR];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int r = 0; r < RH_MAXR; ++r) acc[u][r] = mk(0.0, 0.0);
        for (int t0 = 0; t0 < T; t0 += TCR) {
            const int tc = min(TCR, T - t0);
            __syncthreads();
            for (int e = threadIdx.x; e < tc * N1; e += GR_THREADS) sPsi[e] = psi_b[(size_t)t0 * N1 + e];
            for (int e = threadIdx.x; e < tc * NTX * n_rx; e += GR_THREADS) {
                const int tt = e / (NTX * n_rx), ii = (e / n_rx) % NTX, r = e % n_rx;
                sRr[e] = cconj(cmul(m_b[(size_t)(t0 + tt) * NTX + ii], y_b[(size_t)(t0 + tt) * n_rx + r]));
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (have[u]) {
                    for (int tt = 0; tt < tc; ++tt) {
                        const cplx p = sPsi[tt * N1 + ns[u]];
                        const cplx* z = sRr + (tt * NTX + is[u]) * n_rx;
#pragma unroll
                        for (int r = 0; r < RH_MAXR; ++r)
                            if (r < n_rx) cfma(acc[u][r], p, z[r]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (have[u]) {
#pragma unroll
                for (int r = 0; r < RH_MAXR; ++r)
                    if (r < n_rx) {
                        const size_t o = (size_t)(d.Lp + r) * d.Lp + ls[u];
                        cplx v = acc[u][r];
                        if (Gi) v = cadd(v, Gi[o]);
                        Gb[o] = v;
                    }
            }
        }
    }
    // padding: identity on the padded diagonal rows L..Lp-1, zero padded columns / rows of B^H
    const int padrows = (d.Lp - d.L) + d.RP;
    for (int e = threadIdx.x; e < padrows * d.Lp; e += GR_THREADS) {
        const int pr = e / d.Lp, c = e % d.Lp;
        if (pr < d.Lp - d.L) {
            const int row = d.L + pr;
            Gb[(size_t)row * d.Lp + c] = (c == row) ? mk(1.0, 0.0) : mk(0.0, 0.0);
        } else {
            const int r = pr - (d.Lp - d.L);
            if (r >= d.n_rx || c >= d.L) Gb[(size_t)(d.Lp + r) * d.Lp + c] = mk(0.0, 0.0);
        }
    }
}

// CTAs 0 .. nP-1: one thread per lower-triangular (n >= n') pair of RIS indices.
// CTA nP (the last one of a trial): the right-hand side rows B^H stored under the matrix,
//   row Lp + r, column l = n*n_tx + i :  conj(B[l][r]) = sum_t psi[t,n] conj(m_t[i] y_t[r]),
// one thread per column l with n_rx accumulators, plus the identity padding of the trapezoid.
// Both kinds of CTA stage the same psi chunks in shared memory.
template <int NTX>
__global__ void __launch_bounds__(GR_THREADS, 2) k_gram(Dims d, int T, int TC, const cplx* __restrict__ Psi,
                                                     const cplx* __restrict__ sR, const cplx* __restrict__ Y,
                                                     const cplx* __restrict__ sm, const cplx* __restrict__ Ginit,
                                                     cplx* __restrict__ Gout, const int32_t* __restrict__ active) {
    extern __shared__ double2 gsm[];
    const int b = blockIdx.y;
    if (active != nullptr && active[b] == 0) return;
    const int N1 = d.N1;
    const int P = N1 * (N1 + 1) / 2;
    cplx* sPsi = gsm;            // [TC][N1]
    cplx* sRr = sPsi + TC * N1;  // [TC][NTX*NTX]  (rhs CTA: sized from the dynamic shared memory it finds)
    const cplx* psi_b = Psi + (size_t)(d.psi_shared ? 0 : b) * T * N1;
    const size_t gstride = (size_t)d.Ltot * d.Lp;
    cplx* Gb = Gout + (size_t)b * gstride;
    const cplx* Gi = Ginit ? Ginit + (size_t)b * gstride : nullptr;

    if (blockIdx.x == gridDim.x - 1) {
        gram_rhs_cta<NTX>(d, T, b, sPsi, sRr, psi_b, Y, sm, Gi, Gb);
        return;
    }

    const int item = blockIdx.x * GR_THREADS + threadIdx.x;
    const bool is_pair = item < P;
    int n = 0, np = 0;
    if (is_pair) {
        n = (int)((sqrt(8.0 * item + 1.0) - 1.0) * 0.5);
        while ((n + 1) * (n + 2) / 2 <= item) ++n;
        while (n * (n + 1) / 2 > item) --n;
        np = item - n * (n + 1) / 2;
    }

    // R_t is Hermitian: only its real diagonal and its upper triangle are used.  For i<j the two block
    // entries  acc[i][j] += p R_ij  and  acc[j][i] += p conj(R_ij)  share their four real products, so we
    // accumulate U = sum pr Rr, V = sum pi Ri, W = sum pr Ri, Z = sum pi Rr (4 FMAs instead of 8) and
    // combine at the end:  acc[i][j] = (U - V, W + Z),  acc[j][i] = (U + V, Z - W).
    constexpr int NPAIR = NTX * (NTX - 1) / 2;
    constexpr int NPAIR1 = NPAIR > 0 ? NPAIR : 1;
    double dg_re[NTX], dg_im[NTX];
    double U[NPAIR1], V[NPAIR1], W[NPAIR1], Z[NPAIR1];
#pragma unroll
    for (int i = 0; i < NTX; ++i) { dg_re[i] = 0.0; dg_im[i] = 0.0; }
#pragma unroll
    for (int q = 0; q < NPAIR1; ++q) { U[q] = 0.0; V[q] = 0.0; W[q] = 0.0; Z[q] = 0.0; }

    const cplx* R_b = sR + (size_t)b * T * NTX * NTX;
    // two-stage cp.async pipeline: chunk c+1 streams into the other buffer while chunk c is consumed
    const int stage_elems = TC * (N1 + NTX * NTX);      // complex elements per stage
    auto issue = [&](int chunk, int buf) {
        const int t0 = chunk * TC;
        const int tc = min(TC, T - t0);
        cplx* dpsi = gsm + buf * stage_elems;
        cplx* dR = dpsi + TC * N1;
        const cplx* spsi = psi_b + (size_t)t0 * N1;
        const cplx* sRg = R_b + (size_t)t0 * NTX * NTX;
        for (int e = threadIdx.x; e < tc * N1; e += GR_THREADS) cp_async16(dpsi + e, spsi + e);
        for (int e = threadIdx.x; e < tc * NTX * NTX; e += GR_THREADS) cp_async16(dR + e, sRg + e);
        cp_async_commit();
    };
    const int nchunk = (T + TC - 1) / TC;
    if (nchunk > 0) issue(0, 0);
    for (int ck = 0; ck < nchunk; ++ck) {
        const int buf = ck & 1;
        if (ck + 1 < nchunk) { issue(ck + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const int tc = min(TC, T - ck * TC);
        const cplx* cPsi = gsm + buf * stage_elems;
        const cplx* cR = cPsi + TC * N1;
        if (is_pair) {
#pragma unroll 4
            for (int tt = 0; tt < tc; ++tt) {
                const cplx a = cPsi[tt * N1 + n];
                const cplx c = cPsi[tt * N1 + np];
                const cplx p = cmulc(c, a);  // conj(psi[t,n]) psi[t,n']
                const cplx* Rt = cR + tt * NTX * NTX;
#pragma unroll
                for (int i = 0; i < NTX; ++i) {
                    const double r = Rt[i * NTX + i].x;
                    dg_re[i] = fma(p.x, r, dg_re[i]);
                    dg_im[i] = fma(p.y, r, dg_im[i]);
                }
#pragma unroll
                for (int i = 0; i < NTX; ++i)
#pragma unroll
                    for (int j = i + 1; j < NTX; ++j) {
                        const int q = i * NTX - i * (i + 1) / 2 + (j - i - 1);
                        const cplx r = Rt[i * NTX + j];
                        U[q] = fma(p.x, r.x, U[q]);
                        V[q] = fma(p.y, r.y, V[q]);
                        W[q] = fma(p.x, r.y, W[q]);
                        Z[q] = fma(p.y, r.x, Z[q]);
                    }
            }
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled (chunk ck+2)
    }
    if (is_pair) {
        auto put = [&](int i, int j, cplx v) {
            const size_t o = (size_t)(n * NTX + i) * d.Lp + (np * NTX + j);
            if (Gi) v = cadd(v, Gi[o]);
            Gb[o] = v;
        };
#pragma unroll
        for (int i = 0; i < NTX; ++i) put(i, i, mk(dg_re[i], dg_im[i]));
#pragma unroll
        for (int i = 0; i < NTX; ++i)
#pragma unroll
            for (int j = i + 1; j < NTX; ++j) {
                const int q = i * NTX - i * (i + 1) / 2 + (j - i - 1);
                put(i, j, mk(U[q] - V[q], W[q] + Z[q]));
                put(j, i, mk(U[q] + V[q], Z[q] - W[q]));
            }
    }
}

// ---------------------------------------------------------------------------
// Gram on the FP64 tensor path, n_tx = 4.
// With the Hermitian product sharing above, the 32 real accumulators of a RIS pair are
//   [pr ; pi] (2 x T)  times  Bmat (T x 16 real columns = [Rd0..Rd3, (Rr,Ri) of the 6 upper entries]),
// i.e. the whole Gram is ONE real GEMM with M = 2 x (number of pairs), N = 16, K = T whose A operand
// p_t = conj(psi[t,n]) psi[t,n'] is generated on the fly into the mma.m16n8k8 fragment layout from the
// staged psi chunk.  A warp owns two 16-pair tiles; per 8 symbols it issues 8 MMAs (= 32 DMMA.8x8x4)
// against 20 shared-memory loads, so the kernel is bound by the FP64 pipe, not by shared memory.
// B-matrix column -> double offset inside the raw 4x4 complex R_t (row-major, interleaved):
__device__ __forceinline__ int gram_bcol_offset(int col) {
    // cols 0..3: real diagonal (i,i) ; cols 4+2q, 5+2q: re, im of the q-th upper entry (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
    const int up[6] = {1, 2, 3, 6, 7, 11};  // 4*i + j
    if (col < 4) return 2 * (5 * col);
    const int q = (col - 4) >> 1;
    return 2 * up[q] + ((col - 4) & 1);
}

// ---------------------------------------------------------------------------
// k_gram_tma4: the FP64 tensor-path Gram fed by the TMA engine.  (A cp.async double-buffered predecessor spent
// only half of its warp time in the DMMA loop, profiles/r01h: 9 % issuing ~10 cp.async per thread and
// chunk, 7 % at the two CTA barriers per chunk that couple all eight warps, 15 % in the epilogue.)  The
// staged operands are CONTIGUOUS in global memory -- 16 symbols of psi ([16][N+1] complex) and of R_t
// ([16][16] complex) -- so one elected lane of a producer warp moves each chunk with two 1-D bulk copies
// (cp.async.bulk ... mbarrier::complete_tx) into a 4-stage ring; the eight DMMA warps never touch the
// staging: each waits on the stage's `full` mbarrier, runs its DMMAs and arrives on `empty` on its own
// -- no __syncthreads in the loop, so a slow warp no longer stalls the other seven.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int GT_TC = 16;        // symbols per stage
constexpr int GT_STAGES = 4;     // ring depth
constexpr int GT_THREADS = GR_THREADS;        // 8 DMMA warps; warp 0's elected lane also drives the TMA ring

__global__ void __launch_bounds__(GT_THREADS, 2) k_gram_tma4(Dims d, int T, const cplx* __restrict__ Psi,
                                                            const cplx* __restrict__ sR, const cplx* __restrict__ Y,
                                                            const cplx* __restrict__ sm, const cplx* __restrict__ Ginit,
                                                            cplx* __restrict__ Gout, const int32_t* __restrict__ active) {
    constexpr int NTX = 4;
    extern __shared__ __align__(128) double2 gsm[];
    __shared__ __align__(8) unsigned long long bar_full[GT_STAGES], bar_empty[GT_STAGES];
    const int b = blockIdx.y;
    if (active != nullptr && active[b] == 0) return;
    const int N1 = d.N1;
    const int P = N1 * (N1 + 1) / 2;
    const cplx* psi_b = Psi + (size_t)(d.psi_shared ? 0 : b) * T * N1;
    const size_t gstride = (size_t)d.Ltot * d.Lp;
    cplx* Gb = Gout + (size_t)b * gstride;
    const cplx* Gi = Ginit ? Ginit + (size_t)b * gstride : nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    if (blockIdx.x == gridDim.x - 1) {
        gram_rhs_cta<NTX>(d, T, b, gsm, gsm + GR_TC * N1, psi_b, Y, sm, Gi, Gb);
        return;
    }
    const int stage_elems = GT_TC * (N1 + NTX * NTX);   // complex elements per stage: [TC][N1] psi, [TC][16] R
    const cplx* R_b = sR + (size_t)b * T * NTX * NTX;
    const int nchunk = (T + GT_TC - 1) / GT_TC;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], GR_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer: fills stage (ck % STAGES) with chunk ck; executed by warp 0 (all lanes for the ragged
    // zero fill, the elected lane for the barrier and the two bulk copies)
    auto produce = [&](int ck) {
        const int s = ck % GT_STAGES;
        const int t0 = ck * GT_TC;
        const int tc = min(GT_TC, T - t0);
        cplx* dpsi = gsm + s * stage_elems;
        cplx* dR = dpsi + GT_TC * N1;
        if (tc < GT_TC) {   // ragged last chunk: symbols beyond T contribute zero
            for (int e = tc * N1 + lane; e < GT_TC * N1; e += 32) dpsi[e] = mk(0.0, 0.0);
            for (int e = tc * NTX * NTX + lane; e < GT_TC * NTX * NTX; e += 32) dR[e] = mk(0.0, 0.0);
            __syncwarp();
        }
        if (lane == 0) {
            const unsigned bpsi = (unsigned)(tc * N1 * sizeof(cplx)), bR = (unsigned)(tc * NTX * NTX * sizeof(cplx));
            mbar_expect_tx(&bar_full[s], bpsi + bR);
            tma_bulk_g2s(dpsi, psi_b + (size_t)t0 * N1, bpsi, &bar_full[s]);
            tma_bulk_g2s(dR, R_b + (size_t)t0 * NTX * NTX, bR, &bar_full[s]);
        }
        __syncwarp();
    };
    if (warp == 0)
        for (int ck = 0; ck < min(GT_STAGES, nchunk); ++ck) produce(ck);

    // ---------------- DMMA warps: pairs of this lane: tile u (0,1), row half h (0,1) -> item
    int pn[2][2], pnp[2][2];
    bool pv[2][2], tile_live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        tile_live[u] = (blockIdx.x * GR_THREADS + (2 * warp + u) * 16) < P;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int item = blockIdx.x * GR_THREADS + (2 * warp + u) * 16 + g + 8 * h;
            pv[u][h] = item < P;
            item = min(item, P - 1);
            int n = (int)((sqrt(8.0 * item + 1.0) - 1.0) * 0.5);
            while ((n + 1) * (n + 2) / 2 <= item) ++n;
            while (n * (n + 1) / 2 > item) --n;
            pn[u][h] = n;
            pnp[u][h] = item - n * (n + 1) / 2;
        }
    }
    const int boff0 = gram_bcol_offset(g), boff1 = gram_bcol_offset(8 + g);
    double accr[2][2][4], acci[2][2][4];  // [tile][n-tile][c0..c3]
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) { accr[u][nt][e] = 0.0; acci[u][nt][e] = 0.0; }

    for (int ck = 0; ck < nchunk; ++ck) {
        const int s = ck % GT_STAGES;
        // refill the stage of the PREVIOUS chunk (three chunks of look-ahead): by now the other warps have
        // almost always released it, so the wait on its `empty` barrier rarely blocks warp 0
        if (warp == 0 && ck >= 1 && ck - 1 + GT_STAGES < nchunk) {
            mbar_wait(&bar_empty[(ck - 1) % GT_STAGES], ((ck - 1) / GT_STAGES) & 1);
            produce(ck - 1 + GT_STAGES);
        }
        mbar_wait(&bar_full[s], (ck / GT_STAGES) & 1);
        const cplx* cPsi = gsm + s * stage_elems;
        const double* cR = (const double*)(cPsi + GT_TC * N1);
#pragma unroll
        for (int ks = 0; ks < GT_TC / 8; ++ks) {
            const int tlo = ks * 8 + tig, thi = tlo + 4;
            const double b00 = cR[tlo * 32 + boff0], b01 = cR[thi * 32 + boff0];
            const double b10 = cR[tlo * 32 + boff1], b11 = cR[thi * 32 + boff1];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!tile_live[u]) continue;   // warp-uniform: ragged last pair CTA of a trial
                // A fragment order: (row g, t lo), (row g+8, t lo), (row g, t hi), (row g+8, t hi)
                const cplx p0 = cmulc(cPsi[tlo * N1 + pnp[u][0]], cPsi[tlo * N1 + pn[u][0]]);
                const cplx p1 = cmulc(cPsi[tlo * N1 + pnp[u][1]], cPsi[tlo * N1 + pn[u][1]]);
                const cplx p2 = cmulc(cPsi[thi * N1 + pnp[u][0]], cPsi[thi * N1 + pn[u][0]]);
                const cplx p3 = cmulc(cPsi[thi * N1 + pnp[u][1]], cPsi[thi * N1 + pn[u][1]]);
                const double pr[4] = {p0.x, p1.x, p2.x, p3.x};
                const double pi[4] = {p0.y, p1.y, p2.y, p3.y};
                dmma16x8x8(accr[u][0], pr, b00, b01);
                dmma16x8x8(accr[u][1], pr, b10, b11);
                dmma16x8x8(acci[u][0], pi, b00, b01);
                dmma16x8x8(acci[u][1], pi, b10, b11);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[s]);   // this warp is done with the stage
    }
    // epilogue: accumulator (row g + 8h, cols 2 tig, 2 tig + 1 of n-tile nt)
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!pv[u][h]) continue;
            const int n = pn[u][h], np = pnp[u][h];
            auto put = [&](int i, int j, cplx v) {
                const size_t o = (size_t)(n * NTX + i) * d.Lp + (np * NTX + j);
                if (Gi) v = cadd(v, Gi[o]);
                Gb[o] = v;
            };
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const double r0 = accr[u][nt][2 * h], r1 = accr[u][nt][2 * h + 1];
                const double i0 = acci[u][nt][2 * h], i1 = acci[u][nt][2 * h + 1];
                const int col = 8 * nt + 2 * tig;
                if (col < 4) {  // two diagonal entries: (pr Rd, pi Rd)
                    put(col, col, mk(r0, i0));
                    put(col + 1, col + 1, mk(r1, i1));
                } else {        // one upper pair q: U = pr Rr, W = pr Ri, Z = pi Rr, V = pi Ri
                    const int q = (col - 4) >> 1;
                    const int qi = (q < 3) ? 0 : (q < 5 ? 1 : 2);
                    const int qj = (q < 3) ? q + 1 : (q < 5 ? q - 1 : 3);
                    put(qi, qj, mk(r0 - i1, r1 + i0));
                    put(qj, qi, mk(r0 + i1, i0 - r1));
                }
            }
        }
}

// Generic tensor-path Gram, n_tx = 4..8.  Column layout of the real B matrix (NC columns, padded to a
// multiple of 8): [0, NTX) the real diagonal of R_t; from the even offset PO on, (re, im) of the upper
// entries (i<j) in row-major order.  A warp owns one 16-pair tile and all NT column tiles; the chunk
// length TC (multiple of 8) is chosen at launch so that two cp.async stages fit in shared memory.
template <int NTX>
struct GramCols {
    static constexpr int NPAIR = NTX * (NTX - 1) / 2;
    static constexpr int PO = (NTX + 1) & ~1;
    static constexpr int NC = PO + 2 * NPAIR;
    static constexpr int NT = (NC + 7) / 8;
    __device__ __forceinline__ static void pair_ij(int q, int& i, int& j) {
        i = 0;
        while (q >= NTX - 1 - i) { q -= NTX - 1 - i; ++i; }
        j = i + 1 + q;
    }
    // column -> double offset inside the raw R_t (row-major NTX x NTX interleaved complex); -1: padding
    __device__ __forceinline__ static int offset(int col) {
        if (col < NTX) return 2 * (col * NTX + col);
        if (col < PO || col >= NC) return -1;
        int i, j;
        pair_ij((col - PO) >> 1, i, j);
        return 2 * (i * NTX + j) + ((col - PO) & 1);
    }
};

constexpr int GW_PAIRS = (GR_THREADS / 32) * 16;   // pairs per CTA of the generic kernel

template <int NTX>
__global__ void __launch_bounds__(GR_THREADS, 1) k_gram_mma(Dims d, int T, int TC, const cplx* __restrict__ Psi,
                                                          const cplx* __restrict__ sR, const cplx* __restrict__ Y,
                                                          const cplx* __restrict__ sm, const cplx* __restrict__ Ginit,
                                                          cplx* __restrict__ Gout, const int32_t* __restrict__ active) {
    typedef GramCols<NTX> GC;
    constexpr int NT = GC::NT;
    constexpr int RS = 2 * NTX * NTX;   // doubles per staged R_t
    extern __shared__ double2 gsm[];
    const int b = blockIdx.y;
    if (active != nullptr && active[b] == 0) return;
    const int N1 = d.N1;
    const int P = N1 * (N1 + 1) / 2;
    const cplx* psi_b = Psi + (size_t)(d.psi_shared ? 0 : b) * T * N1;
    const size_t gstride = (size_t)d.Ltot * d.Lp;
    cplx* Gb = Gout + (size_t)b * gstride;
    const cplx* Gi = Ginit ? Ginit + (size_t)b * gstride : nullptr;
    if (blockIdx.x == gridDim.x - 1) {
        gram_rhs_cta<NTX>(d, T, b, gsm, gsm + GR_TC * N1, psi_b, Y, sm, Gi, Gb);
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    int pn[2], pnp[2];
    bool pv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int item = blockIdx.x * GW_PAIRS + warp * 16 + g + 8 * h;
        pv[h] = item < P;
        item = min(item, P - 1);
        int n = (int)((sqrt(8.0 * item + 1.0) - 1.0) * 0.5);
        while ((n + 1) * (n + 2) / 2 <= item) ++n;
        while (n * (n + 1) / 2 > item) --n;
        pn[h] = n;
        pnp[h] = item - n * (n + 1) / 2;
    }
    int boff[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) boff[nt] = GC::offset(8 * nt + g);
    double accr[NT][4], acci[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { accr[nt][e] = 0.0; acci[nt][e] = 0.0; }

    const cplx* R_b = sR + (size_t)b * T * NTX * NTX;
    const int stage_elems = TC * (N1 + NTX * NTX);
    auto issue = [&](int chunk, int buf) {
        const int t0 = chunk * TC;
        const int tc = min(TC, T - t0);
        cplx* dpsi = gsm + buf * stage_elems;
        cplx* dR = dpsi + TC * N1;
        const cplx* spsi = psi_b + (size_t)t0 * N1;
        const cplx* sRg = R_b + (size_t)t0 * NTX * NTX;
        for (int e = threadIdx.x; e < tc * N1; e += GR_THREADS) cp_async16(dpsi + e, spsi + e);
        for (int e = threadIdx.x; e < tc * NTX * NTX; e += GR_THREADS) cp_async16(dR + e, sRg + e);
        if (tc < TC) {  // ragged last chunk: symbols beyond T contribute zero
            for (int e = tc * N1 + threadIdx.x; e < TC * N1; e += GR_THREADS) dpsi[e] = mk(0.0, 0.0);
            for (int e = tc * NTX * NTX + threadIdx.x; e < TC * NTX * NTX; e += GR_THREADS) dR[e] = mk(0.0, 0.0);
        }
        cp_async_commit();
    };
    const int nchunk = (T + TC - 1) / TC;
    if (nchunk > 0) issue(0, 0);
    for (int ck = 0; ck < nchunk; ++ck) {
        const int buf = ck & 1;
        if (ck + 1 < nchunk) { issue(ck + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const cplx* cPsi = gsm + buf * stage_elems;
        const double* cR = (const double*)(cPsi + TC * N1);
        for (int ks = 0; ks < TC / 8; ++ks) {
            const int tlo = ks * 8 + tig, thi = tlo + 4;
            // A fragment order: (row g, t lo), (row g+8, t lo), (row g, t hi), (row g+8, t hi)
            const cplx p0 = cmulc(cPsi[tlo * N1 + pnp[0]], cPsi[tlo * N1 + pn[0]]);
            const cplx p1 = cmulc(cPsi[tlo * N1 + pnp[1]], cPsi[tlo * N1 + pn[1]]);
            const cplx p2 = cmulc(cPsi[thi * N1 + pnp[0]], cPsi[thi * N1 + pn[0]]);
            const cplx p3 = cmulc(cPsi[thi * N1 + pnp[1]], cPsi[thi * N1 + pn[1]]);
            const double pr[4] = {p0.x, p1.x, p2.x, p3.x};
            const double pi[4] = {p0.y, p1.y, p2.y, p3.y};
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double b0 = boff[nt] >= 0 ? cR[tlo * RS + boff[nt]] : 0.0;
                const double b1 = boff[nt] >= 0 ? cR[thi * RS + boff[nt]] : 0.0;
                dmma16x8x8(accr[nt], pr, b0, b1);
                dmma16x8x8(acci[nt], pi, b0, b1);
            }
        }
        __syncthreads();
    }
    // epilogue: accumulator (row g + 8h, cols 8 nt + 2 tig, + 1)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (!pv[h]) continue;
        const int n = pn[h], np = pnp[h];
        auto put = [&](int i, int j, cplx v) {
            const size_t o = (size_t)(n * NTX + i) * d.Lp + (np * NTX + j);
            if (Gi) v = cadd(v, Gi[o]);
            Gb[o] = v;
        };
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const double r0 = accr[nt][2 * h], r1 = accr[nt][2 * h + 1];
            const double i0 = acci[nt][2 * h], i1 = acci[nt][2 * h + 1];
            const int col = 8 * nt + 2 * tig;
            if (col < NTX) {  // diagonal entries: (pr Rd, pi Rd)
                put(col, col, mk(r0, i0));
                if (col + 1 < NTX) put(col + 1, col + 1, mk(r1, i1));
            } else if (col >= GC::PO && col < GC::NC) {  // upper pair: U = pr Rr, W = pr Ri, Z = pi Rr, V = pi Ri
                int qi, qj;
                GC::pair_ij((col - GC::PO) >> 1, qi, qj);
                put(qi, qj, mk(r0 - i1, r1 + i0));
                put(qj, qi, mk(r0 + i1, i0 - r1));
            }
        }
    }
}

// largest chunk length (multiple of 8, at most 32) whose two stages fit in `budget` bytes
static int gram_chunk(int N1, int ntx, size_t budget) {
    int tc = 32;
    while (tc > 8 && sizeof(cplx) * (size_t)(2 * tc * (N1 + ntx * ntx)) > budget) tc >>= 1;
    return tc;
}

// Shared-memory ceiling of the Gram kernels: two stages of at least 8 symbols must fit in 227 KB.  abi.cu's
// make_dims() rejects longer phase rows with SBCE_E_UNSUPPORTED before any launch (N + 1 + n_tx^2 <= 908).
bool gram_supports(int N1, int n_tx) { return sizeof(cplx) * (size_t)(2 * 8 * (N1 + n_tx * n_tx)) <= 227 * 1024; }

template <int NTX>
static cudaError_t run_gram_wide(const Dims& d, int nb, const double* Psi, int T, const double* sR, const double* Y,
                                 const double* sm, const double* Ginit, double* Gout, const int32_t* active,
                                 cudaStream_t s) {
    static SmemOptIn optin;
    const int P = d.N1 * (d.N1 + 1) / 2;
    dim3 grid((P + GW_PAIRS - 1) / GW_PAIRS + 1, nb);   // + 1: the right-hand-side / padding CTA
    int tc = gram_chunk(d.N1, NTX, 100 * 1024);
    if (sizeof(cplx) * (size_t)(2 * tc * (d.N1 + NTX * NTX)) > 227 * 1024) return cudaErrorInvalidValue;
    const int zsz = NTX * (d.n_rx > NTX ? d.n_rx : NTX);
    size_t smem = sizeof(cplx) * (size_t)(2 * tc * (d.N1 + NTX * NTX));
    const size_t smem_rhs = sizeof(cplx) * (size_t)(GR_TC * d.N1 + GR_TC * zsz);
    if (smem_rhs > smem) smem = smem_rhs;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = opt_in_smem(optin, (const void*)k_gram_mma<NTX>, smem);
    if (e != cudaSuccess) return e;
    k_gram_mma<NTX><<<grid, GR_THREADS, smem, s>>>(d, T, tc, (const cplx*)Psi, (const cplx*)sR, (const cplx*)Y,
                                                   (const cplx*)sm, (const cplx*)Ginit, (cplx*)Gout, active);
    count_launch();
    return cudaGetLastError();
}

// n_tx <= 3: scalar Hermitian-shared kernel, chunk length shrunk for very long phase rows
template <int NTX>
static cudaError_t run_gram_scalar(const Dims& d, int nb, const double* Psi, int T, const double* sR, const double* Y,
                                   const double* sm, const double* Ginit, double* Gout, const int32_t* active,
                                   cudaStream_t s) {
    static SmemOptIn optin;
    const int P = d.N1 * (d.N1 + 1) / 2;
    dim3 grid((P + GR_THREADS - 1) / GR_THREADS + 1, nb);   // + 1: the right-hand-side / padding CTA
    int tc = GR_TC;
    while (tc > 2 && sizeof(cplx) * (size_t)(2 * tc * (d.N1 + NTX * NTX)) > 100 * 1024) tc >>= 1;
    const int zsz = NTX * (d.n_rx > NTX ? d.n_rx : NTX);
    size_t smem = sizeof(cplx) * (size_t)(2 * tc * (d.N1 + NTX * NTX));          // two cp.async stages (pair CTAs)
    const size_t smem_rhs = sizeof(cplx) * (size_t)(d.N1 + zsz);                  // rhs CTA: at least one symbol
    if (smem_rhs > smem) smem = smem_rhs;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = opt_in_smem(optin, (const void*)k_gram<NTX>, smem);
    if (e != cudaSuccess) return e;
    k_gram<NTX><<<grid, GR_THREADS, smem, s>>>(d, T, tc, (const cplx*)Psi, (const cplx*)sR, (const cplx*)Y,
                                               (const cplx*)sm, (const cplx*)Ginit, (cplx*)Gout, active);
    count_launch();
    return cudaGetLastError();
}

// n_tx = 4: TMA-fed tensor-path kernel while its four-stage ring fits in 160 KB (N <= 143), else the generic one
static cudaError_t run_gram4(const Dims& d, int nb, const double* Psi, int T, const double* sR, const double* Y,
                             const double* sm, const double* Ginit, double* Gout, const int32_t* active,
                             cudaStream_t s) {
    constexpr int NTX = 4;
    static SmemOptIn optin;
    size_t smem = sizeof(cplx) * (size_t)(GT_STAGES * GT_TC * (d.N1 + NTX * NTX));
    if (smem > 160 * 1024) return run_gram_wide<4>(d, nb, Psi, T, sR, Y, sm, Ginit, Gout, active, s);
    const int P = d.N1 * (d.N1 + 1) / 2;
    dim3 grid((P + GR_THREADS - 1) / GR_THREADS + 1, nb);   // + 1: the right-hand-side / padding CTA
    const int zsz = NTX * (d.n_rx > NTX ? d.n_rx : NTX);
    const size_t smem_rhs = sizeof(cplx) * (size_t)(GR_TC * d.N1 + GR_TC * zsz);
    if (smem_rhs > smem) smem = smem_rhs;
    cudaError_t e = opt_in_smem(optin, (const void*)k_gram_tma4, smem);
    if (e != cudaSuccess) return e;
    k_gram_tma4<<<grid, GT_THREADS, smem, s>>>(d, T, (const cplx*)Psi, (const cplx*)sR, (const cplx*)Y,
                                               (const cplx*)sm, (const cplx*)Ginit, (cplx*)Gout, active);
    count_launch();
    return cudaGetLastError();
}

// Normal-equation build: lower triangle of G, B^H rows and padding, one launch.
cudaError_t launch_normal_equations(const Dims& d, int nb, const double* Psi, int T, const double* Y,
                                    const double* sm, const double* sR, const double* Ginit, double* Gout,
                                    const int32_t* active, cudaStream_t s) {
    if (d.n_rx > RH_MAXR) return cudaErrorInvalidValue;
    switch (d.n_tx) {
        case 1: return run_gram_scalar<1>(d, nb, Psi, T, sR, Y, sm, Ginit, Gout, active, s);
        case 2: return run_gram_scalar<2>(d, nb, Psi, T, sR, Y, sm, Ginit, Gout, active, s);
        case 3: return run_gram_scalar<3>(d, nb, Psi, T, sR, Y, sm, Ginit, Gout, active, s);
        case 4: return run_gram4(d, nb, Psi, T, sR, Y, sm, Ginit, Gout, active, s);
        case 5: return run_gram_wide<5>(d, nb, Psi, T, sR, Y, sm, Ginit, Gout, active, s);
        case 6: return run_gram_wide<6>(d, nb, Psi, T, sR, Y, sm, Ginit, Gout, active, s);
        case 7: return run_gram_wide<7>(d, nb, Psi, T, sR, Y, sm, Ginit, Gout, active, s);
        case 8: return run_gram_wide<8>(d, nb, Psi, T, sR, Y, sm, Ginit, Gout, active, s);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------
// Blocked Cholesky of the augmented trapezoid + back substitution
// ---------------------------------------------------------------------------
constexpr int CH_NB = 16;            // panel width
constexpr int CH_DS = CH_NB + 1;     // row stride (complex) of the shared 16x16 blocks


// k_chol2: LEFT-looking blocked complex Cholesky (panel width 16) on the FP64 tensor path (mma.sync
// m16n8k8.f64, SASS DMMA) with LOOK-AHEAD.  Iteration k works on the 16-row tiles below the diagonal block of
// panel k, and every tile is taken through TWO steps by the warp that claimed it:
//   A. X = S W_k^H            (triangular solve of panel k as a GEMM against the inverse of the diagonal
//                              block, 16 DMMA quads), then
//   B. S' = A - C[rows, 0:g0] C[g0:g0+16, 0:g0]^H   (update for panel k+1 against ALL previous columns,
//      operand fragments double-buffered in registers so the loads of step q+1 are in flight during the
//      DMMAs of step q).
// Warp 0 takes tile 0 -- the next diagonal block row -- first, publishes "block row ready" through a
// shared flag (the other warps need its freshly solved columns as the B operand of step B), then factors
// the next diagonal block and its inverse W_{k+1} into the other shared buffer while the other warps chew
// through the remaining tiles: one CTA barrier per panel.  The trailing matrix is never touched; the B^H
// rows carried under the matrix come out as (C^-1 B)^H, i.e. the forward substitution is free.  (Its
// two-barrier predecessor with a shared-memory diagonal factor is documented in profiles/r01h-r01j.)
// ---------------------------------------------------------------------------

// Diagonal block of the look-ahead kernel, register resident.  Profiling (profiles/r01i) showed the
// shared-memory version above on warp 0's critical path for 43 % of the kernel: ~55k cycles per panel of
// LDS -> DFMA -> STS round trips that the compiler cannot overlap (possible aliasing), a DSQRT + DDIV per
// column and a DDIV per row of the inverse.  Here lane r (and its mirror r + 16) keeps row r of the block in
// registers; column c is a left-looking dot product against row c, fetched with width-16 shuffles; the
// pivot is broadcast, inverted once with rsqrt (no division anywhere); the inverse W = L^-1 is built column
// per lane from broadcast reads of L in shared memory.  W's strict lower part is parked in the unused
// strict UPPER triangle of the diagonal block in global memory (A[k0+j][k0+i] = W[i][j], i > j) so that
// the back substitution can apply W^H instead of running a serial triangular solve.
__device__ __forceinline__ void chol_diag_factor_reg(const Dims& d, cplx* A, int ld, int k0, int nb, cplx* sD, cplx* sW,
                                                     int lane, double& maxpiv, int* s_bad) {
    const unsigned full = 0xffffffffu;
    const int r = lane & 15;
    cplx a[CH_NB];
    {
        const cplx* rowp = A + (size_t)(k0 + min(r, nb - 1)) * ld + k0;
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) a[c] = (r < nb && c <= r && c < nb) ? rowp[c] : mk(0.0, 0.0);
    }
    double invs[CH_NB];
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) {
        cplx s = a[c];
#pragma unroll
        for (int q = 0; q < c; ++q) {
            const cplx lcq = mk(__shfl_sync(full, a[q].x, c, 16), __shfl_sync(full, a[q].y, c, 16));
            cfmsc(s, a[q], lcq);   // s -= L[r][q] conj(L[c][q])
        }
        double piv = __shfl_sync(full, s.x, c, 16);
        const bool live = c < nb;
        if (live) {
            // numerically singular: non-positive, or below 1e-13 of the largest pivot so far; identity padding exempt
            if (!(piv > ((k0 + c < d.L) ? 1e-13 * maxpiv : 0.0))) {
                if (lane == 0) *s_bad = 1;
                piv = 1.0;
            }
            maxpiv = fmax(maxpiv, piv);
        } else {
            piv = 1.0;
        }
        const double inv = rsqrt(piv);
        invs[c] = live ? inv : 0.0;
        a[c] = (!live || r < c) ? mk(0.0, 0.0) : (r == c ? mk(piv * inv, 0.0) : cscale(s, inv));
    }
    if (lane < CH_NB) {
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) sD[r * CH_DS + c] = a[c];
    }
    __syncwarp();
    // W = L^-1, lane = column j:  W[i][j] = -inv_i sum_{q<i} L[i][q] W[q][j]  (W[q][j] = 0 for q < j)
    cplx w[CH_NB];
#pragma unroll
    for (int i = 0; i < CH_NB; ++i) {
        cplx acc0 = mk(0.0, 0.0), acc1 = mk(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < i; ++q) {
            if (q & 1) cfma(acc1, sD[i * CH_DS + q], w[q]); else cfma(acc0, sD[i * CH_DS + q], w[q]);
        }
        w[i] = (i == r) ? mk(invs[i], 0.0) : mk(-invs[i] * (acc0.x + acc1.x), -invs[i] * (acc0.y + acc1.y));
    }
    if (lane < CH_NB) {
#pragma unroll
        for (int i = 0; i < CH_NB; ++i) sW[i * CH_DS + r] = w[i];
        if (r < nb) {   // row k0 + r of the block: L up to the diagonal, then W^T
            cplx* rowp = A + (size_t)(k0 + r) * ld + k0;
#pragma unroll
            for (int c = 0; c < CH_NB; ++c)
                if (c < nb) rowp[c] = (c <= r) ? sD[r * CH_DS + c] : w[c];
        }
    }
    __syncwarp();
}

template <int CH_THREADS, int CH_MINB>
__global__ void __launch_bounds__(CH_THREADS, CH_MINB) k_chol2(Dims d, cplx* __restrict__ Gall,
                                                               cplx* __restrict__ theta,
                                                               const int32_t* __restrict__ active,
                                                               int32_t* __restrict__ stat, cplx* th_global) {
    extern __shared__ double2 csm[];
    const int b = blockIdx.x;
    if (active != nullptr && active[b] == 0) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int Lp = d.Lp, Ltot = d.Ltot, ld = d.Lp;
    cplx* A = Gall + (size_t)b * Ltot * Lp;

    cplx* sD = csm;                       // [16][17] diagonal block being factored (later: back substitution)
    cplx* sWb = sD + CH_NB * CH_DS;       // [2][16][17] inverse of the diagonal block, double buffered
    cplx* th = th_global ? th_global + (size_t)b * d.Lp * d.n_rx : sWb + 2 * CH_NB * CH_DS;
    __shared__ int s_bad, s_next[2];
    __shared__ volatile int s_ready;      // panels whose diagonal block ROW is final (written by warp 0)
    if (tid == 0) { s_bad = 0; s_next[0] = 1; s_next[1] = 1; s_ready = 0; }
    double maxpiv = 0.0;
    const int npan = (Lp + CH_NB - 1) / CH_NB;
    if (warp == 0) chol_diag_factor_reg(d, A, ld, 0, min(CH_NB, Lp), sD, sWb, lane, maxpiv, &s_bad);
    __syncthreads();

    auto claim = [&](int* counter) {
        int v = 0;
        if (lane == 0) v = atomicAdd(counter, 1);
        return __shfl_sync(0xffffffffu, v, 0);
    };

    for (int k = 0; k < npan; ++k) {
        const int k0 = k * CH_NB;
        const int nb = min(CH_NB, Lp - k0);
        const int g0 = k0 + nb;                               // first row below the diagonal block = next panel
        const int nbn = (k + 1 < npan) ? min(CH_NB, Lp - g0) : 0;
        const int ntile = (Ltot - g0 + 15) >> 4;
        const cplx* sW = sWb + (k & 1) * CH_NB * CH_DS;
        if (tid == 0) s_next[(k + 1) & 1] = 1;                // next iteration's counter (idle during this one)

        auto tile = [&](int t) {
            const int r0 = g0 + (t << 4);
            const int ra = min(r0 + g, Ltot - 1), rb8 = min(r0 + g + 8, Ltot - 1);
            double cr[2][4], ci[2][4];
            // ---- step A: X = S W^H on columns k0 .. k0+nb-1
            {
                const cplx* pa0 = A + (size_t)ra * ld + k0 + tig;
                const cplx* pa1 = A + (size_t)rb8 * ld + k0 + tig;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) { cr[j][e] = 0.0; ci[j][e] = 0.0; }
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const int q0 = 8 * kk + tig;
                    cplx a0 = mk(0, 0), a1 = mk(0, 0), a2 = mk(0, 0), a3 = mk(0, 0);
                    if (q0 < nb) { a0 = pa0[8 * kk]; a1 = pa1[8 * kk]; }
                    if (q0 + 4 < nb) { a2 = pa0[8 * kk + 4]; a3 = pa1[8 * kk + 4]; }
                    const double ar[4] = {a0.x, a1.x, a2.x, a3.x};
                    const double ai[4] = {a0.y, a1.y, a2.y, a3.y};
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const cplx w0 = sW[(8 * j + g) * CH_DS + 8 * kk + tig];
                        const cplx w1 = sW[(8 * j + g) * CH_DS + 8 * kk + tig + 4];
                        dmma16x8x8(cr[j], ar, w0.x, w1.x);
                        dmma16x8x8(cr[j], ai, w0.y, w1.y);
                        dmma16x8x8(ci[j], ai, w0.x, w1.x);
                        dmma16x8x8(ci[j], ar, -w0.y, -w1.y);
                    }
                }
                __syncwarp();  // all lanes have read S before anyone overwrites it
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int row = r0 + g + 8 * h;
                        const int c = 8 * j + 2 * tig;
                        if (row < Ltot && c < nb) {
                            cplx* p2 = A + (size_t)row * ld + k0 + c;
                            p2[0] = mk(cr[j][2 * h], ci[j][2 * h]);
                            p2[1] = mk(cr[j][2 * h + 1], ci[j][2 * h + 1]);
                        }
                    }
            }
            if (nbn == 0) return;
            __threadfence_block();
            __syncwarp();
            if (t == 0) {
                if (lane == 0) s_ready = k + 1;               // block row g0.. is final in columns < g0
            } else {
                while (s_ready < k + 1) __nanosleep(32);
                __threadfence_block();
            }
            // ---- step B: S' = A[rows, g0:g0+nbn] - C[rows, 0:g0] C[g0:g0+16, 0:g0]^H
            {
                const cplx* pa0 = A + (size_t)ra * ld + tig;
                const cplx* pa1 = A + (size_t)rb8 * ld + tig;
                const cplx* pb0 = A + (size_t)min(g0 + g, Ltot - 1) * ld + tig;
                const cplx* pb1 = A + (size_t)min(g0 + 8 + g, Ltot - 1) * ld + tig;
                // the block to be updated is only needed at the very end: start fetching it now
                if (tig == 0) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)ra * ld + g0));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)ra * ld + g0 + 8));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)rb8 * ld + g0));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)rb8 * ld + g0 + 8));
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) { cr[j][e] = 0.0; ci[j][e] = 0.0; }
                cplx fa[4], fb[4];
                fa[0] = pa0[0]; fa[1] = pa1[0]; fa[2] = pa0[4]; fa[3] = pa1[4];
                fb[0] = pb0[0]; fb[1] = pb0[4]; fb[2] = pb1[0]; fb[3] = pb1[4];
#pragma unroll 1
                for (int q0 = 0; q0 < g0; q0 += 8) {
                    cplx na[4], nbq[4];
                    const int qn = (q0 + 8 < g0) ? q0 + 8 : q0;   // last step reloads itself (harmless, L1 hit)
                    na[0] = pa0[qn]; na[1] = pa1[qn]; na[2] = pa0[qn + 4]; na[3] = pa1[qn + 4];
                    nbq[0] = pb0[qn]; nbq[1] = pb0[qn + 4]; nbq[2] = pb1[qn]; nbq[3] = pb1[qn + 4];
                    const double ar[4] = {fa[0].x, fa[1].x, fa[2].x, fa[3].x};
                    const double ai[4] = {fa[0].y, fa[1].y, fa[2].y, fa[3].y};
                    // sum_q a conj(b):  re += ar br + ai bi ;  im += ai br - ar bi
                    dmma16x8x8(cr[0], ar, fb[0].x, fb[1].x);
                    dmma16x8x8(cr[0], ai, fb[0].y, fb[1].y);
                    dmma16x8x8(ci[0], ai, fb[0].x, fb[1].x);
                    dmma16x8x8(ci[0], ar, -fb[0].y, -fb[1].y);
                    dmma16x8x8(cr[1], ar, fb[2].x, fb[3].x);
                    dmma16x8x8(cr[1], ai, fb[2].y, fb[3].y);
                    dmma16x8x8(ci[1], ai, fb[2].x, fb[3].x);
                    dmma16x8x8(ci[1], ar, -fb[2].y, -fb[3].y);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { fa[e] = na[e]; fb[e] = nbq[e]; }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int row = r0 + g + 8 * h;
                        const int c = 8 * j + 2 * tig;
                        if (row < Ltot && c < nbn) {
                            cplx* p2 = A + (size_t)row * ld + g0 + c;
                            const cplx v0 = p2[0], v1 = p2[1];
                            p2[0] = mk(v0.x - cr[j][2 * h], v0.y - ci[j][2 * h]);
                            p2[1] = mk(v1.x - cr[j][2 * h + 1], v1.y - ci[j][2 * h + 1]);
                        }
                    }
            }
        };

        if (warp == 0) {
            tile(0);
            if (nbn > 0) {
                __threadfence_block();
                __syncwarp();
                chol_diag_factor_reg(d, A, ld, g0, nbn, sD, sWb + ((k + 1) & 1) * CH_NB * CH_DS, lane, maxpiv, &s_bad);
            }
        }
        for (int t = claim(&s_next[k & 1]); t < ntile; t = claim(&s_next[k & 1])) tile(t);
        __syncthreads();   // W_{k+1} is ready, every tile of this iteration is written
    }
    if (tid == 0 && s_bad && stat) atomicOr(&stat[b], SBCE_ST_NOT_PD);

    // ---- back substitution  C^H theta = z,  z[l][r] = conj(A[Lp + r][l])  (as in k_chol)
    const int nrx = d.n_rx;
    for (int e = tid; e < Lp * nrx; e += CH_THREADS) {
        const int l = e / nrx, r = e % nrx;
        th[e] = cconj(A[(size_t)(Lp + r) * ld + l]);
    }
    for (int k0 = ((Lp - 1) / CH_NB) * CH_NB; k0 >= 0; k0 -= CH_NB) {
        const int nb = min(CH_NB, Lp - k0);
        __syncthreads();  // th updates of the previous block are complete
        // D^H x = rhs  <=>  x = W^H rhs with W = D^-1 parked in the block's strict upper triangle:
        // x[c] = rhs[c] / D[c][c] + sum_{q>c} conj(W[q][c]) rhs[q], one thread per (c, right-hand side)
        const bool act = tid < nb * nrx;
        const int bc = act ? tid / nrx : 0, br = act ? tid % nrx : 0;
        cplx xv = mk(0.0, 0.0);
        if (act) {
            const cplx* wrow = A + (size_t)(k0 + bc) * ld + k0;
            xv = cscale(th[(k0 + bc) * nrx + br], 1.0 / wrow[bc].x);
            for (int q = bc + 1; q < nb; ++q) cfmac(xv, th[(k0 + q) * nrx + br], wrow[q]);
        }
        __syncthreads();
        if (act) th[(k0 + bc) * nrx + br] = xv;
        __syncthreads();
        // th[c][:] -= sum_q conj(C[k0+q][c]) x[q][:]   for all c < k0
        for (int c = tid; c < k0; c += CH_THREADS) {
            cplx cq[CH_NB];
#pragma unroll
            for (int q = 0; q < CH_NB; ++q) cq[q] = (q < nb) ? A[(size_t)(k0 + q) * ld + c] : mk(0.0, 0.0);
            for (int r = 0; r < nrx; ++r) {
                cplx v = th[c * nrx + r];
#pragma unroll
                for (int q = 0; q < CH_NB; ++q)
                    if (q < nb) cfmsc(v, th[(k0 + q) * nrx + r], cq[q]);
                th[c * nrx + r] = v;
            }
        }
    }
    __syncthreads();
    cplx* out = theta + (size_t)b * d.L * nrx;
    bool bad = false;
    for (int e = tid; e < d.L * nrx; e += CH_THREADS) {
        const cplx v = th[e];
        out[e] = v;
        if (!isfinite(v.x) || !isfinite(v.y)) bad = true;
    }
    if (bad && stat) atomicOr(&stat[b], SBCE_ST_NONFINITE);
}

template <int T, int MB>
static cudaError_t run_chol2(const Dims& d, int nb, double* G, double* theta, const int32_t* active, int32_t* stat,
                             size_t smem, double* thg, cudaStream_t s) {
    static SmemOptIn optin;
    cudaError_t e = opt_in_smem(optin, (const void*)k_chol2<T, MB>, smem);
    if (e != cudaSuccess) return e;
    k_chol2<T, MB><<<nb, T, smem, s>>>(d, (cplx*)G, (cplx*)theta, active, stat, (cplx*)thg);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_chol_solve(const Dims& d, int nb, double* G, double* theta, const int32_t* active, int32_t* stat,
                              double* th_scratch, cudaStream_t s) {
    size_t thsz = (size_t)d.Lp * d.n_rx;
    size_t smem = sizeof(cplx) * (3 * CH_NB * CH_DS + thsz);
    double* thg = nullptr;
    if (smem > 48 * 1024) {   // keep four trials resident per SM: the solution vector moves to global scratch
        if (!th_scratch) return cudaErrorInvalidValue;
        thg = th_scratch;
        smem = sizeof(cplx) * (3 * CH_NB * CH_DS);
    }
    // CTA shapes measured on B200 (N=64, 4x4: L=260, 592 trials) for the look-ahead kernel:
    // 128 threads x 4 CTAs/SM 0.98 ms, 96 x 5 1.27 ms, 256 x 2 1.57 ms, 160 x 3 1.45 ms, 128 x 5 (96 registers)
    // 1.11 ms -- latency bound, more resident trials per SM win even though their factors no longer fit in L2.
    return run_chol2<128, 4>(d, nb, G, theta, active, stat, smem, thg, s);
}

}  // namespace sbce
