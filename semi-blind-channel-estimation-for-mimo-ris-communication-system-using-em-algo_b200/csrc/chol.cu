// Solve step of the M-step: blocked complex Cholesky of the augmented lower trapezoid [G ; B^H] produced by the
// normal-equation kernels (mstep.cu) and the back substitution, one CTA per trial.
//
// Reference semantics (/root/reference/Proposed_method_NMSEvsTp.py:66):
//   theta = np.linalg.solve(denom, numer)   with denom = G (x) I_nrx  =>  Theta = G^-1 B   (L x n_rx)
#include <math.h>

#include "common.cuh"
#include "tensor.cuh"

namespace sbce {

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
                                                               int32_t* __restrict__ stat, cplx* th_global,
                                                               int pf_cols) {
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
                    if (pf_cols > 0 && tig == 0 && q0 + pf_cols < g0) {   // pull the lines of a later step into L2
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pa0 + q0 + pf_cols));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pa1 + q0 + pf_cols));
                    }
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
    k_chol2<T, MB><<<nb, T, smem, s>>>(d, (cplx*)G, (cplx*)theta, active, stat, (cplx*)thg, 0);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// k_chol3: the same look-ahead left-looking factorisation as a BARRIER-FREE task pipeline.
//
// What ncu showed on k_chol2 (profiles/r01m, 1184 trials): DMMA pipe 53 % active; of the warp-cycles not
// issuing, 35 % wait on global loads (register double buffering looks one 8-column step ahead, the operand
// stream misses L2: 1184 factors = 650 MB), 21 % sit at the per-panel CTA barrier (tile counts per panel are
// not multiples of the warp count and warp 0 carries the diagonal factor), 17 % on back-to-back dependent DMMAs.
// Here
//  * the (panel k, row tile t) pairs form ONE ordered task list; warps claim tasks from a shared counter and
//    synchronise through flags in shared memory only -- progress per 16-row block (s_prog), "block row k+1 is
//    final" (s_rowready), "W_k is available" (s_wready) -- so a warp that runs out of tiles of panel k
//    starts on panel k+1 as soon as its inputs exist.  Every wait is on a task EARLIER in the list and tasks
//    are claimed in order, so the earliest unfinished task can always run: no deadlock.  The warp that owns
//    tile (k, 0) factors the next diagonal block; the inverses W_k live in a 3-deep ring (a writer checks that
//    all tiles of panel k-2 are past their triangular solve).
//  * both operands of the update  S' = A - C[rows, 0:g0] C[g0:g0+16, 0:g0]^H  stream through a per-warp
//    cp.async ring in shared memory (16 x 8 complex chunks, XOR-swizzled so that the fragment loads are
//    conflict-free): loads run NSTAGE-1 steps ahead without holding registers, which frees them for
//  * the 3-multiplication complex product (GAUSS): re = P1 + P2, im = P3 - P1 + P2 with P1 = ar br,
//    P2 = ai bi, P3 = (ar + ai)(br - bi): 6 instead of 8 DMMA quads per step for 8 extra DADDs.
// With GAUSS = false the per-accumulator operation order equals k_chol2's: results are bit-identical.
// ---------------------------------------------------------------------------
constexpr int C3_WARPS = 4;
constexpr int C3_CHUNK = 16 * 8;   // complex elements of one staged operand chunk: 16 rows x 8 columns

__device__ __forceinline__ void spin_until_ge(const volatile int* p, int v) {
    while (*p < v) __nanosleep(20);
    __threadfence_block();
}

template <int NSTAGE, bool GAUSS, int MINB>
__global__ void __launch_bounds__(C3_WARPS * 32, MINB) k_chol3(Dims d, cplx* __restrict__ Gall,
                                                               cplx* __restrict__ theta,
                                                               const int32_t* __restrict__ active,
                                                               int32_t* __restrict__ stat, cplx* th_global) {
    static_assert(NSTAGE == 0 || NSTAGE >= 2, "the triangular solve stages its 16 x 16 tile in two ring chunks");
    constexpr bool RING = NSTAGE > 0;      // false: operand fragments straight from global / L1 (as k_chol2)
    constexpr int CH_THREADS = C3_WARPS * 32;
    extern __shared__ __align__(16) double2 csm[];
    const int b = blockIdx.x;
    if (active != nullptr && active[b] == 0) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int Lp = d.Lp, Ltot = d.Ltot, ld = d.Lp;
    cplx* A = Gall + (size_t)b * Ltot * Lp;
    const int npan = (Lp + CH_NB - 1) / CH_NB;
    const int nrb = (Ltot + 15) / 16 + 1;

    cplx* sD = csm;                                            // [16][17] diagonal block scratch
    cplx* sWb = sD + CH_NB * CH_DS;                            // [3][16][17] inverses of the diagonal blocks
    cplx* ring = sWb + 3 * CH_NB * CH_DS;                      // [warps][NSTAGE][A chunk, B chunk]; then the solution
    const int ring_elems = RING ? C3_WARPS * NSTAGE * 2 * C3_CHUNK : (th_global ? 0 : Lp * d.n_rx);
    int* s_prog = (int*)(ring + ring_elems);                   // [nrb] panels completed per aligned 16-row block
    int* s_adone = s_prog + nrb;                               // [npan] tiles of panel k past their triangular solve
    int* s_start = s_adone + npan;                             // [npan + 1] first task id of panel k
    __shared__ int s_next, s_bad;
    __shared__ volatile int s_wready, s_rowready;
    __shared__ double s_maxpiv;

    for (int i = tid; i < nrb + npan; i += CH_THREADS) s_prog[i] = 0;
    if (tid == 0) {
        s_next = 0; s_bad = 0; s_wready = 0; s_rowready = 0; s_maxpiv = 0.0;
        int acc = 0;
        for (int k = 0; k < npan; ++k) {
            s_start[k] = acc;
            const int g0 = min(k * CH_NB + CH_NB, Lp);
            acc += (Ltot - g0 + 15) >> 4;
        }
        s_start[npan] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        double maxpiv = 0.0;
        chol_diag_factor_reg(d, A, ld, 0, min(CH_NB, Lp), sD, sWb, lane, maxpiv, &s_bad);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) { s_maxpiv = maxpiv; s_wready = 1; }
    }
    __syncthreads();
    const int ntasks = s_start[npan];

    // lane-constant pieces of the staging maps
    const int l3 = lane >> 3, col = lane & 7;                     // cp.async: rows l3 + 4 i, column `col` of a chunk
    const int dst0 = l3 * 8 + (col ^ ((l3 & 1) << 2));            // + 32 i
    const int sw = (g & 1) << 2;
    const int oa0 = g * 8 + (tig ^ sw), oa2 = g * 8 + ((tig + 4) ^ sw);   // fragment elements (row g, col tig / tig + 4)
    cplx* myring = ring + (size_t)warp * NSTAGE * 2 * C3_CHUNK;

    int myk = 0;
    for (;;) {
        int id = 0;
        if (lane == 0) id = atomicAdd(&s_next, 1);
        id = __shfl_sync(0xffffffffu, id, 0);
        if (id >= ntasks) break;
        while (id >= s_start[myk + 1]) ++myk;
        const int k = myk, t = id - s_start[myk];
        const int k0 = k * CH_NB;
        const int nb = min(CH_NB, Lp - k0);
        const int g0 = k0 + nb;                                   // first row below the diagonal block = next panel
        const int nbn = (k + 1 < npan) ? min(CH_NB, Lp - g0) : 0;
        const int r0 = g0 + (t << 4);
        const cplx* sW = sWb + (k % 3) * CH_NB * CH_DS;

        // rows of this tile must have been taken through panels 0 .. k-1; W_k must exist
        if (k > 0) {
            spin_until_ge((const volatile int*)&s_prog[r0 >> 4], k);
            spin_until_ge((const volatile int*)&s_prog[(min(r0 + 16, Ltot) - 1) >> 4], k);
        }
        spin_until_ge(&s_wready, k + 1);

        int rowA[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rowA[i] = min(r0 + l3 + 4 * i, Ltot - 1);
        double cr[2][4], ci[2][4];
        // ---- step A: X = S W^H on columns k0 .. k0+nb-1
        {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { cr[j][e] = 0.0; ci[j][e] = 0.0; }
            if constexpr (RING) {   // S staged in ring stages 0 and 1
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    cplx* sA = myring + kk * 2 * C3_CHUNK;
                    const int c = 8 * kk + col;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        cp_async16_zfill(sA + dst0 + 32 * i, A + (size_t)rowA[i] * ld + k0 + min(c, nb - 1), c < nb);
                }
                cp_async_commit();
                cp_async_wait<0>();
                __syncwarp();
            }
            const int ra = min(r0 + g, Ltot - 1), rb8 = min(r0 + g + 8, Ltot - 1);
            const cplx* pa0 = A + (size_t)ra * ld + k0 + tig;
            const cplx* pa1 = A + (size_t)rb8 * ld + k0 + tig;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                cplx a0 = mk(0, 0), a1 = mk(0, 0), a2 = mk(0, 0), a3 = mk(0, 0);
                if constexpr (RING) {
                    const cplx* sA = myring + kk * 2 * C3_CHUNK;
                    a0 = sA[oa0]; a1 = sA[oa0 + 64]; a2 = sA[oa2]; a3 = sA[oa2 + 64];
                } else {
                    const int q0 = 8 * kk + tig;
                    if (q0 < nb) { a0 = pa0[8 * kk]; a1 = pa1[8 * kk]; }
                    if (q0 + 4 < nb) { a2 = pa0[8 * kk + 4]; a3 = pa1[8 * kk + 4]; }
                }
                const double ar[4] = {a0.x, a1.x, a2.x, a3.x};
                const double ai[4] = {a0.y, a1.y, a2.y, a3.y};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const cplx w0 = sW[(8 * j + g) * CH_DS + 8 * kk + tig];
                    const cplx w1 = sW[(8 * j + g) * CH_DS + 8 * kk + tig + 4];
                    dmma16x8x8(cr[j], ar, w0.x, w1.x);
                    dmma16x8x8(ci[j], ai, w0.x, w1.x);
                    dmma16x8x8(cr[j], ai, w0.y, w1.y);
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
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atomicAdd(&s_adone[k], 1);
        if (nbn == 0) continue;                                   // last panel: nothing left to update
        if (t == 0) {
            if (lane == 0) s_rowready = k + 1;                    // block row g0.. is final in columns < g0
        } else {
            spin_until_ge(&s_rowready, k + 1);
        }
        // ---- step B: S' = A[rows, g0:g0+nbn] - C[rows, 0:g0] C[g0:g0+16, 0:g0]^H
        {
            int rowB[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) rowB[i] = min(g0 + l3 + 4 * i, Ltot - 1);
            const int ra = min(r0 + g, Ltot - 1), rb8 = min(r0 + g + 8, Ltot - 1);
            if (tig == 0) {   // the block to be updated is only needed at the very end: start fetching it now
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)ra * ld + g0));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)ra * ld + g0 + 8));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)rb8 * ld + g0));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)rb8 * ld + g0 + 8));
            }
            const int nq = g0 >> 3;                               // 8-column steps (g0 is a multiple of 16 here)
            double p3[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { cr[j][e] = 0.0; ci[j][e] = 0.0; p3[j][e] = 0.0; }
            // one 8-column step on the fragments (a0..a3: rows g / g+8 x columns tig / tig+4 of the tile;
            // b00, b01 / b10, b11: the same columns of block rows g and 8+g)
            auto mma_step = [&](const cplx& a0, const cplx& a1, const cplx& a2, const cplx& a3, const cplx& b00,
                                const cplx& b01, const cplx& b10, const cplx& b11) {
                const double ar[4] = {a0.x, a1.x, a2.x, a3.x};
                const double ai[4] = {a0.y, a1.y, a2.y, a3.y};
                if (GAUSS) {
                    // cr = sum ar br, ci = sum ai bi, p3 = sum (ar + ai)(br - bi)
                    const double as[4] = {a0.x + a0.y, a1.x + a1.y, a2.x + a2.y, a3.x + a3.y};
                    dmma16x8x8(cr[0], ar, b00.x, b01.x);
                    dmma16x8x8(ci[0], ai, b00.y, b01.y);
                    dmma16x8x8(p3[0], as, b00.x - b00.y, b01.x - b01.y);
                    dmma16x8x8(cr[1], ar, b10.x, b11.x);
                    dmma16x8x8(ci[1], ai, b10.y, b11.y);
                    dmma16x8x8(p3[1], as, b10.x - b10.y, b11.x - b11.y);
                } else {
                    // sum_q a conj(b):  re += ar br + ai bi ;  im += ai br - ar bi
                    dmma16x8x8(cr[0], ar, b00.x, b01.x);
                    dmma16x8x8(ci[0], ai, b00.x, b01.x);
                    dmma16x8x8(cr[1], ar, b10.x, b11.x);
                    dmma16x8x8(ci[1], ai, b10.x, b11.x);
                    dmma16x8x8(cr[0], ai, b00.y, b01.y);
                    dmma16x8x8(ci[0], ar, -b00.y, -b01.y);
                    dmma16x8x8(cr[1], ai, b10.y, b11.y);
                    dmma16x8x8(ci[1], ar, -b10.y, -b11.y);
                }
            };
            if constexpr (RING) {
                auto issue = [&](int q) {
                    cplx* sA = myring + (q % NSTAGE) * 2 * C3_CHUNK;
                    cplx* sB = sA + C3_CHUNK;
                    const int c = 8 * q + col;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        cp_async16_ca(sA + dst0 + 32 * i, A + (size_t)rowA[i] * ld + c);
                        cp_async16_ca(sB + dst0 + 32 * i, A + (size_t)rowB[i] * ld + c);
                    }
                };
#pragma unroll
                for (int q = 0; q < NSTAGE - 1; ++q) {
                    if (q < nq) issue(q);
                    cp_async_commit();
                }
#pragma unroll 1
                for (int q = 0; q < nq; ++q) {
                    if (q + NSTAGE - 1 < nq) issue(q + NSTAGE - 1);
                    cp_async_commit();
                    cp_async_wait<(NSTAGE > 0 ? NSTAGE - 1 : 0)>();
                    __syncwarp();
                    const cplx* sA = myring + (q % NSTAGE) * 2 * C3_CHUNK;
                    const cplx* sB = sA + C3_CHUNK;
                    mma_step(sA[oa0], sA[oa0 + 64], sA[oa2], sA[oa2 + 64], sB[oa0], sB[oa2], sB[oa0 + 64], sB[oa2 + 64]);
                    __syncwarp();   // every lane is done with this stage before a later issue() refills it
                }
                cp_async_wait<0>();
            } else {
                // fragments straight from global memory, double-buffered in registers (the loads of step q+1 are
                // in flight during the DMMAs of step q); the block row hits L1, the tile rows stream from L2
                const cplx* pa0 = A + (size_t)ra * ld + tig;
                const cplx* pa1 = A + (size_t)rb8 * ld + tig;
                const cplx* pb0 = A + (size_t)min(g0 + g, Ltot - 1) * ld + tig;
                const cplx* pb1 = A + (size_t)min(g0 + 8 + g, Ltot - 1) * ld + tig;
                cplx fa[4], fb[4];
                fa[0] = pa0[0]; fa[1] = pa1[0]; fa[2] = pa0[4]; fa[3] = pa1[4];
                fb[0] = pb0[0]; fb[1] = pb0[4]; fb[2] = pb1[0]; fb[3] = pb1[4];
#pragma unroll 1
                for (int q0 = 0; q0 < g0; q0 += 8) {
                    cplx na[4], nbq[4];
                    const int qn = (q0 + 8 < g0) ? q0 + 8 : q0;   // last step reloads itself (harmless, L1 hit)
                    na[0] = pa0[qn]; na[1] = pa1[qn]; na[2] = pa0[qn + 4]; na[3] = pa1[qn + 4];
                    nbq[0] = pb0[qn]; nbq[1] = pb0[qn + 4]; nbq[2] = pb1[qn]; nbq[3] = pb1[qn + 4];
                    mma_step(fa[0], fa[1], fa[2], fa[3], fb[0], fb[1], fb[2], fb[3]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { fa[e] = na[e]; fb[e] = nbq[e]; }
                }
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
                        double re0, im0, re1, im1;
                        if (GAUSS) {
                            re0 = cr[j][2 * h] + ci[j][2 * h];
                            im0 = p3[j][2 * h] - cr[j][2 * h] + ci[j][2 * h];
                            re1 = cr[j][2 * h + 1] + ci[j][2 * h + 1];
                            im1 = p3[j][2 * h + 1] - cr[j][2 * h + 1] + ci[j][2 * h + 1];
                        } else {
                            re0 = cr[j][2 * h]; im0 = ci[j][2 * h]; re1 = cr[j][2 * h + 1]; im1 = ci[j][2 * h + 1];
                        }
                        p2[0] = mk(v0.x - re0, v0.y - im0);
                        p2[1] = mk(v1.x - re1, v1.y - im1);
                    }
                }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *((volatile int*)&s_prog[r0 >> 4]) = k + 1;
        if (t == 0) {
            // next diagonal block: its inverse goes into the ring slot last read by the tiles of panel k-2
            if (k >= 2) spin_until_ge((const volatile int*)&s_adone[k - 2], s_start[k - 1] - s_start[k - 2]);
            double maxpiv = s_maxpiv;
            chol_diag_factor_reg(d, A, ld, g0, nbn, sD, sWb + ((k + 1) % 3) * CH_NB * CH_DS, lane, maxpiv, &s_bad);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) { s_maxpiv = maxpiv; s_wready = k + 2; }
        }
    }
    __syncthreads();
    if (tid == 0 && s_bad && stat) atomicOr(&stat[b], SBCE_ST_NOT_PD);

    // ---- back substitution  C^H theta = z,  z[l][r] = conj(A[Lp + r][l])  (as in k_chol2; the solution vector
    // lives in the operand ring, idle by now, or in global scratch when it is longer than the ring)
    cplx* th = th_global ? th_global + (size_t)b * d.Lp * d.n_rx : ring;   // ring is idle by now
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

template <int NSTAGE, bool GAUSS, int MINB>
static cudaError_t run_chol3(const Dims& d, int nb, double* G, double* theta, const int32_t* active, int32_t* stat,
                             double* th_scratch, cudaStream_t s) {
    static SmemOptIn optin;
    const int npan = (d.Lp + CH_NB - 1) / CH_NB, nrb = (d.Ltot + 15) / 16 + 1;
    const size_t fixed = sizeof(cplx) * (size_t)(4 * CH_NB * CH_DS) + sizeof(int) * (size_t)(nrb + 2 * npan + 2);
    const size_t thb = sizeof(cplx) * (size_t)d.Lp * d.n_rx;
    size_t ring = sizeof(cplx) * (size_t)(C3_WARPS * NSTAGE * 2 * C3_CHUNK);
    double* thg = nullptr;
    if (NSTAGE > 0) {
        if (thb > ring) thg = th_scratch;        // solution vector longer than the (idle) operand ring
    } else {
        // register-fed operands: the solution vector sits in shared memory while MINB trials still fit an SM
        // with room left for L1 (the block-row operand lives there), else in global scratch
        if (fixed + thb > 24 * 1024) thg = th_scratch; else ring = thb;
    }
    if ((NSTAGE > 0 ? thb > ring : fixed + thb > 24 * 1024) && !th_scratch) return cudaErrorInvalidValue;
    const size_t smem = fixed + ring;
    cudaError_t e = opt_in_smem(optin, (const void*)k_chol3<NSTAGE, GAUSS, MINB>, smem);
    if (e != cudaSuccess) return e;
    k_chol3<NSTAGE, GAUSS, MINB><<<nb, C3_WARPS * 32, smem, s>>>(d, (cplx*)G, (cplx*)theta, active, stat, (cplx*)thg);
    count_launch();
    return cudaGetLastError();
}

// L2 cache-hinted 16-byte accesses (createpolicy + ld/st.global.L2::cache_hint)
__device__ __forceinline__ unsigned long long l2_policy(bool keep) {
    unsigned long long pol;
    if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <int L2POL>
__device__ __forceinline__ cplx ldg_c(const cplx* p, unsigned long long pol) {
    if constexpr (L2POL == 0) {
        return *p;
    } else {
        cplx v;
        asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
        return v;
    }
}
template <int L2POL>
__device__ __forceinline__ void stg_c(cplx* p, cplx v, unsigned long long pol) {
    if constexpr (L2POL == 0) {
        *p = v;
    } else {
        asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
    }
}

// ---------------------------------------------------------------------------
// k_chol4: task pipeline of k_chol3 with TWO PANELS PER PASS over the factor.
//
// Measured (gpurun_out/r02b-r02c, 1184 trials): k_chol2 1.94 ms; k_chol3 with both operands through a cp.async
// ring 2.1-2.8 ms (the ring costs L1: the block-row operand then misses); k_chol3 with register-fed operands
// 2.03 ms, 1.89 ms with the 3-multiplication product.  Neither barriers nor DMMA count were the limiter: every
// panel streams all previous columns of the rows below it from L2/DRAM (4.7 GB per launch, 7x the factor).
// Halving that stream is what this kernel does: on EVEN panels k the update step takes the tile through
// panels k+1 AND k+2 at once,
//   S'[rows, panel k+1 | panel k+2] = A - C[rows, 0:g0] C[block rows k+1 | k+2, 0:g0]^H     (32 columns wide)
// so the A fragments (the streamed operand) feed 16 instead of 8 DMMA quads per 8-column step; on ODD panels
// only the missing 16-column slice is applied to panel k+2:  S' -= C[rows, k0:g0] C[block row, k0:g0]^H.
// Same flags as k_chol3 plus "block row k+2 is final left of g0" (s_rowready2, set by tile 1).
// ---------------------------------------------------------------------------
template <bool GAUSS, int MINB, int PF, bool WIDE, bool BSPF, int L2POL>
__global__ void __launch_bounds__(C3_WARPS * 32, MINB) k_chol4(Dims d, cplx* __restrict__ Gall,
                                                               cplx* __restrict__ theta,
                                                               const int32_t* __restrict__ active,
                                                               int32_t* __restrict__ stat, cplx* th_global) {
    constexpr int CH_THREADS = C3_WARPS * 32;
    extern __shared__ __align__(16) double2 csm[];
    const int b = blockIdx.x;
    if (active != nullptr && active[b] == 0) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int Lp = d.Lp, Ltot = d.Ltot, ld = d.Lp;
    cplx* A = Gall + (size_t)b * Ltot * Lp;
    const int npan = (Lp + CH_NB - 1) / CH_NB;
    const int nrb = (Ltot + 15) / 16 + 1;
    // L2POL 1: the factors of every third trial are kept in L2 (evict_last), the others stream through it
    // (evict_first): a cyclic working set of ~590 resident factors (340 MB) thrashes a 126 MB LRU completely,
    // pinning a third of it turns a third of the operand stream into L2 hits.  L2POL 2: everything evict_first
    // except the block rows (control experiment).
    const unsigned long long pol = L2POL ? l2_policy(L2POL == 1 && (blockIdx.x % 3) == 0) : 0ull;

    cplx* sD = csm;                                            // [16][17] diagonal block scratch
    cplx* sWb = sD + CH_NB * CH_DS;                            // [3][16][17] inverses of the diagonal blocks
    cplx* sTh = sWb + 3 * CH_NB * CH_DS;                       // [Lp][n_rx] solution vector (unless in global scratch)
    int* s_prog = (int*)(sTh + (th_global ? 0 : Lp * d.n_rx)); // [nrb] panels completed per aligned 16-row block
    int* s_adone = s_prog + nrb;                               // [npan] tiles of panel k past their triangular solve
    int* s_start = s_adone + npan;                             // [npan + 1] first task id of panel k
    __shared__ int s_next, s_bad;
    __shared__ volatile int s_wready, s_rowready, s_rowready2;
    __shared__ double s_maxpiv;

    for (int i = tid; i < nrb + npan; i += CH_THREADS) s_prog[i] = 0;
    if (tid == 0) {
        s_next = 0; s_bad = 0; s_wready = 0; s_rowready = 0; s_rowready2 = 0; s_maxpiv = 0.0;
        int acc = 0;
        for (int k = 0; k < npan; ++k) {
            s_start[k] = acc;
            const int g0 = min(k * CH_NB + CH_NB, Lp);
            acc += (Ltot - g0 + 15) >> 4;
        }
        s_start[npan] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        double maxpiv = 0.0;
        chol_diag_factor_reg(d, A, ld, 0, min(CH_NB, Lp), sD, sWb, lane, maxpiv, &s_bad);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) { s_maxpiv = maxpiv; s_wready = 1; }
    }
    __syncthreads();
    const int ntasks = s_start[npan];

    int myk = 0;
    for (;;) {
        int id = 0;
        if (lane == 0) id = atomicAdd(&s_next, 1);
        id = __shfl_sync(0xffffffffu, id, 0);
        if (id >= ntasks) break;
        while (id >= s_start[myk + 1]) ++myk;
        const int k = myk, t = id - s_start[myk];
        const int k0 = k * CH_NB;
        const int nb = min(CH_NB, Lp - k0);
        const int g0 = k0 + nb;                                   // first row below the diagonal block = next panel
        const int nbn = (k + 1 < npan) ? min(CH_NB, Lp - g0) : 0;
        const bool wide = WIDE && (k & 1) == 0;
        // width of panel k+2 where this tile updates it as well (tile 0 sits above that panel's diagonal block)
        const int nbn2 = (wide && t >= 1 && nbn == CH_NB && k + 2 < npan) ? min(CH_NB, Lp - (g0 + CH_NB)) : 0;
        const int r0 = g0 + (t << 4);
        const cplx* sW = sWb + (k % 3) * CH_NB * CH_DS;

        if (k > 0) {
            spin_until_ge((const volatile int*)&s_prog[r0 >> 4], k);
            spin_until_ge((const volatile int*)&s_prog[(min(r0 + 16, Ltot) - 1) >> 4], k);
        }
        spin_until_ge(&s_wready, k + 1);

        const int ra = min(r0 + g, Ltot - 1), rb8 = min(r0 + g + 8, Ltot - 1);
        double cr[4][4], ci[4][4];
        // ---- step A: X = S W^H on columns k0 .. k0+nb-1
        {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { cr[j][e] = 0.0; ci[j][e] = 0.0; }
            const cplx* pa0 = A + (size_t)ra * ld + k0 + tig;
            const cplx* pa1 = A + (size_t)rb8 * ld + k0 + tig;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const int q0 = 8 * kk + tig;
                cplx a0 = mk(0, 0), a1 = mk(0, 0), a2 = mk(0, 0), a3 = mk(0, 0);
                if (q0 < nb) { a0 = ldg_c<L2POL>(pa0 + 8 * kk, pol); a1 = ldg_c<L2POL>(pa1 + 8 * kk, pol); }
                if (q0 + 4 < nb) { a2 = ldg_c<L2POL>(pa0 + 8 * kk + 4, pol); a3 = ldg_c<L2POL>(pa1 + 8 * kk + 4, pol); }
                const double ar[4] = {a0.x, a1.x, a2.x, a3.x};
                const double ai[4] = {a0.y, a1.y, a2.y, a3.y};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const cplx w0 = sW[(8 * j + g) * CH_DS + 8 * kk + tig];
                    const cplx w1 = sW[(8 * j + g) * CH_DS + 8 * kk + tig + 4];
                    dmma16x8x8(cr[j], ar, w0.x, w1.x);
                    dmma16x8x8(ci[j], ai, w0.x, w1.x);
                    dmma16x8x8(cr[j], ai, w0.y, w1.y);
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
                        stg_c<L2POL>(p2, mk(cr[j][2 * h], ci[j][2 * h]), pol);
                        stg_c<L2POL>(p2 + 1, mk(cr[j][2 * h + 1], ci[j][2 * h + 1]), pol);
                    }
                }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atomicAdd(&s_adone[k], 1);
        if (nbn == 0) continue;                                   // last panel: nothing left to update
        if (lane == 0) {
            if (t == 0) s_rowready = k + 1;                       // block row k+1 is final in columns < g0
            // block row k+2 likewise; tiles (k, 1) of successive panels may finish out of order on the odd
            // (narrow) panels, where nobody waits for them: keep the flag monotonic
            if (t == 1) atomicMax((int*)&s_rowready2, k + 1);
        }
        if (t >= 1) spin_until_ge(&s_rowready, k + 1);
        if (nbn2 > 0 && t >= 2) spin_until_ge(&s_rowready2, k + 1);
        // ---- step B
        {
            const int qbeg = (WIDE && !wide) ? k0 : 0;            // odd panels: only the 16 columns of panel k are missing
            const bool four = nbn2 > 0;                           // warp-uniform: panel k+2 is updated too
            if (tig == 0) {   // the blocks to be updated are only needed at the very end: start fetching them now
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)ra * ld + g0));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)rb8 * ld + g0));
            }
            double p3[GAUSS ? 4 : 1][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { cr[j][e] = 0.0; ci[j][e] = 0.0; if (GAUSS) p3[GAUSS ? j : 0][e] = 0.0; }
            const cplx* pa0 = A + (size_t)ra * ld + tig;
            const cplx* pa1 = A + (size_t)rb8 * ld + tig;
            const cplx* pb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) pb[j] = A + (size_t)min(g0 + 8 * j + g, Ltot - 1) * ld + tig;
            // fragments are double-buffered in registers: the loads of step q+1 are in flight during the DMMAs of
            // step q (the tile rows stream from L2 / DRAM, the block rows hit L1)
            constexpr int NJ = WIDE ? 4 : 2;
            cplx fa[4], fb[NJ][2];
            fa[0] = ldg_c<L2POL>(pa0 + qbeg, pol); fa[1] = ldg_c<L2POL>(pa1 + qbeg, pol);
            fa[2] = ldg_c<L2POL>(pa0 + qbeg + 4, pol); fa[3] = ldg_c<L2POL>(pa1 + qbeg + 4, pol);
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (j < 2 || four) { fb[j][0] = pb[j][qbeg]; fb[j][1] = pb[j][qbeg + 4]; }
            auto pair = [&](int j, const double (&ar)[4], const double (&ai)[4], const double (&as)[4], const cplx& b0,
                            const cplx& b1) {
                if (GAUSS) {
                    dmma16x8x8(cr[j], ar, b0.x, b1.x);
                    dmma16x8x8(ci[j], ai, b0.y, b1.y);
                    dmma16x8x8(p3[GAUSS ? j : 0], as, b0.x - b0.y, b1.x - b1.y);
                } else {
                    dmma16x8x8(cr[j], ar, b0.x, b1.x);
                    dmma16x8x8(ci[j], ai, b0.x, b1.x);
                    dmma16x8x8(cr[j], ai, b0.y, b1.y);
                    dmma16x8x8(ci[j], ar, -b0.y, -b1.y);
                }
            };
            // L1 prefetch of the streamed operand PF steps ahead (no registers held): lanes 0-15 touch the line
            // that holds the first, lanes 16-31 the line that holds the last element of their row's 128-byte segment
            const cplx* ppf = A + (size_t)min(r0 + (lane & 15), Ltot - 1) * ld + ((lane >> 4) ? 7 : 0);
#pragma unroll 1
            for (int q0 = qbeg; q0 < g0; q0 += 8) {
                cplx na[4], nbf[NJ][2];
                const int qn = (q0 + 8 < g0) ? q0 + 8 : q0;   // last step reloads itself (harmless, L1 hit)
                if (PF > 0 && q0 + 8 * PF < g0) asm volatile("prefetch.global.L1 [%0];" ::"l"(ppf + q0 + 8 * PF));
                na[0] = ldg_c<L2POL>(pa0 + qn, pol); na[1] = ldg_c<L2POL>(pa1 + qn, pol);
                na[2] = ldg_c<L2POL>(pa0 + qn + 4, pol); na[3] = ldg_c<L2POL>(pa1 + qn + 4, pol);
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    if (j < 2 || four) { nbf[j][0] = pb[j][qn]; nbf[j][1] = pb[j][qn + 4]; }
                const double ar[4] = {fa[0].x, fa[1].x, fa[2].x, fa[3].x};
                const double ai[4] = {fa[0].y, fa[1].y, fa[2].y, fa[3].y};
                const double as[4] = {fa[0].x + fa[0].y, fa[1].x + fa[1].y, fa[2].x + fa[2].y, fa[3].x + fa[3].y};
                pair(0, ar, ai, as, fb[0][0], fb[0][1]);
                pair(1, ar, ai, as, fb[1][0], fb[1][1]);
                if (WIDE && four) {
                    pair(NJ - 2, ar, ai, as, fb[NJ - 2][0], fb[NJ - 2][1]);
                    pair(NJ - 1, ar, ai, as, fb[NJ - 1][0], fb[NJ - 1][1]);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) fa[e] = na[e];
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    if (j < 2 || four) { fb[j][0] = nbf[j][0]; fb[j][1] = nbf[j][1]; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j >= 2 && !four) break;
                const int cbase = g0 + 8 * j;                     // panel k+1: columns g0 .. ; panel k+2: g0 + 16 ..
                const int width = (j < 2) ? g0 + nbn : g0 + CH_NB + nbn2;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = r0 + g + 8 * h;
                    const int c = cbase + 2 * tig;
                    if (row < Ltot && c < width) {
                        cplx* p2 = A + (size_t)row * ld + c;
                        const cplx v0 = ldg_c<L2POL>(p2, pol), v1 = ldg_c<L2POL>(p2 + 1, pol);
                        double re0, im0, re1, im1;
                        if (GAUSS) {
                            re0 = cr[j][2 * h] + ci[j][2 * h];
                            im0 = p3[GAUSS ? j : 0][2 * h] - cr[j][2 * h] + ci[j][2 * h];
                            re1 = cr[j][2 * h + 1] + ci[j][2 * h + 1];
                            im1 = p3[GAUSS ? j : 0][2 * h + 1] - cr[j][2 * h + 1] + ci[j][2 * h + 1];
                        } else {
                            re0 = cr[j][2 * h]; im0 = ci[j][2 * h]; re1 = cr[j][2 * h + 1]; im1 = ci[j][2 * h + 1];
                        }
                        stg_c<L2POL>(p2, mk(v0.x - re0, v0.y - im0), pol);
                        stg_c<L2POL>(p2 + 1, mk(v1.x - re1, v1.y - im1), pol);
                    }
                }
            }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *((volatile int*)&s_prog[r0 >> 4]) = k + 1;
        if (t == 0) {
            // next diagonal block: its inverse goes into the ring slot last read by the tiles of panel k-2
            if (k >= 2) spin_until_ge((const volatile int*)&s_adone[k - 2], s_start[k - 1] - s_start[k - 2]);
            double maxpiv = s_maxpiv;
            chol_diag_factor_reg(d, A, ld, g0, nbn, sD, sWb + ((k + 1) % 3) * CH_NB * CH_DS, lane, maxpiv, &s_bad);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) { s_maxpiv = maxpiv; s_wready = k + 2; }
        }
    }
    __syncthreads();
    if (tid == 0 && s_bad && stat) atomicOr(&stat[b], SBCE_ST_NOT_PD);

    // ---- back substitution  C^H theta = z,  z[l][r] = conj(A[Lp + r][l])  (as in k_chol2).  The factor rows it
    // walks were written long ago (592 resident factors do not fit L2): with BSPF the 16 rows of the NEXT block
    // are pulled into L2 while the current block is processed.
    cplx* th = th_global ? th_global + (size_t)b * d.Lp * d.n_rx : sTh;
    const int nrx = d.n_rx;
    for (int e = tid; e < Lp * nrx; e += CH_THREADS) {
        const int l = e / nrx, r = e % nrx;
        th[e] = cconj(A[(size_t)(Lp + r) * ld + l]);
    }
    for (int k0 = ((Lp - 1) / CH_NB) * CH_NB; k0 >= 0; k0 -= CH_NB) {
        const int nb = min(CH_NB, Lp - k0);
        if (BSPF && k0 >= CH_NB) {   // rows k0-16 .. k0-1, columns 0 .. k0-1: one 128-byte line per request
            const int lines = (k0 * (int)sizeof(cplx) + 127) >> 7;
            for (int e = tid; e < CH_NB * lines; e += CH_THREADS)
                asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)(A + (size_t)(k0 - CH_NB + e / lines) * ld) +
                                                               ((size_t)(e % lines) << 7)));
        }
        __syncthreads();  // th updates of the previous block are complete
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

template <bool GAUSS, int MINB, int PF, bool WIDE, bool BSPF, int L2POL = 0>
static cudaError_t run_chol4(const Dims& d, int nb, double* G, double* theta, const int32_t* active, int32_t* stat,
                             double* th_scratch, cudaStream_t s) {
    static SmemOptIn optin;
    const int npan = (d.Lp + CH_NB - 1) / CH_NB, nrb = (d.Ltot + 15) / 16 + 1;
    const size_t fixed = sizeof(cplx) * (size_t)(4 * CH_NB * CH_DS) + sizeof(int) * (size_t)(nrb + 2 * npan + 2);
    const size_t thb = sizeof(cplx) * (size_t)d.Lp * d.n_rx;
    // the solution vector sits in shared memory while that leaves most of the SM's L1 to the block-row operand
    double* thg = nullptr;
    size_t smem = fixed + thb;
    if (smem > 24 * 1024) {
        if (!th_scratch) return cudaErrorInvalidValue;
        thg = th_scratch;
        smem = fixed;
    }
    cudaError_t e = opt_in_smem(optin, (const void*)k_chol4<GAUSS, MINB, PF, WIDE, BSPF, L2POL>, smem);
    if (e != cudaSuccess) return e;
    k_chol4<GAUSS, MINB, PF, WIDE, BSPF, L2POL><<<nb, C3_WARPS * 32, smem, s>>>(d, (cplx*)G, (cplx*)theta, active, stat, (cplx*)thg);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_chol_solve(const Dims& d, int nb, double* G, double* theta, const int32_t* active, int32_t* stat,
                              double* th_scratch, cudaStream_t s) {
#ifdef SBCE_DEV
    // tuning builds only (libsbce_dev.so, tools/build_dev.sh): pick the Cholesky variant at run time
    switch (dev_knob("SBCE_CHOL", 0)) {
        case 2: break;   // k_chol2 below
        case 30: return run_chol3<2, false, 4>(d, nb, G, theta, active, stat, th_scratch, s);
        case 31: return run_chol3<2, true, 4>(d, nb, G, theta, active, stat, th_scratch, s);
        case 32: return run_chol3<3, false, 3>(d, nb, G, theta, active, stat, th_scratch, s);
        case 33: return run_chol3<3, true, 3>(d, nb, G, theta, active, stat, th_scratch, s);
        case 34: return run_chol3<2, true, 3>(d, nb, G, theta, active, stat, th_scratch, s);
        case 35: return run_chol3<4, true, 3>(d, nb, G, theta, active, stat, th_scratch, s);
        case 36: return run_chol3<0, false, 4>(d, nb, G, theta, active, stat, th_scratch, s);
        case 37: return run_chol3<0, true, 4>(d, nb, G, theta, active, stat, th_scratch, s);
        case 38: return run_chol3<0, true, 3>(d, nb, G, theta, active, stat, th_scratch, s);
        case 39: return run_chol3<0, false, 3>(d, nb, G, theta, active, stat, th_scratch, s);
        // k_chol4<GAUSS, MINB, PF, WIDE, BSPF>
        case 40: return run_chol4<false, 3, 0, true, false>(d, nb, G, theta, active, stat, th_scratch, s);
        case 41: return run_chol4<true, 3, 0, true, false>(d, nb, G, theta, active, stat, th_scratch, s);
        case 42: return run_chol4<false, 3, 2, true, true>(d, nb, G, theta, active, stat, th_scratch, s);
        case 43: return run_chol4<true, 3, 2, true, true>(d, nb, G, theta, active, stat, th_scratch, s);
        case 50: return run_chol4<true, 4, 0, false, false>(d, nb, G, theta, active, stat, th_scratch, s);   // = k_chol3<0,true,4>
        case 51: return run_chol4<true, 4, 0, false, true>(d, nb, G, theta, active, stat, th_scratch, s);
        case 52: return run_chol4<true, 4, 2, false, true>(d, nb, G, theta, active, stat, th_scratch, s);
        case 53: return run_chol4<true, 4, 4, false, true>(d, nb, G, theta, active, stat, th_scratch, s);
        case 54: return run_chol4<false, 4, 2, false, true>(d, nb, G, theta, active, stat, th_scratch, s);
        case 55: return run_chol4<true, 3, 2, false, true>(d, nb, G, theta, active, stat, th_scratch, s);
        case 56: return run_chol4<true, 4, 2, false, true, 1>(d, nb, G, theta, active, stat, th_scratch, s);
        case 57: return run_chol4<true, 4, 2, false, true, 2>(d, nb, G, theta, active, stat, th_scratch, s);
        case 58: return run_chol4<true, 4, 0, false, true, 1>(d, nb, G, theta, active, stat, th_scratch, s);
        default: break;
    }
#endif
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
