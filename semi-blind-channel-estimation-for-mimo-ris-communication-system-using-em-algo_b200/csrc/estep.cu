// E-step of the semi-blind EM estimator: posterior over all M^n_tx QAM
// hypotheses of every data symbol.
//
// Reference semantics (one data symbol t, /root/reference/Proposed_method_NMSEvsTp.py:53-62):
//   d2_k   = || y_t - Z(x_k, psi~_t) theta ||^2          for all K = M^n_tx hypotheses
//   beta_k = exp(-d2_k / varn^2) / sum_k' exp(-d2_k'/varn^2)
//   numer += beta_k Z^H y ,  denom += beta_k Z^H Z
// which collapses to per-symbol statistics
//   m_t = sum_k beta_k conj(x_k)          R_t = sum_k beta_k conj(x_k) x_k^T
// and, for the hard-decision variant (Proposed method/ML_detecctor.py:66-77),
//   k*  = argmax_k beta_k (first index on ties), rank-one statistics at x_{k*}.
//
// B200 design (not a translation of the double Python loop):
//  1. k_heff_qr: one LANE per data symbol.  The Kronecker design row
//     psi~_t^T (x) x^T (x) I is never formed; the lane contracts theta over the
//     RIS index into the n_rx x n_tx effective channel Heff_t and immediately
//     reduces [Heff_t | y_t] to upper-triangular form with Householder
//     reflections in registers (real non-negative diagonal).  Then
//     d2(x) = c0 + sum_i | ytilde_i - sum_{j>=i} R_ij x_j |^2  exactly.
//  2. k_enum: one WARP per data symbol walks the hypothesis tree stream
//     n_tx-1 -> 1 with partial residuals in registers; the top levels are split
//     across lanes, the inner levels are warp-uniform loops fed by shared-memory
//     tables R_ij*c_m.  Because R_00 is real, the last stream separates into its
//     in-phase and quadrature PAM components, so the sum over the M leaves of a
//     node is a product of two sqrt(M)-term sums: every hypothesis is accounted
//     for exactly, at O(1) work per node when the node's best leaf is more than
//     64 varn^2 above the running minimum (its whole weight is < 2e-28 of the
//     normaliser and rounds away in FP64, exactly as in the reference where
//     float(beta) underflows).  Online max-subtracted accumulation per lane,
//     warp-shuffle log-sum-exp merge at the end; the incomplete-data
//     log-likelihood sum_k exp(-d2/varn^2) falls out of the same reduction.
#include <math.h>

#include "common.cuh"
#include "tensor.cuh"

namespace sbce {

static __device__ __forceinline__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------
// pilots: symbols known with probability one
// ---------------------------------------------------------------------------
__global__ void k_pilot_stats(int n_tx, int total, const cplx* __restrict__ Xp, cplx* __restrict__ pm,
                              cplx* __restrict__ pR) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;  // flat (b,t)
    if (s >= total) return;
    const cplx* x = Xp + (size_t)s * n_tx;
    for (int i = 0; i < n_tx; ++i) {
        cplx xi = x[i];
        pm[(size_t)s * n_tx + i] = cconj(xi);
        for (int j = 0; j < n_tx; ++j) pR[((size_t)s * n_tx + i) * n_tx + j] = cmulc(x[j], xi);  // conj(x_i) x_j
    }
}

cudaError_t launch_pilot_stats(const Dims& d, int nb, const double* Xp, double* pil_m, double* pil_R, cudaStream_t s) {
    int total = nb * d.T_p;
    if (total == 0) return cudaSuccess;
    k_pilot_stats<<<(total + 127) / 128, 128, 0, s>>>(d.n_tx, total, (const cplx*)Xp, (cplx*)pil_m, (cplx*)pil_R);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// 1. effective channel + Householder QR, one lane per symbol
// ---------------------------------------------------------------------------
// Per-lane tail shared by the effective-channel kernels: A = Heff_t (rows >= NRX zero), then
// y' = y - Heff o_t (superimposed pilots), Householder QR of [A | y] in registers, record store.
template <int NTX, int NRX>
__device__ __forceinline__ void heff_qr_finish(const Dims& d, int b, int t, cplx (&A)[cmax(NTX, NRX)][NTX],
                                               const cplx* __restrict__ Yd, const cplx* __restrict__ Xoff,
                                               double* __restrict__ qr) {
    constexpr int NR = cmax(NTX, NRX);
    cplx y[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) y[r] = (r < NRX) ? Yd[((size_t)b * d.T_d + t) * NRX + r] : mk(0.0, 0.0);
    if (Xoff != nullptr) {   // superimposed pilots: hypotheses are x_k + o_t  <=>  y' = y - Heff o_t
        const cplx* o = Xoff + ((size_t)b * d.T_d + t) * NTX;
#pragma unroll
        for (int j = 0; j < NTX; ++j) {
            const cplx oj = o[j], noj = mk(-oj.x, -oj.y);
#pragma unroll
            for (int r = 0; r < NRX; ++r) cfma(y[r], A[r][j], noj);
        }
    }

    // Householder, column k; afterwards row k is rotated so that R_kk = ||x|| >= 0 is real
#pragma unroll
    for (int k = 0; k < NTX; ++k) {
        double nrm2 = 0.0;
#pragma unroll
        for (int i = k; i < NR; ++i) nrm2 += cnorm2(A[i][k]);
        if (nrm2 > 0.0) {
            const double nrm = sqrt(nrm2);
            const cplx a0 = A[k][k];
            const double abs0 = sqrt(cnorm2(a0));
            const cplx phase = abs0 > 0.0 ? mk(a0.x / abs0, a0.y / abs0) : mk(1.0, 0.0);
            // v = x - alpha e_k with alpha = -phase*nrm  =>  v_k = phase*(abs0+nrm)
            const cplx vk = cscale(phase, abs0 + nrm);
            const double beta = 1.0 / (nrm * (nrm + abs0));  // 2 / (v^H v)
#pragma unroll
            for (int c = k + 1; c <= NTX; ++c) {  // c == NTX is the y column
                cplx w = mk(0.0, 0.0);
                if (c < NTX) {
                    cfmac(w, A[k][c], vk);  // conj(v_k) * a
#pragma unroll
                    for (int i = k + 1; i < NR; ++i) cfmac(w, A[i][c], A[i][k]);
                    w = cscale(w, beta);
                    cplx nw = mk(-w.x, -w.y);
                    cfma(A[k][c], nw, vk);
#pragma unroll
                    for (int i = k + 1; i < NR; ++i) cfma(A[i][c], nw, A[i][k]);
                } else {
                    cfmac(w, y[k], vk);
#pragma unroll
                    for (int i = k + 1; i < NR; ++i) cfmac(w, y[i], A[i][k]);
                    w = cscale(w, beta);
                    cplx nw = mk(-w.x, -w.y);
                    cfma(y[k], nw, vk);
#pragma unroll
                    for (int i = k + 1; i < NR; ++i) cfma(y[i], nw, A[i][k]);
                }
            }
            // rotate row k by conj(-phase): diagonal becomes +nrm
            const cplx rot = mk(-phase.x, phase.y);
#pragma unroll
            for (int c = k + 1; c < NTX; ++c) A[k][c] = cmul(A[k][c], rot);
            y[k] = cmul(y[k], rot);
            A[k][k] = mk(nrm, 0.0);
        } else {
            A[k][k] = mk(0.0, 0.0);
        }
    }

    double* rec = qr + ((size_t)b * d.T_d + t) * d.rec;
    int o = 0;
#pragma unroll
    for (int i = 0; i < NTX; ++i)
#pragma unroll
        for (int j = i; j < NTX; ++j) {
            rec[o++] = A[i][j].x;
            rec[o++] = (j == i) ? 0.0 : A[i][j].y;
        }
#pragma unroll
    for (int i = 0; i < NTX; ++i) {
        rec[o++] = y[i].x;
        rec[o++] = y[i].y;
    }
    double c0 = 0.0;
#pragma unroll
    for (int i = NTX; i < NR; ++i) c0 += cnorm2(y[i]);
    rec[o++] = c0;
    rec[o++] = 0.0;
}

template <int NTX, int NRX>
__global__ void __launch_bounds__(128) k_heff_qr(Dims d, const cplx* __restrict__ Yd, const cplx* __restrict__ PsiD,
                                                 const cplx* __restrict__ theta, const int32_t* __restrict__ active,
                                                 const cplx* __restrict__ Xoff, double* __restrict__ qr, int use_smem) {
    constexpr int NR = cmax(NTX, NRX);
    extern __shared__ double2 heff_smem[];
    const int b = blockIdx.y;
    if (active != nullptr && active[b] == 0) return;
    // theta of the trial is read by every lane at warp-uniform addresses, 16 values per RIS index: staged in
    // shared memory once per CTA (coalesced) the inner loop issues LDS broadcasts instead of global loads
    // (the global-load queue was the kernel's top stall, profiles/r01m); use_smem = 0: theta too long
    const cplx* th = theta + (size_t)b * d.L * NRX;
    if (use_smem) {
        for (int e = threadIdx.x; e < d.L * NRX; e += blockDim.x) heff_smem[e] = th[e];
        __syncthreads();
        th = heff_smem;
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.T_d) return;

    const cplx* psi = PsiD + ((size_t)(d.psi_shared ? 0 : b) * d.T_d + t) * d.N1;

    cplx A[NR][NTX];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < NTX; ++j) A[r][j] = mk(0.0, 0.0);

    // Heff[r][j] = sum_n' psi~[t][n'] Theta[n'*n_tx + j][r]; theta reads are warp-uniform (broadcast)
#pragma unroll 4
    for (int n = 0; n < d.N1; ++n) {
        const cplx p = psi[n];
        const cplx* row = th + (size_t)n * NTX * NRX;
#pragma unroll
        for (int j = 0; j < NTX; ++j)
#pragma unroll
            for (int r = 0; r < NRX; ++r) cfma(A[r][j], p, row[j * NRX + r]);
    }
    heff_qr_finish<NTX, NRX>(d, b, t, A, Yd, Xoff, qr);
}

// ---------------------------------------------------------------------------
// 1a. the same on the FP64 tensor path.  Heff = Psi_d (T_d x N+1) . Theta (N+1 x n_tx n_rx) is a dense complex
// GEMM per trial (the RIS contraction, Proposed_method_NMSEvsTp.py:56 without the Kronecker products); the
// per-lane kernel above streams each lane's psi row with stride-(N+1) 16-byte loads (17 % of HBM peak, the
// load queue is its top stall, profiles/r01m).  Here a warp owns 32 symbols = two 16-row MMA tiles: A
// fragments (psi) are 128-byte row segments per 8-column step straight from global memory, B fragments
// (Theta, shared by the whole CTA) come from a padded shared-memory copy; the 2 x NT x 4 accumulators are
// transposed through shared memory so that lane t ends up with Heff_t in registers and runs the same
// Householder tail.
// ---------------------------------------------------------------------------
template <int NTX, int NRX>
__global__ void __launch_bounds__(128, 4) k_heff_qr_mma(Dims d, const cplx* __restrict__ Yd,
                                                        const cplx* __restrict__ PsiD, const cplx* __restrict__ theta,
                                                        const int32_t* __restrict__ active,
                                                        const cplx* __restrict__ Xoff, double* __restrict__ qr) {
    constexpr int NR = cmax(NTX, NRX);
    constexpr int NC = NTX * NRX;          // complex columns of the GEMM: (j, r) -> j * NRX + r
    constexpr int NT = (NC + 7) / 8;       // 8-column MMA tiles
    constexpr int TS = NT * 8 + 2;         // row stride of the staged Theta: fragment loads are conflict-free
    constexpr int HS = NC + 1;             // row stride of the transposed result
    extern __shared__ double2 heff_smem[];
    const int b = blockIdx.y;
    if (active != nullptr && active[b] == 0) return;
    const int N1 = d.N1, KP = (N1 + 7) & ~7;
    cplx* sTh = heff_smem;                 // [KP][TS], rows >= N1 and columns >= NC zero
    cplx* sH = sTh + KP * TS;              // [4 warps][32][HS]
    {
        const cplx* th = theta + (size_t)b * d.L * NRX;
        for (int e = threadIdx.x; e < KP * TS; e += blockDim.x) {
            const int n = e / TS, c = e % TS;
            sTh[e] = (n < N1 && c < NC) ? th[n * NC + c] : mk(0.0, 0.0);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const int t0 = blockIdx.x * 128 + warp * 32;
    if (t0 >= d.T_d) return;
    const cplx* psi_b = PsiD + (size_t)(d.psi_shared ? 0 : b) * d.T_d * N1;
    const cplx* prow[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) prow[mt][h] = psi_b + (size_t)min(t0 + mt * 16 + g + 8 * h, d.T_d - 1) * N1;
    {   // pull this warp's 32 phase rows into L2 before the k loop touches them (lane = row): the fragment loads
        // below are then L2 hits instead of DRAM round trips (0.200 -> 0.174 ms at the north-star size)
        const char* rowp = (const char*)(psi_b + (size_t)min(t0 + lane, d.T_d - 1) * N1);
        for (int o = 0; o < N1 * (int)sizeof(cplx); o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + o));
    }
    double cr[2][NT][4], ci[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) { cr[mt][nt][e] = 0.0; ci[mt][nt][e] = 0.0; }
#pragma unroll 2
    for (int k0 = 0; k0 < KP; k0 += 8) {
        const int ka = k0 + tig, kb = ka + 4;
        const bool va = ka < N1, vb = kb < N1;
        double ar[2][4], ai[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            // A fragment order: (row g, k lo), (row g+8, k lo), (row g, k hi), (row g+8, k hi)
            const cplx a0 = va ? prow[mt][0][ka] : mk(0.0, 0.0), a1 = va ? prow[mt][1][ka] : mk(0.0, 0.0);
            const cplx a2 = vb ? prow[mt][0][kb] : mk(0.0, 0.0), a3 = vb ? prow[mt][1][kb] : mk(0.0, 0.0);
            ar[mt][0] = a0.x; ar[mt][1] = a1.x; ar[mt][2] = a2.x; ar[mt][3] = a3.x;
            ai[mt][0] = a0.y; ai[mt][1] = a1.y; ai[mt][2] = a2.y; ai[mt][3] = a3.y;
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const cplx b0 = sTh[ka * TS + 8 * nt + g], b1 = sTh[kb * TS + 8 * nt + g];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                // (ar + i ai)(br + i bi): re += ar br - ai bi ; im += ar bi + ai br
                dmma16x8x8(cr[mt][nt], ar[mt], b0.x, b1.x);
                dmma16x8x8(ci[mt][nt], ar[mt], b0.y, b1.y);
                dmma16x8x8(cr[mt][nt], ai[mt], -b0.y, -b1.y);
                dmma16x8x8(ci[mt][nt], ai[mt], b0.x, b1.x);
            }
        }
    }
    // transpose through shared memory: accumulator (row g + 8h, columns 8 nt + 2 tig, + 1) -> H[symbol][column]
    cplx* H = sH + (size_t)warp * 32 * HS;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = mt * 16 + g + 8 * h, c = 8 * nt + 2 * tig;
                if (c < NC) H[row * HS + c] = mk(cr[mt][nt][2 * h], ci[mt][nt][2 * h]);
                if (c + 1 < NC) H[row * HS + c + 1] = mk(cr[mt][nt][2 * h + 1], ci[mt][nt][2 * h + 1]);
            }
    __syncwarp();
    const int t = t0 + lane;
    if (t >= d.T_d) return;
    cplx A[NR][NTX];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < NTX; ++j) A[r][j] = (r < NRX) ? H[lane * HS + j * NRX + r] : mk(0.0, 0.0);
    heff_qr_finish<NTX, NRX>(d, b, t, A, Yd, Xoff, qr);
}

// ---------------------------------------------------------------------------
// 1b. wide arrays (n_tx = 5..8): the (8 x (n_tx+1)) work matrix [Heff_t | y_t] no longer fits the
// registers of one lane, so EIGHT lanes share a symbol, lane r owning receive row r.  The theta reads of
// the RIS contraction are then 8 consecutive complex values per (n', j) -- one 128-byte line per symbol
// group; the Householder inner products become 3-step xor-shuffle sums inside the aligned 8-lane group.
// All lanes execute every shuffle (no group-divergent branches); same record layout as k_heff_qr.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double group8_sum(double v) {
    v += shfl_xor_d(v, 4);
    v += shfl_xor_d(v, 2);
    v += shfl_xor_d(v, 1);
    return v;
}

template <int NTX>
__global__ void __launch_bounds__(128) k_heff_qr_rows(Dims d, const cplx* __restrict__ Yd,
                                                      const cplx* __restrict__ PsiD, const cplx* __restrict__ theta,
                                                      const int32_t* __restrict__ active,
                                                      const cplx* __restrict__ Xoff, double* __restrict__ qr) {
    const int b = blockIdx.y;
    if (active != nullptr && active[b] == 0) return;
    const int r = threadIdx.x & 7;
    const int tq = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
    const bool tv = tq < d.T_d;
    const int t = tv ? tq : d.T_d - 1;   // surplus groups redo the last symbol (they must stay in the shuffles)
    const int nrx = d.n_rx;
    const bool rv = r < nrx;
    const int rr = rv ? r : 0;

    const cplx* psi = PsiD + ((size_t)(d.psi_shared ? 0 : b) * d.T_d + t) * d.N1;
    const cplx* th = theta + (size_t)b * d.L * nrx;

    cplx A[NTX + 1];   // row r of [Heff | y]
#pragma unroll
    for (int j = 0; j <= NTX; ++j) A[j] = mk(0.0, 0.0);
#pragma unroll 2
    for (int n = 0; n < d.N1; ++n) {
        const cplx p = psi[n];
        const cplx* row = th + (size_t)n * NTX * nrx + rr;
#pragma unroll
        for (int j = 0; j < NTX; ++j) cfma(A[j], p, __ldg(&row[j * nrx]));
    }
    A[NTX] = Yd[((size_t)b * d.T_d + t) * nrx + rr];
    if (Xoff != nullptr) {   // superimposed pilots: y' = y - Heff o_t
        const cplx* o = Xoff + ((size_t)b * d.T_d + t) * NTX;
#pragma unroll
        for (int j = 0; j < NTX; ++j) {
            const cplx oj = o[j];
            cfma(A[NTX], A[j], mk(-oj.x, -oj.y));
        }
    }
    if (!rv) {
#pragma unroll
        for (int j = 0; j <= NTX; ++j) A[j] = mk(0.0, 0.0);
    }

#pragma unroll
    for (int k = 0; k < NTX; ++k) {
        const double nrm2 = group8_sum(r >= k ? cnorm2(A[k]) : 0.0);
        const bool ok = nrm2 > 0.0;
        const double nrm = sqrt(nrm2);
        const int src = (threadIdx.x & 24) + k;   // lane that owns row k of this group
        const cplx a0 = mk(shfl_d(A[k].x, src), shfl_d(A[k].y, src));
        const double abs0 = sqrt(cnorm2(a0));
        const cplx phase = abs0 > 0.0 ? mk(a0.x / abs0, a0.y / abs0) : mk(1.0, 0.0);
        const cplx vk = cscale(phase, abs0 + nrm);                      // v_k = phase (|a0| + ||x||)
        const double beta = ok ? 1.0 / (nrm * (nrm + abs0)) : 0.0;      // 2 / (v^H v)
        const cplx vr = (r == k) ? vk : (r > k ? A[k] : mk(0.0, 0.0));  // this lane's entry of v
#pragma unroll
        for (int c = k + 1; c <= NTX; ++c) {
            cplx w = mk(0.0, 0.0);
            cfmac(w, A[c], vr);   // conj(v_r) a_rc
            w = mk(group8_sum(w.x) * beta, group8_sum(w.y) * beta);
            const cplx nw = mk(-w.x, -w.y);
            cfma(A[c], nw, vr);
        }
        if (r == k) {   // rotate row k so that its diagonal entry is +||x||
            const cplx rot = mk(-phase.x, phase.y);
#pragma unroll
            for (int c = k + 1; c <= NTX; ++c) A[c] = cmul(A[c], rot);
            A[k] = mk(ok ? nrm : 0.0, 0.0);
        }
    }
    const double c0 = group8_sum(r >= NTX ? cnorm2(A[NTX]) : 0.0);
    if (!tv) return;
    double* rec = qr + ((size_t)b * d.T_d + t) * d.rec;
    if (r < NTX) {
        double* o = rec + 2 * (r * NTX - (r * (r - 1)) / 2);
#pragma unroll
        for (int j = 0; j < NTX; ++j)
            if (j >= r) {
                o[2 * (j - r)] = A[j].x;
                o[2 * (j - r) + 1] = (j == r) ? 0.0 : A[j].y;
            }
        rec[NTX * (NTX + 1) + 2 * r] = A[NTX].x;
        rec[NTX * (NTX + 1) + 2 * r + 1] = A[NTX].y;
    }
    if (r == 0) {
        rec[NTX * (NTX + 1) + 2 * NTX] = c0;
        rec[NTX * (NTX + 1) + 2 * NTX + 1] = 0.0;
    }
}

// ---------------------------------------------------------------------------
// 2. hypothesis-tree enumeration, one warp per symbol
// ---------------------------------------------------------------------------
constexpr double SBCE_THR = 64.0;  // nodes whose best leaf is > THR*varn^2 above the incumbent weigh < e^-64
constexpr int ENUM_QCAP = 1024;    // per-warp candidate queue (node codes)

template <int NTX, int SQM>
struct EnumT {
    static constexpr int M = SQM * SQM;
    static constexpr int BITS = (SQM == 2 ? 2 : (SQM == 4 ? 4 : 6));
    static constexpr int HB = BITS / 2;
    static constexpr int NODE_STREAMS = NTX - 1;
    static constexpr int PLWANT = (5 + BITS - 1) / BITS;  // prefix streams so that M^PL >= 32
    static constexpr int PL = NODE_STREAMS < PLWANT ? NODE_STREAMS : PLWANT;
    static constexpr int NPREF = 1 << (BITS * PL);
    static constexpr int NPAIR = NTX * (NTX - 1) / 2;
    static constexpr int NPAIR1 = NPAIR > 0 ? NPAIR : 1;
    static constexpr int NACC = 1 + 3 * NTX + 2 * NPAIR;  // S, m (re,im), R diag, R upper (re,im)
    // per-warp shared block (doubles): tables, PAM levels, ytilde, c0, accumulators, aref, then ints
    static constexpr int O_TAB = 0;
    static constexpr int O_G = O_TAB + 2 * NPAIR1 * M;
    static constexpr int O_YT = O_G + NTX * SQM;
    static constexpr int O_C0 = O_YT + 2 * NTX;
    static constexpr int O_ACC = O_C0 + 1;
    static constexpr int O_AREF = O_ACC + NACC;
    static constexpr int O_S2 = O_AREF + 1;     // inv_s2
    static constexpr int O_CNT = O_S2 + 1;      // int counter lives in this double slot
    static constexpr int O_Q = O_CNT + 1;       // ENUM_QCAP ints = ENUM_QCAP/2 doubles
    static constexpr int O_SCR = O_Q + ENUM_QCAP / 2;   // [NACC][32] reduction scratch of the flush
    static constexpr int WS_DOUBLES = ((O_SCR + NACC * 32) + 1) & ~1;

    __device__ __forceinline__ static double pam(int a) { return (double)(2 * a - SQM + 1); }
    __device__ __forceinline__ static cplx cval(int m) { return mk(pam(m & (SQM - 1)), pam(m >> HB)); }
    __device__ __forceinline__ static int pair(int i, int s) { return s * (s - 1) / 2 + i; }

    // nearest PAM level to v among gs[0..SQM) (increasing), first index on ties; dist = squared distance
    __device__ __forceinline__ static int slice(double v, const double* gs, double& dist) {
        int a = 0;
        double e = v - gs[0];
        dist = e * e;
#pragma unroll
        for (int q = 1; q < SQM; ++q) {
            e = v - gs[q];
            const double dq = e * e;
            if (dq < dist) { dist = dq; a = q; }
        }
        return a;
    }

    // distance from v to the nearest level of the symmetric PAM set {+-r, +-3r, ...}: fold |v| around the
    // midpoints (for 4 levels: ||v| - 2r| - r), no comparisons, 1 + log2(SQM/2) subtractions
    __device__ __forceinline__ static double fold(double v, double r) {
        double a = fabs(v);
        if (SQM == 8) a = fabs(a - 4.0 * r);
        if (SQM >= 4) a = fabs(a - 2.0 * r);
        return a - r;
    }
};

// Cooperative, out-of-line processing of the queued candidate nodes of one symbol: every lane takes
// queue entries round-robin, rebuilds the node's residuals from its code, sums its M leaves in closed
// form (separable in-phase / quadrature PAM sums) and the warp adds the result to the shared
// accumulators, kept relative to the reference distance `aref` (rescaled when the incumbent improves).
template <int NTX, int SQM>
__device__ __noinline__ void enum_flush(double* ws, double warp_best) {
    typedef EnumT<NTX, SQM> E;
    constexpr int M = E::M;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    int* cntp = (int*)(ws + E::O_CNT);
    const int n = min(*cntp, ENUM_QCAP);
    const int* q = (const int*)(ws + E::O_Q);
    const cplx* tab = (const cplx*)(ws + E::O_TAB);
    const double* g = ws + E::O_G;
    const double inv_s2 = ws[E::O_S2];
    double aref = ws[E::O_AREF];
    if (warp_best < aref) {
        const double f = exp((warp_best - aref) * inv_s2);  // aref = +inf initially -> f = 0
        for (int i = lane; i < E::NACC; i += 32) ws[E::O_ACC + i] *= f;
        aref = warp_best;
    }
    double* A = ws + E::O_ACC;
    // Narrow path.  At operating SNRs the queue almost always holds ONE node (the arg-min node, whose M leaves
    // carry the whole posterior) or two; the general path below would run all 32 lanes through ~700
    // instructions (9 exponentials in sequence per lane) for it.  Here the warp shares one entry: the lanes of
    // a 2*SQM group take one PAM level of the in-phase or of the quadrature component each (one exponential
    // per lane, min / sums by xor shuffles), every lane then holds the node's closed-form leaf sums and lane i
    // adds statistic i.
    if (n <= 2) {
        constexpr int LG = (SQM == 2 ? 1 : (SQM == 4 ? 2 : 3));
        const int a = lane & (SQM - 1);
        const bool isQ = (lane >> LG) & 1;
        const double pq = E::pam(a);
        for (int e = 0; e < n; ++e) {
            const int code = q[e];
            double base = ws[E::O_C0];
            cplx t0;
            {
                cplx acc[NTX];
#pragma unroll
                for (int i = 0; i < NTX; ++i) acc[i] = mk(ws[E::O_YT + 2 * i], ws[E::O_YT + 2 * i + 1]);
#pragma unroll
                for (int s = NTX - 1; s >= 1; --s) {
                    const int m = (code >> (E::BITS * (NTX - 1 - s))) & (M - 1);
                    const double eI = acc[s].x - g[s * SQM + (m & (SQM - 1))];
                    const double eQ = acc[s].y - g[s * SQM + (m >> E::HB)];
                    base += fma(eI, eI, eQ * eQ);
#pragma unroll
                    for (int i = 0; i < NTX; ++i)
                        if (i < s) acc[i] = csub(acc[i], tab[E::pair(i, s) * M + m]);
                }
                t0 = acc[0];
            }
            const double ev = (isQ ? t0.y : t0.x) - g[a];
            const double dd = ev * ev;
            double mn = dd;
#pragma unroll
            for (int o = SQM / 2; o > 0; o >>= 1) mn = fmin(mn, shfl_xor_d(mn, o));
            const double w = exp((mn - dd) * inv_s2);
            double s0 = w, s1 = pq * w, s2 = pq * pq * w;
#pragma unroll
            for (int o = SQM / 2; o > 0; o >>= 1) {
                s0 += shfl_xor_d(s0, o);
                s1 += shfl_xor_d(s1, o);
                s2 += shfl_xor_d(s2, o);
            }
            const double o0 = shfl_xor_d(s0, SQM), o1 = shfl_xor_d(s1, SQM), o2 = shfl_xor_d(s2, SQM);
            const double om = shfl_xor_d(mn, SQM);
            const double EI = isQ ? o0 : s0, A1 = isQ ? o1 : s1, A2 = isQ ? o2 : s2, minI = isQ ? om : mn;
            const double EQ = isQ ? s0 : o0, B1 = isQ ? s1 : o1, B2 = isQ ? s2 : o2, minQ = isQ ? mn : om;
            const double W = exp((aref - (base + minI + minQ)) * inv_s2);
            const double E0 = W * EI * EQ;
            const cplx F0 = mk(W * A1 * EQ, -W * EI * B1);  // sum over leaves of e * conj(x_0)
            const double Q0 = W * (A2 * EQ + EI * B2);      // sum over leaves of e * |x_0|^2
            // statistic idx belongs to lane idx & 31: every lane picks its values out of the (warp-uniform) results
            // with predicated register moves and then makes ONE shared-memory update per 32 statistics (as ~25
            // single-lane read-modify-writes behind divergent branches this was 14 % of the kernel's stall
            // samples, profiles/r02m); per entry, so the order of the additions is unchanged
            constexpr int NR = (E::NACC + 31) / 32;
            double mine[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) mine[r] = 0.0;
            auto add = [&](int idx, double v) { if ((idx & 31) == lane) mine[idx >> 5] = v; };
            add(0, E0);
            add(1, F0.x);
            add(1 + NTX, F0.y);
            add(1 + 2 * NTX, Q0);
#pragma unroll
            for (int s = 1; s < NTX; ++s) {
                const cplx xs = E::cval((code >> (E::BITS * (NTX - 1 - s))) & (M - 1));
                add(1 + s, E0 * xs.x);
                add(1 + NTX + s, -E0 * xs.y);
                add(1 + 2 * NTX + s, E0 * cnorm2(xs));
                const cplx v = cmul(F0, xs);  // conj(x_0) x_s summed over the leaves
                add(1 + 3 * NTX + 2 * E::pair(0, s), v.x);
                add(1 + 3 * NTX + 2 * E::pair(0, s) + 1, v.y);
#pragma unroll
                for (int i = 1; i < NTX; ++i)
                    if (i < s) {
                        const cplx xi = E::cval((code >> (E::BITS * (NTX - 1 - i))) & (M - 1));
                        const cplx cx = cmulc(xs, xi);  // conj(x_i) x_s
                        add(1 + 3 * NTX + 2 * E::pair(i, s), E0 * cx.x);
                        add(1 + 3 * NTX + 2 * E::pair(i, s) + 1, E0 * cx.y);
                    }
            }
#pragma unroll
            for (int r = 0; r < NR; ++r)
                if (32 * r + lane < E::NACC) A[32 * r + lane] += mine[r];
            __syncwarp();
        }
        if (lane == 0) {
            ws[E::O_AREF] = aref;
            *cntp = 0;
        }
        __syncwarp();
        return;
    }
    // General path: one queue entry per lane per round; the warp reduces every statistic right away (rare
    // path: keep the register footprint small so that it does not limit the occupancy of the scan loop)
    // per round every lane drops the NACC contributions of its queue entry into a shared scratch
    // [NACC][32]; afterwards lane i sums statistic i over the lanes that held an entry (a transpose
    // instead of NACC butterfly reductions: ~25 stores + nround loads per lane, no shuffles)
    double* scr = ws + E::O_SCR;
    int nround = 0;
    auto radd = [&](int idx, double v) { scr[idx * 32 + lane] = v; };
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        const bool have = e < n;
        nround = min(32, n - e0);
        const int code = have ? q[e] : 0;
        cplx t0;
        double base = ws[E::O_C0];
        {
            cplx acc[NTX];
#pragma unroll
            for (int i = 0; i < NTX; ++i) acc[i] = mk(ws[E::O_YT + 2 * i], ws[E::O_YT + 2 * i + 1]);
#pragma unroll
            for (int s = NTX - 1; s >= 1; --s) {
                const int m = (code >> (E::BITS * (NTX - 1 - s))) & (M - 1);
                const double eI = acc[s].x - g[s * SQM + (m & (SQM - 1))];
                const double eQ = acc[s].y - g[s * SQM + (m >> E::HB)];
                base += fma(eI, eI, eQ * eQ);
#pragma unroll
                for (int i = 0; i < NTX; ++i)
                    if (i < s) acc[i] = csub(acc[i], tab[E::pair(i, s) * M + m]);
            }
            t0 = acc[0];
        }
        double minI = 1e300, minQ = 1e300;
#pragma unroll
        for (int a = 0; a < SQM; ++a) {
            const double eI = t0.x - g[a], eQ = t0.y - g[a];
            minI = fmin(minI, eI * eI);
            minQ = fmin(minQ, eQ * eQ);
        }
        const double W = have ? exp((aref - (base + minI + minQ)) * inv_s2) : 0.0;
        double EI = 0, A1 = 0, A2 = 0, EQ = 0, B1 = 0, B2 = 0;
#pragma unroll 1
        for (int a = 0; a < SQM; ++a) {
            const double pq = E::pam(a);
            const double eI = t0.x - g[a], eQ = t0.y - g[a];
            const double wI = exp((minI - eI * eI) * inv_s2), wQ = exp((minQ - eQ * eQ) * inv_s2);
            EI += wI; A1 = fma(pq, wI, A1); A2 = fma(pq * pq, wI, A2);
            EQ += wQ; B1 = fma(pq, wQ, B1); B2 = fma(pq * pq, wQ, B2);
        }
        const double E0 = W * EI * EQ;
        const cplx F0 = mk(W * A1 * EQ, -W * EI * B1);  // sum over leaves of e * conj(x_0)
        const double Q0 = W * (A2 * EQ + EI * B2);      // sum over leaves of e * |x_0|^2
        radd(0, E0);
        radd(1, F0.x);
        radd(1 + NTX, F0.y);
        radd(1 + 2 * NTX, Q0);
#pragma unroll
        for (int s = 1; s < NTX; ++s) {
            const cplx xs = E::cval((code >> (E::BITS * (NTX - 1 - s))) & (M - 1));
            radd(1 + s, E0 * xs.x);
            radd(1 + NTX + s, -E0 * xs.y);
            radd(1 + 2 * NTX + s, E0 * cnorm2(xs));
            const cplx v = cmul(F0, xs);  // conj(x_0) x_s summed over the leaves
            radd(1 + 3 * NTX + 2 * E::pair(0, s), v.x);
            radd(1 + 3 * NTX + 2 * E::pair(0, s) + 1, v.y);
#pragma unroll
            for (int i = 1; i < NTX; ++i)
                if (i < s) {
                    const cplx xi = E::cval((code >> (E::BITS * (NTX - 1 - i))) & (M - 1));
                    const cplx cx = cmulc(xs, xi);  // conj(x_i) x_s
                    radd(1 + 3 * NTX + 2 * E::pair(i, s), E0 * cx.x);
                    radd(1 + 3 * NTX + 2 * E::pair(i, s) + 1, E0 * cx.y);
                }
        }
        __syncwarp();
        for (int i = lane; i < E::NACC; i += 32) {
            double acc = 0.0;
            for (int e2 = 0; e2 < nround; ++e2) acc += scr[i * 32 + e2];
            A[i] += acc;
        }
        __syncwarp();
    }
    __syncwarp();
    if (lane == 0) {
        ws[E::O_AREF] = aref;
        *cntp = 0;
    }
    __syncwarp();
}

template <int NTX, int SQM, bool HARD>
struct Scan {
    typedef EnumT<NTX, SQM> E;
    static constexpr int M = E::M;
    double* ws;
    const cplx* tab;
    const double* g;
    double r0;        // R_00 (real): PAM unit of the leaf stream
    double thr;       // 64 varn^2
    double best, lim; // incumbent distance and enqueue limit best + thr
    int bestk;
    bool prune;       // skip subtrees whose partial distance already exceeds lim (bit-identical results)

    // the M leaves below a node: t0 = residual on row 0, nb = partial distance of the node, code = node digits
    __device__ __forceinline__ void leaf(cplx t0, double nb, int code) {
        const double fI = E::fold(t0.x, r0), fQ = E::fold(t0.y, r0);
        const double nodemin = fma(fI, fI, fma(fQ, fQ, nb));
        if (nodemin <= lim) {  // rare at operating SNRs
            if (nodemin <= best) {
                double dI, dQ;
                const int iI = E::slice(t0.x, g, dI), iQ = E::slice(t0.y, g, dQ);
                const double val = nb + dI + dQ;
                const int k = code + ((iQ * SQM + iI) << (E::BITS * (NTX - 1)));
                if (val < best || (val == best && k < bestk)) {
                    best = val;
                    bestk = k;
                    lim = HARD ? val : val + thr;
                }
            }
            if (!HARD) {
                int* cntp = (int*)(ws + E::O_CNT);
                const int slot = atomicAdd(cntp, 1);
                if (slot < ENUM_QCAP) ((int*)(ws + E::O_Q))[slot] = code;
            }
        }
    }

    __device__ __forceinline__ void maybe_flush() {
        if (!HARD) {
            __syncwarp();
            const int c = *(volatile int*)(ws + E::O_CNT);
            if (c > ENUM_QCAP - 32 * SQM) {
                double wb = best;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wb = fmin(wb, shfl_xor_d(wb, o));
                enum_flush<NTX, SQM>(ws, wb);
            }
        }
    }

    // Branch and bound.  Every leaf below a node has d2 >= the node's partial distance `base`; once
    // base > lim (incumbent + 64 varn^2) none of them can improve the arg-min or pass the enqueue test,
    // so the subtree is "dead": visiting it would change nothing.  Lanes carry a dead flag instead of
    // returning (the loops contain warp collectives and must stay convergent); a row or subtree is
    // skipped only when a warp vote says it is dead for all 32 lanes.

    // streams S_ ... 1 enumerated with warp-uniform loops
    template <int S_>
    __device__ __forceinline__ void inner(const cplx (&acc)[NTX], double base, int code, cplx r01, bool dead) {
        dead = dead || (prune && base > lim);
        if constexpr (S_ == 0) {
            if (!dead) leaf(acc[0], base, code);
        } else if constexpr (S_ == 1) {
            if (__all_sync(0xffffffffu, dead)) return;
            // innermost node level: |u_1|^2 and R_01 x_1 are separable in the in-phase / quadrature indices
            const double* g1 = g + SQM;
            double dI1[SQM];
#pragma unroll
            for (int a = 0; a < SQM; ++a) {
                const double eI = acc[1].x - g1[a];
                dI1[a] = eI * eI;
            }
#pragma unroll 1
            for (int iQ = 0; iQ < SQM; ++iQ) {
                // the quadrature term is formed here (a table indexed by the loop counter lived in local memory);
                // product and sum are rounded separately, as the table version did
                const double eQ = acc[1].y - g1[iQ];
                const double bq = base + __dmul_rn(eQ, eQ);
                const bool deadq = dead || (prune && bq > lim);
                if (__all_sync(0xffffffffu, deadq)) continue;
                const double pQ = E::pam(iQ);
                // t = acc0 - pQ * (i r01) = acc0 - pQ*(-r01.y + i r01.x)
                const cplx tq = mk(fma(pQ, r01.y, acc[0].x), fma(-pQ, r01.x, acc[0].y));
                const int cq = code + ((iQ * SQM) << (E::BITS * (NTX - 2)));
                if (!deadq) {
#pragma unroll
                    for (int iI = 0; iI < SQM; ++iI) {
                        const double pI = E::pam(iI);
                        const cplx t0 = mk(fma(-pI, r01.x, tq.x), fma(-pI, r01.y, tq.y));
                        leaf(t0, bq + dI1[iI], cq + (iI << (E::BITS * (NTX - 2))));
                    }
                }
                maybe_flush();
            }
        } else {
            if (__all_sync(0xffffffffu, dead)) return;
#pragma unroll 1
            for (int m = 0; m < M; ++m) {
                const double* gs = g + S_ * SQM;
                const double eI = acc[S_].x - gs[m & (SQM - 1)];
                const double eQ = acc[S_].y - gs[m >> E::HB];
                const double nb = base + fma(eI, eI, eQ * eQ);
                cplx nacc[NTX];
#pragma unroll
                for (int i = 0; i < NTX; ++i) nacc[i] = (i < S_) ? csub(acc[i], tab[E::pair(i, S_) * M + m]) : acc[i];
                inner<S_ - 1>(nacc, nb, code + (m << (E::BITS * (NTX - 1 - S_))), r01, dead);
            }
        }
    }

    // prefix streams (lane-varying digits taken from p), then the uniform inner levels
    template <int S_, int LEFT>
    __device__ __forceinline__ void prefix(const cplx (&acc)[NTX], double base, int code, int p, cplx r01, bool dead) {
        if constexpr (LEFT == 0) {
            inner<S_>(acc, base, code, r01, dead);
        } else {
            const int m = (p >> (E::BITS * (LEFT - 1))) & (M - 1);
            const double* gs = g + S_ * SQM;
            const double eI = acc[S_].x - gs[m & (SQM - 1)];
            const double eQ = acc[S_].y - gs[m >> E::HB];
            const double nb = base + fma(eI, eI, eQ * eQ);
            cplx nacc[NTX];
#pragma unroll
            for (int i = 0; i < NTX; ++i) nacc[i] = (i < S_) ? csub(acc[i], tab[E::pair(i, S_) * M + m]) : acc[i];
            prefix<S_ - 1, LEFT - 1>(nacc, nb, code + (m << (E::BITS * (NTX - 1 - S_))), p, r01, dead);
        }
    }
};

// (4 CTAs per SM at 128 registers with a few spilled words beat 3 CTAs at 140 registers without: 1.16 vs 1.22 ms)
template <int NTX, int SQM, bool HARD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (NTX > 4 ? 2 : 4)) k_enum(Dims d, const double* __restrict__ qr,
                                                     const double* __restrict__ varn,
                                                     const int32_t* __restrict__ active, cplx* __restrict__ stat_m,
                                                     cplx* __restrict__ stat_R, int32_t* __restrict__ kstar,
                                                     double* __restrict__ lse_sym) {
    typedef EnumT<NTX, SQM> E;
    constexpr int M = E::M;
    extern __shared__ __align__(16) double enum_smem[];   // [WARPS][E::WS_DOUBLES]
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * WARPS + warp;
    if (t >= d.T_d) return;
    // the three global reads a warp starts with -- trial flag, noise variance, first 32 doubles of the QR record --
    // are issued together (one after the other they were three DRAM round trips before the first useful
    // instruction, 10 % of the kernel's stall samples in profiles/r02m)
    const double* grec = qr + ((size_t)b * d.T_d + t) * d.rec;
    const int32_t act = (active != nullptr) ? active[b] : 1;
    const double vn = varn[b];
    const double rec0 = (lane < d.rec) ? grec[lane] : 0.0;
    if (act == 0) return;

    double* ws = enum_smem + (size_t)warp * E::WS_DOUBLES;
    cplx* tab = (cplx*)(ws + E::O_TAB);
    double* g = ws + E::O_G;
    // the QR record of the symbol: one coalesced load by the warp into the (still idle) flush scratch, then
    // broadcast reads from shared memory (every lane needs all of it; 30+ uniform global loads per lane before)
    const double* rec = ws + E::O_SCR;
    {
        if (lane < d.rec) ws[E::O_SCR + lane] = rec0;
        for (int i = lane + 32; i < d.rec; i += 32) ws[E::O_SCR + i] = grec[i];
        __syncwarp();
    }
    cplx yt[NTX];
    cplx r01 = mk(0.0, 0.0);
    double r00;
    {
        cplx Rm[NTX][NTX];
        int o = 0;
#pragma unroll
        for (int i = 0; i < NTX; ++i)
#pragma unroll
            for (int j = i; j < NTX; ++j) {
                Rm[i][j] = mk(rec[o], rec[o + 1]);
                o += 2;
            }
#pragma unroll
        for (int i = 0; i < NTX; ++i) {
            yt[i] = mk(rec[o], rec[o + 1]);
            o += 2;
        }
        r00 = Rm[0][0].x;
        if (NTX > 1) r01 = Rm[0][NTX > 1 ? 1 : 0];
        // tables: R_is * c_m for i<s, and the real PAM levels R_ss * pam_a
#pragma unroll
        for (int s = 1; s < NTX; ++s)
#pragma unroll
            for (int i = 0; i < s; ++i)
                for (int m = lane; m < M; m += 32) tab[E::pair(i, s) * M + m] = cmul(Rm[i][s], E::cval(m));
        if (lane < NTX * SQM) {
            const int s = lane / SQM, a = lane % SQM;
            double dd = 0.0;
#pragma unroll
            for (int q = 0; q < NTX; ++q)
                if (q == s) dd = Rm[q][q].x;
            g[lane] = dd * E::pam(a);
        }
    }
    const double c0 = rec[NTX * (NTX + 1) + 2 * NTX];
    const double s2 = vn * vn, inv_s2 = 1.0 / s2;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NTX; ++i) {
            ws[E::O_YT + 2 * i] = yt[i].x;
            ws[E::O_YT + 2 * i + 1] = yt[i].y;
        }
        ws[E::O_C0] = c0;
        ws[E::O_AREF] = 1e300;
        ws[E::O_S2] = inv_s2;
        *(int*)(ws + E::O_CNT) = 0;
    }
    for (int i = lane; i < E::NACC; i += 32) ws[E::O_ACC + i] = 0.0;
    __syncwarp();

    Scan<NTX, SQM, HARD> sc;
    sc.ws = ws;
    sc.tab = tab;
    sc.g = g;
    sc.r0 = r00;
    sc.thr = SBCE_THR * s2;
    sc.prune = (d.flags & SBCE_FLAG_FULL_SCAN) == 0;
    // Babai point (successive slicing): a tight upper bound on min d2 -> incumbent
    {
        cplx acc[NTX];
#pragma unroll
        for (int i = 0; i < NTX; ++i) acc[i] = yt[i];
        double base = c0;
        int k = 0;
#pragma unroll
        for (int s = NTX - 1; s >= 0; --s) {
            double dI, dQ;
            const int iI = E::slice(acc[s].x, g + s * SQM, dI), iQ = E::slice(acc[s].y, g + s * SQM, dQ);
            const int m = iQ * SQM + iI;
            base += dI + dQ;
            k += m << (E::BITS * (NTX - 1 - s));
#pragma unroll
            for (int i = 0; i < NTX; ++i)
                if (i < s) acc[i] = csub(acc[i], tab[E::pair(i, s) * M + m]);
        }
        sc.best = base;
        sc.bestk = k;
        sc.lim = HARD ? base : base + sc.thr;
    }

    // Top-level pre-pass (prefixes of two or more streams): lane m tests the FIRST prefix stream's symbol m against
    // the Babai bound once.  A round of 32 prefixes whose top symbols are all dead would only compute its prefix
    // levels to find every lane dead; it is skipped outright.  Dead under the initial bound implies dead under any
    // later (tighter) one, and a dead round contributes nothing in either mode, so the sequence of visited live
    // nodes -- hence the queue order and every output bit -- is unchanged.
    unsigned top_alive = 0xffffffffu;
    constexpr int TOP_SHIFT = E::BITS * (E::PL - 1);
    if (E::PL >= 2 && sc.prune) {
        bool a = false;
        if (lane < M) {
            const double* gs = g + (NTX - 1) * SQM;
            const double eI = yt[NTX - 1].x - gs[lane & (SQM - 1)];
            const double eQ = yt[NTX - 1].y - gs[lane >> E::HB];
            a = !(c0 + fma(eI, eI, eQ * eQ) > sc.lim);
        }
        top_alive = __ballot_sync(0xffffffffu, a);
    }
    // all lanes iterate the same number of prefixes (invalid ones carry an infinite partial distance)
    for (int pb = 0; pb < E::NPREF; pb += 32) {
        if (E::PL >= 2) {
            const int lo = pb >> TOP_SHIFT, hi = min(pb + 31, E::NPREF - 1) >> TOP_SHIFT;
            const unsigned bits = (hi >= 31 ? 0xffffffffu : ((2u << hi) - 1u)) & ~((1u << lo) - 1u);
            if ((top_alive & bits) == 0u) continue;
        }
        const int p = pb + lane;
        const bool valid = p < E::NPREF;
        sc.template prefix<NTX - 1, E::PL>(yt, c0, 0, valid ? p : 0, r01, !valid);
    }

    // ---- warp merge of the incumbent -------------------------------------------
    double best = sc.best;
    int bestk = sc.bestk;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = shfl_xor_d(best, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
        if (ob < best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
    }
    const size_t sidx = (size_t)b * d.T_d + t;
    if (kstar != nullptr && lane == 0) kstar[sidx] = bestk;

    bool collapsed = HARD;
    double S = 0.0;
    if (!HARD) {
        enum_flush<NTX, SQM>(ws, best);
        S = ws[E::O_ACC];
        // Nothing within 64 varn^2 of the incumbent was queued: only happens when |d2| is so large that
        // its rounding error exceeds the window (e.g. a garbage start vector of norm 1e13); every weight
        // but the best one underflows, so the posterior is the one-hot at k*.
        if (!(S > 0.0) || !isfinite(S)) collapsed = true;
    }
    if (collapsed) {
        if (lane == 0) {
            cplx xs[NTX];
#pragma unroll
            for (int s = 0; s < NTX; ++s) xs[s] = E::cval((bestk >> (E::BITS * (NTX - 1 - s))) & (M - 1));
#pragma unroll
            for (int i = 0; i < NTX; ++i) {
                stat_m[sidx * NTX + i] = cconj(xs[i]);
#pragma unroll
                for (int j = 0; j < NTX; ++j) stat_R[(sidx * NTX + i) * NTX + j] = cmulc(xs[j], xs[i]);
            }
            if (lse_sym != nullptr) lse_sym[sidx] = -best * inv_s2;
        }
        return;
    }
    // outputs: lane l < n_tx^2 writes R[i][j] with (i, j) = (l / n_tx, l % n_tx), lane j < n_tx writes m[j] -- two
    // coalesced stores (lane 0 alone issued all n_tx + n_tx^2 of them one after the other before)
    {
        const double invS = 1.0 / S;
        const double* A = ws + E::O_ACC;
        for (int l = lane; l < NTX * NTX; l += 32) {
            const int i = l / NTX, j = l % NTX;
            cplx v;
            if (i == j) {
                v = mk(A[1 + 2 * NTX + j] * invS, 0.0);
            } else {
                const int lo = min(i, j), hi = max(i, j);
                const int pr = 1 + 3 * NTX + 2 * E::pair(lo, hi);
                const double a = A[pr] * invS, c = A[pr + 1] * invS;
                v = mk(a, i < j ? c : -c);
            }
            stat_R[sidx * NTX * NTX + l] = v;
        }
        if (lane < NTX) stat_m[sidx * NTX + lane] = mk(A[1 + lane] * invS, A[1 + NTX + lane] * invS);
        if (lse_sym != nullptr && lane == 0) lse_sym[sidx] = -ws[E::O_AREF] * inv_s2 + log(S);
    }
}

// ---------------------------------------------------------------------------
// superimposed pilots: statistics of x~ = x + o from those of x
//   m~_i = m_i + conj(o_i),   R~_ij = R_ij + m_i o_j + conj(o_i) conj(m_j) + conj(o_i) o_j
// ---------------------------------------------------------------------------
__global__ void k_superimpose_stats(Dims d, int nb, const cplx* __restrict__ Xoff, const int32_t* __restrict__ active,
                                    cplx* __restrict__ stat_m, cplx* __restrict__ stat_R) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;   // flat (b, t)
    if (s >= nb * d.T_d) return;
    if (active != nullptr && active[s / d.T_d] == 0) return;
    const int n = d.n_tx;
    const cplx* o = Xoff + (size_t)s * n;
    cplx* m = stat_m + (size_t)s * n;
    cplx* R = stat_R + (size_t)s * n * n;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            cplx v = R[i * n + j];
            cfma(v, m[i], o[j]);
            v = cadd(v, cconj(cmul(o[i], m[j])));
            v = cadd(v, cmulc(o[j], o[i]));   // conj(o_i) o_j
            R[i * n + j] = v;
        }
    for (int i = 0; i < n; ++i) m[i] = cadd(m[i], cconj(o[i]));
}

cudaError_t launch_superimpose_stats(const Dims& d, int nb, const double* Xoff, const int32_t* active, double* stat_m,
                                     double* stat_R, cudaStream_t s) {
    const int total = nb * d.T_d;
    k_superimpose_stats<<<(total + 127) / 128, 128, 0, s>>>(d, nb, (const cplx*)Xoff, active, (cplx*)stat_m, (cplx*)stat_R);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------
template <int NTX, int NRX>
static cudaError_t run_heff(const Dims& d, int nb, const double* Yd, const double* PsiD, const double* theta,
                            const int32_t* active, const double* Xoff, double* qr, cudaStream_t s) {
    static SmemOptIn optin_mma, optin;
    dim3 grid((d.T_d + 127) / 128, nb);
    // tensor path once the contraction is a real GEMM: at least one full 8-column tile and two 8-deep steps
    constexpr int NC = NTX * NRX, NT = (NC + 7) / 8;
    const int KP = (d.N1 + 7) & ~7;
    const size_t smem_mma = sizeof(cplx) * ((size_t)KP * (NT * 8 + 2) + 4 * 32 * (NC + 1));
    if (NC >= 8 && d.N1 >= 16 && smem_mma <= 100 * 1024) {
        cudaError_t e = opt_in_smem(optin_mma, (const void*)k_heff_qr_mma<NTX, NRX>, smem_mma);
        if (e != cudaSuccess) return e;
        k_heff_qr_mma<NTX, NRX><<<grid, 128, smem_mma, s>>>(d, (const cplx*)Yd, (const cplx*)PsiD, (const cplx*)theta,
                                                            active, (const cplx*)Xoff, qr);
        count_launch();
        return cudaGetLastError();
    }
    size_t smem = sizeof(cplx) * (size_t)d.L * NRX;
    const int use_smem = smem <= 64 * 1024;
    if (!use_smem) smem = 0;
    cudaError_t e = opt_in_smem(optin, (const void*)k_heff_qr<NTX, NRX>, smem);
    if (e != cudaSuccess) return e;
    k_heff_qr<NTX, NRX><<<grid, 128, smem, s>>>(d, (const cplx*)Yd, (const cplx*)PsiD, (const cplx*)theta, active,
                                                (const cplx*)Xoff, qr, use_smem);
    count_launch();
    return cudaGetLastError();
}

template <int NTX>
static cudaError_t run_heff_ntx(const Dims& d, int nb, const double* Yd, const double* PsiD, const double* theta,
                                const int32_t* active, const double* Xoff, double* qr, cudaStream_t s) {
    switch (d.n_rx) {
        case 1: return run_heff<NTX, 1>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 2: return run_heff<NTX, 2>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 3: return run_heff<NTX, 3>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 4: return run_heff<NTX, 4>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 6: return run_heff<NTX, 6>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 8: return run_heff<NTX, 8>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        default: return cudaErrorInvalidValue;
    }
}

template <int NTX, int SQM>
static cudaError_t run_enum(const Dims& d, int nb, const double* qr, const double* varn, const int32_t* active,
                            double* stat_m, double* stat_R, int32_t* kstar, double* lse_sym, cudaStream_t s) {
    constexpr int WARPS = 4;
    dim3 grid((d.T_d + WARPS - 1) / WARPS, nb);
    constexpr size_t smem = sizeof(double) * WARPS * EnumT<NTX, SQM>::WS_DOUBLES;
    static SmemOptIn optin_hard, optin_soft;
    cudaError_t e;
    if (d.mode == SBCE_MODE_HARD) {
        e = opt_in_smem(optin_hard, (const void*)k_enum<NTX, SQM, true, WARPS>, smem);
        if (e != cudaSuccess) return e;
        k_enum<NTX, SQM, true, WARPS><<<grid, WARPS * 32, smem, s>>>(d, qr, varn, active, (cplx*)stat_m,
                                                                      (cplx*)stat_R, kstar, lse_sym);
    } else {
        e = opt_in_smem(optin_soft, (const void*)k_enum<NTX, SQM, false, WARPS>, smem);
        if (e != cudaSuccess) return e;
        k_enum<NTX, SQM, false, WARPS><<<grid, WARPS * 32, smem, s>>>(d, qr, varn, active, (cplx*)stat_m,
                                                                       (cplx*)stat_R, kstar, lse_sym);
    }
    count_launch();
    return cudaGetLastError();
}

template <int NTX>
static cudaError_t run_heff_rows(const Dims& d, int nb, const double* Yd, const double* PsiD, const double* theta,
                                 const int32_t* active, const double* Xoff, double* qr, cudaStream_t s) {
    dim3 grid((d.T_d + 15) / 16, nb);   // 8 lanes per symbol, 16 symbols per CTA
    k_heff_qr_rows<NTX><<<grid, 128, 0, s>>>(d, (const cplx*)Yd, (const cplx*)PsiD, (const cplx*)theta, active,
                                             (const cplx*)Xoff, qr);
    count_launch();
    return cudaGetLastError();
}

// wide trees: only constellations whose joint hypothesis index fits comfortably in 32 bits are instantiated
// (n_tx log2 M <= 24); abi.cu::make_dims rejects the rest before any launch
template <int NTX>
static cudaError_t run_enum_wide(const Dims& d, int nb, const double* qr, const double* varn, const int32_t* active,
                                 double* stat_m, double* stat_R, int32_t* kstar, double* lse_sym, cudaStream_t s) {
    if (d.sqM == 2) return run_enum<NTX, 2>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
    if constexpr (NTX <= 6) {
        if (d.sqM == 4) return run_enum<NTX, 4>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
    }
    return cudaErrorInvalidValue;
}

template <int NTX>
static cudaError_t run_enum_ntx(const Dims& d, int nb, const double* qr, const double* varn, const int32_t* active,
                                double* stat_m, double* stat_R, int32_t* kstar, double* lse_sym, cudaStream_t s) {
    switch (d.sqM) {
        case 2: return run_enum<NTX, 2>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        case 4: return run_enum<NTX, 4>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        case 8: return run_enum<NTX, 8>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_heff_qr(const Dims& d, int nb, const double* Yd, const double* PsiD, const double* theta,
                           const int32_t* active, const double* Xoff, double* qr, cudaStream_t s) {
    switch (d.n_tx) {
        case 1: return run_heff_ntx<1>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 2: return run_heff_ntx<2>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 3: return run_heff_ntx<3>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 4: return run_heff_ntx<4>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 5: return run_heff_rows<5>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 6: return run_heff_rows<6>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 7: return run_heff_rows<7>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        case 8: return run_heff_rows<8>(d, nb, Yd, PsiD, theta, active, Xoff, qr, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_enum(const Dims& d, int nb, const double* qr, const double* varn, const int32_t* active,
                        double* stat_m, double* stat_R, int32_t* kstar, double* lse_sym, cudaStream_t s) {
    switch (d.n_tx) {
        case 1: return run_enum_ntx<1>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        case 2: return run_enum_ntx<2>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        case 3: return run_enum_ntx<3>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        case 4: return run_enum_ntx<4>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        case 5: return run_enum_wide<5>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        case 6: return run_enum_wide<6>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        case 7: return run_enum_wide<7>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        case 8: return run_enum_wide<8>(d, nb, qr, varn, active, stat_m, stat_R, kstar, lse_sym, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sbce
