"""Host-side signal model of the MIMO-RIS link: channels, symbols, RIS phase
designs, received blocks and the least-squares start.  These are the functions
every reference driver calls before the estimator
(/root/reference/Proposed method/PM.py:11-40,119-148); here they are batched
and produce the dense structure-of-arrays layout the CUDA library consumes
(the Kronecker design matrices Z_t are never built).

`legacy=True` reproduces the reference's global-RNG draw ORDER trial by trial
 so that a seeded trial is byte-identical to what the
reference scripts would have generated; `legacy=False` draws a whole batch at
once from numpy's Generator (throughput runs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .qam import constellation


@dataclass
class TrialBatch:
    """SoA container, trial index outermost, complex128 C-order."""
    h: np.ndarray        # (B, L, n_rx)  true channel, Theta layout
    Xd: np.ndarray       # (B, T_d, n_tx)
    Xp: np.ndarray       # (B, T_p, n_tx)
    idx_d: np.ndarray    # (B, T_d, n_tx) constellation indices of the data
    PsiP: np.ndarray     # (B, T_p, N+1)
    PsiD: np.ndarray     # (B, T_d, N+1)
    Yp: np.ndarray       # (B, T_p, n_rx)
    Yd: np.ndarray       # (B, T_d, n_rx)
    theta0: np.ndarray   # (B, L, n_rx)  LS start
    varn: np.ndarray     # (B,)


def _cn(rs, shape, var):
    """CN(0,var) with the reference's memory trick: normal(size=(..., 2k)).view(complex)."""
    raw = rs.normal(0.0, math.sqrt(var / 2.0), shape[:-1] + (2 * shape[-1],))
    return raw.view(np.complex128)


def cascaded_channel(H_BU, H_BS, H_SU, flatten="F"):
    """Theta (L, n_rx) from the three links (Proposed method/PM.py:15): row
    n'*n_tx + j is the direct link H_BU[:, j] for n'=0 and H_BS[n, j] * H_SU[:, n]
    for n' = n + 1 (Khatri-Rao of H_BS^T and H_SU, never materialised as a matrix
    product).  flatten="C" is the scrambled arrangement of the two top-level
    scripts (Proposed_method_NMSEvsTp.py:14)."""
    n_rx, n_tx = H_BU.shape
    N = H_BS.shape[0]
    if flatten == "C":
        kr = H_BS.T[:, None, :] * H_SU[None, :, :]
        return np.concatenate((H_BU.reshape(-1), kr.reshape(-1))).reshape((N + 1) * n_tx, n_rx)
    Th = np.empty((N + 1, n_tx, n_rx), dtype=np.complex128)
    Th[0] = H_BU.T
    Th[1:] = H_BS[:, :, None] * H_SU.T[:, None, :]
    return Th.reshape((N + 1) * n_tx, n_rx)


def _dft_phases(T, rows, denom):
    # scalar Python complex arithmetic, term by term as the reference forms it
    # (PM.py:124); numpy's vector complex division rounds differently.
    arg = np.empty((T, rows), dtype=np.complex128)
    for t in range(T):
        for n in range(rows):
            arg[t, n] = (-1j * 2 * np.pi * (t) * (n)) / (denom)
    return np.exp(arg)


def pilot_phases(T_p, N, variant="pm"):
    """(T_p, N+1).  "pm": exp(-j2pi t n/N) for n<N in columns 0..N-1, column N zero
    (last element off during pilots, PM.py:120-124).  "top": ones column + exp(-j2pi t n/T_p)
    (Proposed_method_NMSEvsTp.py:77,129)."""
    if variant == "pm":
        P = np.zeros((T_p, N + 1), dtype=np.complex128)
        P[:, :N] = _dft_phases(T_p, N, N)
    else:
        P = np.ones((T_p, N + 1), dtype=np.complex128)
        P[:, 1:] = _dft_phases(T_p, N, T_p)
    return P


def data_phases_random(T_d, N, rs, beta_min=0.0, beta_max=2 * math.pi, amp=1.0, legacy=True):
    """(T_d, N+1): ones column (direct link, PM.py:179) + exp(j U(beta_min,beta_max)) (PM.py:125-129)."""
    D = np.ones((T_d, N + 1), dtype=np.complex128)
    if legacy:
        for t in range(T_d):
            u = rs.uniform(0, 1, (N, 1))
            D[t, 1:] = (amp * np.exp(1j * ((beta_max - beta_min) * u + beta_min)))[:, 0]
    else:
        u = rs.uniform(0, 1, (T_d, N))
        D[:, 1:] = amp * np.exp(1j * ((beta_max - beta_min) * u + beta_min))
    return D


def data_phases_dft(T_d, N):
    """Deterministic variant of Proposed_method_NMSEvsTd.py:92-94 (all N+1 rows DFT over T_d)."""
    return _dft_phases(T_d, N + 1, T_d)


def draw_indices(rs, M, n_tx, T, legacy=True):
    if legacy:
        out = np.empty((T, n_tx), dtype=np.int64)
        for t in range(T):
            out[t] = rs.choice(range(0, M), n_tx, True)
        return out
    return rs.integers(0, M, (T, n_tx))


def design_rows(Psi, X):
    """W[t] = psi~_t (x) x_t, (T, L)."""
    return (Psi[..., :, None] * X[..., None, :]).reshape(Psi.shape[:-1] + (-1,))


def ls_start(Wp, Yp):
    """theta0 = pinv(W_p) Y_p (PM.py:147 with Z_p = W_p (x) I_nrx)."""
    return np.linalg.pinv(Wp) @ Yp


def generate_trial(N, n_tx, n_rx, M, T_p, T_d, varn, rs, order="pm", variant="pm", varh=1.0):
    """One realisation in the reference draw order (legacy RandomState `rs`)."""
    cons = constellation(M)
    sd_shape = lambda a, b: (a, b)
    H_BU = _cn(rs, sd_shape(n_rx, n_tx), varh)
    H_BS = _cn(rs, sd_shape(N, n_tx), varh)
    H_SU = _cn(rs, sd_shape(n_rx, N), varh)
    h = cascaded_channel(H_BU, H_BS, H_SU)
    idx_d = draw_indices(rs, M, n_tx, T_d)
    if order == "rev4":
        idx_p = draw_indices(rs, M, n_tx, T_p)
    PsiP = pilot_phases(T_p, N, "pm" if variant == "pm" else "top")
    PsiD = data_phases_dft(T_d, N) if variant == "top_td" else data_phases_random(T_d, N, rs)
    if order != "rev4":
        idx_p = draw_indices(rs, M, n_tx, T_p)
    Xd, Xp = cons[idx_d], cons[idx_p]
    Yp, Yd = noisy_blocks(PsiP, PsiD, Xp, Xd, h, varn, rs)
    theta0 = ls_start(design_rows(PsiP, Xp), Yp)
    return dict(h=h, Xd=Xd, Xp=Xp, idx_d=idx_d, PsiP=PsiP, PsiD=PsiD, Yp=Yp, Yd=Yd, theta0=theta0)


def noisy_blocks(PsiP, PsiD, Xp, Xd, h, varn, rs, legacy=True):
    """Y = W Theta + CN(0, varn) noise, pilots first then data (PM.py:137-146)."""
    n_rx = h.shape[-1]
    Yp = design_rows(PsiP, Xp) @ h
    Yd = design_rows(PsiD, Xd) @ h
    if legacy:
        for blk in (Yp, Yd):
            for t in range(blk.shape[0]):
                blk[t] += _cn(rs, (n_rx, 1), varn)[:, 0]
    else:
        Yp += _cn(rs, Yp.shape, varn)
        Yd += _cn(rs, Yd.shape, varn)
    return Yp, Yd


def generate_batch(N, n_tx, n_rx, M, T_p, T_d, varn, B, seed=0, legacy=True, order="pm", variant="pm",
                   ls="pinv") -> TrialBatch:
    """B independent trials; trial b uses seed + b when legacy (disjoint, reproducible
    per trial regardless of how the batch is sharded).  `ls` (non-legacy only): "pinv" is the reference's
    LS start (PM.py:147); "normal" solves the pilot normal equations instead -- the same vector to rounding
    when W_p has full column rank (T_p >= L), at O(L^3) instead of an SVD per trial (long channels)."""
    L = (N + 1) * n_tx
    varn_arr = np.broadcast_to(np.asarray(varn, dtype=np.float64), (B,)).copy()
    out = TrialBatch(h=np.empty((B, L, n_rx), np.complex128), Xd=np.empty((B, T_d, n_tx), np.complex128),
                     Xp=np.empty((B, T_p, n_tx), np.complex128), idx_d=np.empty((B, T_d, n_tx), np.int64),
                     PsiP=np.empty((B, T_p, N + 1), np.complex128), PsiD=np.empty((B, T_d, N + 1), np.complex128),
                     Yp=np.empty((B, T_p, n_rx), np.complex128), Yd=np.empty((B, T_d, n_rx), np.complex128),
                     theta0=np.empty((B, L, n_rx), np.complex128), varn=varn_arr)
    if legacy:
        for b in range(B):
            t = generate_trial(N, n_tx, n_rx, M, T_p, T_d, float(varn_arr[b]), np.random.RandomState(seed + b),
                               order=order, variant=variant)
            for k in ("h", "Xd", "Xp", "idx_d", "PsiP", "PsiD", "Yp", "Yd", "theta0"):
                getattr(out, k)[b] = t[k]
        return out
    rng = np.random.default_rng(seed)
    cons = constellation(M)

    def cn(shape, var):
        s = math.sqrt(0.5)
        return (rng.standard_normal(shape) * s + 1j * rng.standard_normal(shape) * s) * np.sqrt(var)

    H_BU, H_BS, H_SU = cn((B, n_rx, n_tx), 1.0), cn((B, N, n_tx), 1.0), cn((B, n_rx, N), 1.0)
    Th = np.empty((B, N + 1, n_tx, n_rx), np.complex128)
    Th[:, 0] = H_BU.transpose(0, 2, 1)
    Th[:, 1:] = H_BS[:, :, :, None] * H_SU.transpose(0, 2, 1)[:, :, None, :]
    out.h[:] = Th.reshape(B, L, n_rx)
    out.idx_d[:] = rng.integers(0, M, (B, T_d, n_tx))
    out.Xd[:] = cons[out.idx_d]
    out.Xp[:] = cons[rng.integers(0, M, (B, T_p, n_tx))]
    out.PsiP[:] = pilot_phases(T_p, N, "pm" if variant == "pm" else "top")[None]
    if variant == "top_td":
        out.PsiD[:] = data_phases_dft(T_d, N)[None]
    else:
        out.PsiD[:, :, 0] = 1.0
        out.PsiD[:, :, 1:] = np.exp(1j * rng.uniform(0.0, 2 * math.pi, (B, T_d, N)))
    Wp = design_rows(out.PsiP, out.Xp)
    Wd = design_rows(out.PsiD, out.Xd)
    vn = varn_arr[:, None, None]
    out.Yp[:] = Wp @ out.h + cn((B, T_p, n_rx), 1.0) * np.sqrt(vn)
    out.Yd[:] = Wd @ out.h + cn((B, T_d, n_rx), 1.0) * np.sqrt(vn)
    if ls == "normal":
        if T_p < L:
            raise ValueError("ls='normal' needs T_p >= L (full column rank pilot block)")
        WpH = Wp.conj().transpose(0, 2, 1)
        out.theta0[:] = np.linalg.solve(WpH @ Wp, WpH @ out.Yp)
    else:
        out.theta0[:] = np.linalg.pinv(Wp) @ out.Yp
    return out


# ---------------------------------------------------------------------------
# reference-named wrappers (same argument order and return objects as
# /root/reference/Proposed method/PM.py:11-40,119-148; global numpy RNG)
# ---------------------------------------------------------------------------

def channelMatrix(n_tx, n_rx, N, varh):
    rs = np.random
    H_BU = _cn(rs, (n_rx, n_tx), varh)
    H_BS = _cn(rs, (N, n_tx), varh)
    H_SU = _cn(rs, (n_rx, N), varh)
    return cascaded_channel(H_BU, H_BS, H_SU).reshape(-1)


def symbols(n_tx, M, T_d):
    import itertools

    cons = constellation(M)
    idx = draw_indices(np.random, M, n_tx, T_d)
    X_d = [cons[idx[t]].reshape(n_tx, 1) for t in range(T_d)]
    table = np.asarray(list(itertools.product(*([cons] * n_tx))))
    return X_d, table, cons


def pilotSymbols(n_tx, M, T_p):
    cons = constellation(M)
    idx = draw_indices(np.random, M, n_tx, T_p)
    return [cons[idx[t]].reshape(n_tx, 1) for t in range(T_p)]


def irsMatrix(T_p, T_d, N, beta_min, amp, beta_max=2 * math.pi):
    """Returns (PsiTilde_tp (N+1,T_p), PsiTilde_td (N,T_d)) like PM.py:119-130; the caller
    inserts the direct-link ones row into PsiTilde_td (PM.py:179)."""
    PsiTilde_tp = pilot_phases(T_p, N, "pm").T.copy()
    PsiTilde_td = data_phases_random(T_d, N, np.random, beta_min, beta_max, amp)[:, 1:].T.copy()
    return PsiTilde_tp, PsiTilde_td


def receivedSignals(T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h, varn, M_symbols):
    """Returns (Y_p, Y_d, Z_p, Z_d, h_initial) like PM.py:132-148.  Z_p / Z_d are
    returned as lazily-expandable design rows (objects exposing the dense
    (n_rx, D) matrix on demand) so that nothing O(n_rx*D) per symbol is built
    unless a caller really indexes them."""
    PsiP = np.asarray(PsiTilde_tp).T
    PsiD = np.asarray(PsiTilde_td).T
    Xp = np.hstack(X_p).T
    Xd = np.hstack(X_d).T
    L = PsiD.shape[1] * n_tx
    Theta = np.asarray(h).reshape(L, n_rx)
    Yp, Yd = noisy_blocks(PsiP, PsiD, Xp, Xd, Theta, varn, np.random)
    Wp, Wd = design_rows(PsiP, Xp), design_rows(PsiD, Xd)
    h_initial = ls_start(Wp, Yp).reshape(-1, 1)
    Y_p = [Yp[t].reshape(n_rx, 1) for t in range(T_p)]
    Y_d = [Yd[t].reshape(n_rx, 1) for t in range(T_d)]
    Z_p = [DesignRow(Wp[t], n_rx) for t in range(T_p)]
    Z_d = [DesignRow(Wd[t], n_rx) for t in range(T_d)]
    return Y_p, Y_d, Z_p, Z_d, h_initial


class DesignRow:
    """Implicit Z_t = w_t^T (x) I_nrx: stores only w_t; indexable like the dense
    (n_rx, D) matrix of the reference (Z[r, l*n_rx + r'] = w[l] delta_rr')."""

    def __init__(self, w, n_rx):
        self.w = np.asarray(w, dtype=np.complex128)
        self.n_rx = int(n_rx)
        self.shape = (self.n_rx, self.w.size * self.n_rx)

    def dense(self):
        return np.kron(self.w[None, :], np.eye(self.n_rx, dtype=np.complex128))

    def __array__(self, dtype=None, copy=None):
        d = self.dense()
        return d if dtype is None else d.astype(dtype)

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2 and key[0] == 0 and key[1] == slice(0, None, self.n_rx):
            return self.w
        return self.dense()[key]


# ---------------------------------------------------------------------------
# structured-product helpers the reference keeps as dense matrices
# ---------------------------------------------------------------------------

def khatri_rao(a, b):
    """Column-wise Kronecker product (Proposed_method_NMSEvsTd.py:9-19 / scipy.linalg.khatri_rao):
    c[i*b.shape[0] + k, j] = a[i, j] * b[k, j]."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("khatri_rao needs two 2-D arrays with the same number of columns")
    return (a[:, None, :] * b[None, :, :]).reshape(a.shape[0] * b.shape[0], a.shape[1])


def commutation_permutation(m, n):
    """Index form of the commutation matrix K(m,n) of `Proposed method/commutation_matrix.py:3-8`:
    K = I[w, :] with w = arange(m*n).reshape((m, n), order='F').T.ravel(order='F'), so that
    K @ vec_F(A) = vec_F(A^T) for an (m, n) matrix A.  Only the permutation is returned (the reference's
    module-level demo materialises a 65536^2 identity, 34 GB); apply it with x[w]."""
    return np.arange(m * n).reshape((m, n), order="F").T.ravel(order="F")


def design_index(n_prime, j, r, n_tx, n_rx):
    """Column of the reference's Kronecker design matrix Z = psi~^T (x) x^T (x) I_nrx that multiplies
    Theta[n'*n_tx + j, r]: (n' * n_tx + j) * n_rx + r.  This index map is all the CUDA kernels keep of
    the Kronecker / Khatri-Rao / commutation structure."""
    return (n_prime * n_tx + j) * n_rx + r
