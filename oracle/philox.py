"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the on-device input generator
(csrc/gen.cu), used by tests/ to check the CUDA generator element by element.

The generator is NOT part of the reference (which draws from numpy's legacy
Mersenne-Twister global state, `Proposed method/PM.py:11-40,119-148`); it
replaces the reference's per-trial host generation for on-device Monte-Carlo
sweeps and follows the same signal model.  What is
pinned here: Philox4x32-10 against the Random123 known-answer vectors, symbol
indices bit-exact, every floating-point array to 1e-12.

Counter layout: Philox(counter = (element, array id, trial lo, trial hi), key = seed lo/hi).
Array ids: 0 H_BU, 1 H_BS, 2 H_SU, 3 data symbols, 4 pilot symbols, 5 data phases,
6 pilot noise, 7 data noise.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs broadcastable unsigned 32-bit values."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & MASK for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _draw(seed, stream, trial, elems):
    elems = np.asarray(elems, dtype=np.uint64)
    return philox4x32_10(elems, stream, trial & 0xFFFFFFFF, (trial >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF,
                         (seed >> 32) & 0xFFFFFFFF)


def _u53(a, b):
    v = (a << np.uint64(21)) ^ (b >> np.uint64(11))
    return (v.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _cn(seed, stream, trial, elems, var):
    x, y, z, w = _draw(seed, stream, trial, elems)
    u1, u2 = _u53(x, y), _u53(z, w)
    rad = np.sqrt(-var * np.log(u1))
    return rad * np.cos(2 * np.pi * u2) + 1j * rad * np.sin(2 * np.pi * u2)


def _dft(T, ns, denom):
    t = np.arange(T, dtype=np.int64)[:, None]
    q = (t * np.asarray(ns, dtype=np.int64)[None, :]) % denom
    return np.exp(-2j * np.pi * q / denom)


def generate_trial(N, n_tx, n_rx, M, T_p, T_d, varn, seed, trial, pilot_design="pm", data_phases="random", varh=1.0,
                   direct_link=True):
    """One trial exactly as csrc/gen.cu generates it.  Returns a dict with the estimator layout of
    oracle/em_numpy.py plus the symbol indices and the noise blocks.  N + 1 phase rows; with
    direct_link=False all of them are RIS elements ("direct vs non direct - T_pv s nmse.py")."""
    s = int(round(M ** 0.5))
    hb = (M.bit_length() - 1) // 2
    if not direct_link:
        return _generate_trial_nodirect(N + 1, n_tx, n_rx, M, T_p, T_d, varn, seed, trial, pilot_design, data_phases, varh)
    H_BU = _cn(seed, 0, trial, np.arange(n_rx * n_tx), varh).reshape(n_rx, n_tx)
    H_BS = _cn(seed, 1, trial, np.arange(N * n_tx), varh).reshape(N, n_tx)
    H_SU = _cn(seed, 2, trial, np.arange(n_rx * N), varh).reshape(n_rx, N)
    Th = np.empty((N + 1, n_tx, n_rx), np.complex128)
    Th[0] = H_BU.T
    Th[1:] = H_BS[:, :, None] * H_SU.T[:, None, :]
    h = Th.reshape((N + 1) * n_tx, n_rx)

    def syms(stream, T):
        idx = (_draw(seed, stream, trial, np.arange(T * n_tx))[0] & np.uint64(M - 1)).astype(np.int64).reshape(T, n_tx)
        return idx, (2 * (idx & (s - 1)) - s + 1) + 1j * (2 * (idx >> hb) - s + 1)

    idx_d, Xd = syms(3, T_d)
    idx_p, Xp = syms(4, T_p)
    if pilot_design == "pm":
        PsiP = np.zeros((T_p, N + 1), np.complex128)
        PsiP[:, :N] = _dft(T_p, np.arange(N), N)
    else:
        PsiP = np.ones((T_p, N + 1), np.complex128)
        PsiP[:, 1:] = _dft(T_p, np.arange(N), T_p)
    if data_phases == "dft":
        PsiD = _dft(T_d, np.arange(N + 1), T_d)
    else:
        x, y, _, _ = _draw(seed, 5, trial, np.arange(T_d * N))
        PsiD = np.ones((T_d, N + 1), np.complex128)
        PsiD[:, 1:] = np.exp(2j * np.pi * _u53(x, y)).reshape(T_d, N)
    noise_p = _cn(seed, 6, trial, np.arange(T_p * n_rx), varn).reshape(T_p, n_rx)
    noise_d = _cn(seed, 7, trial, np.arange(T_d * n_rx), varn).reshape(T_d, n_rx)
    Wp = (PsiP[:, :, None] * Xp[:, None, :]).reshape(T_p, -1)
    Wd = (PsiD[:, :, None] * Xd[:, None, :]).reshape(T_d, -1)
    return dict(h=h, Xp=Xp.astype(np.complex128), Xd=Xd.astype(np.complex128), idx_p=idx_p, idx_d=idx_d, PsiP=PsiP,
                PsiD=PsiD, Yp=Wp @ h + noise_p, Yd=Wd @ h + noise_d, noise_p=noise_p, noise_d=noise_d, Wp=Wp)


def _generate_trial_nodirect(R, n_tx, n_rx, M, T_p, T_d, varn, seed, trial, pilot_design, data_phases, varh):
    """R RIS elements, no BS-user link: Theta[n][j][r] = H_BS[n, j] H_SU[r, n]; every phase row is an element."""
    s = int(round(M ** 0.5))
    hb = (M.bit_length() - 1) // 2
    H_BS = _cn(seed, 1, trial, np.arange(R * n_tx), varh).reshape(R, n_tx)
    H_SU = _cn(seed, 2, trial, np.arange(n_rx * R), varh).reshape(n_rx, R)
    h = (H_BS[:, :, None] * H_SU.T[:, None, :]).reshape(R * n_tx, n_rx)

    def syms(stream, T):
        idx = (_draw(seed, stream, trial, np.arange(T * n_tx))[0] & np.uint64(M - 1)).astype(np.int64).reshape(T, n_tx)
        return idx, (2 * (idx & (s - 1)) - s + 1) + 1j * (2 * (idx >> hb) - s + 1)

    idx_d, Xd = syms(3, T_d)
    idx_p, Xp = syms(4, T_p)
    PsiP = _dft(T_p, np.arange(R), R if pilot_design == "pm" else T_p)
    if data_phases == "dft":
        PsiD = _dft(T_d, np.arange(R), T_d)
    else:
        x, y, _, _ = _draw(seed, 5, trial, np.arange(T_d * R))
        PsiD = np.exp(2j * np.pi * _u53(x, y)).reshape(T_d, R)
    noise_p = _cn(seed, 6, trial, np.arange(T_p * n_rx), varn).reshape(T_p, n_rx)
    noise_d = _cn(seed, 7, trial, np.arange(T_d * n_rx), varn).reshape(T_d, n_rx)
    Wp = (PsiP[:, :, None] * Xp[:, None, :]).reshape(T_p, -1)
    Wd = (PsiD[:, :, None] * Xd[:, None, :]).reshape(T_d, -1)
    return dict(h=h, Xp=Xp.astype(np.complex128), Xd=Xd.astype(np.complex128), idx_p=idx_p, idx_d=idx_d, PsiP=PsiP,
                PsiD=PsiD, Yp=Wp @ h + noise_p, Yd=Wd @ h + noise_d, noise_p=noise_p, noise_d=noise_d, Wp=Wp)
