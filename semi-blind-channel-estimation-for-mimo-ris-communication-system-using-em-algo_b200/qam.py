"""Square QAM constellation and joint-hypothesis indexing used by the estimator.

Restates what the reference takes from its vendored modulation module
(/root/reference/Proposed method/QAM.py:310-322): the UN-normalised grid
c[iQ*sqrt(M)+iI] = (2 iI - sqrt(M) + 1) + 1j (2 iQ - sqrt(M) + 1), and the
itertools.product ordering of joint hypotheses (Proposed method/PM.py:25-31):
k = sum_j idx_j * M**(n_tx-1-j), stream 0 most significant.
"""
from __future__ import annotations

import numpy as np

SUPPORTED_M = (4, 16, 64)


def constellation(M: int) -> np.ndarray:
    root = int(round(float(M) ** 0.5))
    if root * root != M or root & (root - 1):
        raise ValueError("constellation order must be a square power of two, got %r" % (M,))
    levels = 2.0 * np.arange(root) - (root - 1)
    grid = levels.reshape(1, root) + 1j * levels.reshape(root, 1)
    return np.ascontiguousarray(grid.reshape(M), dtype=np.complex128)


def digits_of(k, M: int, n_tx: int) -> np.ndarray:
    """(..., n_tx) per-stream constellation indices of joint hypothesis indices k."""
    k = np.asarray(k, dtype=np.int64)
    shifts = M ** np.arange(n_tx - 1, -1, -1, dtype=np.int64)
    return (k[..., None] // shifts) % M


def symbols_of(k, M: int, n_tx: int) -> np.ndarray:
    """(..., n_tx) complex symbol vectors of joint hypothesis indices k."""
    return constellation(M)[digits_of(k, M, n_tx)]


def constellation_from_table(all_possibleSymbols, M: int) -> np.ndarray:
    """Recover the per-stream constellation from the reference's (K, n_tx) table."""
    tab = np.asarray(all_possibleSymbols)
    n_tx = tab.shape[1]
    return np.ascontiguousarray(tab[:: M ** (n_tx - 1), 0], dtype=np.complex128)
