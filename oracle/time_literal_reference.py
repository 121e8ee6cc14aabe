"""TEST INFRASTRUCTURE ONLY -- times the LITERAL reference `em` (unmodified source, loaded through
oracle/ref_harness.py) in the build container and extrapolates it to the north-star workload
(BASELINE.md section 4.1).  It cannot run on the GPU box (no
/root/reference there); bench.py's CPU arm times the numpy port instead and BASELINE.md records the numbers
this script printed here.

    python oracle/time_literal_reference.py            # prints one JSON object

What is measured
  (A) `Proposed_method_NMSEvsTp.py:43-69` em at its shipped size (N=32, 2x2, QPSK, T_p=40, T_d=50,
      10 iterations), complete: one trial, single Python thread.
  (B) the same function at the north-star channel size (N=64, 4x4 -> D = 1040 unknowns, T_p = 320) on a
      SLICE: T_d <= 5 data symbols and QPSK (K = 256 joint hypotheses) instead of T_d = 256 and 16-QAM
      (K = 65536), one iteration.  The function's cost per (symbol, hypothesis) pair does not depend on M
      or T_d (two np.kron chains, one (D x n_rx)(n_rx x D) product and a D x D accumulation per pair,
      :55-62), so the full workload is  itera * [ T_d * K * c_pair + T_p * c_pilot + c_solve ]  with the
      three constants fitted from the slice; the result is labelled "extrapolated".
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

TWO_PI = 2 * np.pi


def _inputs(ns, N, n_tx, n_rx, M, T_p, T_d, varn):
    """Reference call order of Proposed_method_NMSEvsTp.py:122-136."""
    h = ns["channelMatrix"](n_tx, n_rx, N, 1)
    X_d, aps = ns["symbols"](n_tx, M, T_d)
    PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
    PsiTilde_tp = np.insert(PsiTilde_tp, 0, np.ones((1, T_p), dtype="complex128"), axis=0)
    PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
    X_p = ns["pilotSymbols"](n_tx, M, T_p)
    out = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h, varn, M)
    Y_p, Y_d, Z_p = out[0], out[1], out[2]
    return dict(h=h, aps=aps, PsiTilde_td=PsiTilde_td, Y_p=Y_p, Y_d=Y_d, Z_p=Z_p)


def time_em(N, n_tx, n_rx, M, T_p, T_d, itera, varn=0.1, seed=0):
    ns = rh.load_functions("Proposed_method_NMSEvsTp.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        g = _inputs(ns, N, n_tx, n_rx, M, T_p, T_d, varn)
        t0 = time.perf_counter()
        ns["em"](g["Y_d"], g["Y_p"], T_d, T_p, g["Z_p"], g["PsiTilde_td"], g["aps"], M, varn, itera)
        return time.perf_counter() - t0


def main():
    if not rh.reference_available():
        raise SystemExit("needs /root/reference (build container only)")
    out = {"host": "%d vCPU build container, single Python thread" % (os.cpu_count() or 1)}
    # (A) shipped config 1 point, complete
    tA = time_em(32, 2, 2, 4, 40, 50, 10)
    out["config1_point"] = dict(N=32, n_tx=2, n_rx=2, M=4, T_p=40, T_d=50, itera=10, seconds=tA,
                                trials_per_s=1.0 / tA, kind="measured, complete")
    # (B) north-star channel size, slices that separate the three cost terms
    N, n_tx, n_rx = 64, 4, 4
    time_em(N, n_tx, n_rx, 4, 1, 1, 1)                                        # warm-up (imports, BLAS threads)
    t_11 = min(time_em(N, n_tx, n_rx, 4, 1, 1, 1) for _ in range(2))          # T_p = 1,  T_d = 1: K pairs + 1 pilot + solve
    t_15 = time_em(N, n_tx, n_rx, 4, 1, 5, 1)                                 # T_p = 1,  T_d = 5
    t_41 = time_em(N, n_tx, n_rx, 4, 41, 1, 1)                                # T_p = 41, T_d = 1
    K_slice = 4 ** n_tx
    c_pair = (t_15 - t_11) / (4.0 * K_slice)      # seconds per (symbol, hypothesis) pair, both loops
    c_pilot = max(0.0, (t_41 - t_11) / 40.0)      # seconds per pilot symbol
    c_solve = max(0.0, t_11 - K_slice * c_pair - c_pilot)
    T_p, T_d, K, itera = 320, 256, 16 ** n_tx, 10
    per_iter = T_d * K * c_pair + T_p * c_pilot + c_solve
    out["north_star_extrapolated"] = dict(
        N=N, n_tx=n_tx, n_rx=n_rx, M=16, T_p=T_p, T_d=T_d, itera=itera,
        slice="T_d in {1,5}, T_p in {1,41}, QPSK (K=256), 1 iteration each",
        slice_seconds=dict(tp1_td1=t_11, tp1_td5=t_15, tp41_td1=t_41),
        c_pair_s=c_pair, c_pilot_s=c_pilot, c_solve_s=c_solve,
        seconds_per_iteration=per_iter, seconds_per_trial=itera * per_iter,
        trials_per_s=1.0 / (itera * per_iter), kind="extrapolated from the slice")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
