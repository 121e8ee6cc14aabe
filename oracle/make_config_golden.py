"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/config_*.npz: oracle results at the FULL sizes of the
BASELINE.json workloads (sbce.workloads), where the numpy oracle takes minutes per trial and cannot run
inside the GPU test suite.

    python oracle/make_config_golden.py [headline] [c3] [c4] [c41] [c5]

Each fixture stores, for a few trials of the batch `sbce.workloads.make_batch(w, B)` produces (numpy
Generator, seed in the workload), the oracle's theta after all iterations (oracle/em_numpy.py: em / em_pm --
the restatement pinned on the literal reference by tests/test_oracle_golden.py), its decisions, NMSE and
per-iteration log-sums, plus a SHA-256 of the input arrays so that a test can prove it regenerated the very
same inputs before comparing.  The headline fixture uses bench.py's batch (B = 1184, rank 0) and checks the
first trial of each half of the host route's two-half pipeline (trials 0 and 592).
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import em_numpy as orc  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def input_digest(tb, trials):
    """SHA-256 over the inputs of the listed trials (exactly the arrays both sides consume)."""
    h = hashlib.sha256()
    for b in trials:
        for k in ("Yd", "Yp", "PsiD", "PsiP", "Xp", "theta0", "h"):
            h.update(np.ascontiguousarray(getattr(tb, k)[b]).tobytes())
    return h.hexdigest()


def run_oracle(w, tb, b):
    theta0 = None if w.zero_start else tb.theta0[b]
    if w.mode in ("soft", "hard"):
        th, tr = orc.em(tb.Yd[b], tb.Yp[b], tb.PsiD[b], tb.PsiP[b], tb.Xp[b], w.M, w.varn, w.itera, theta0=theta0,
                        hard=(w.mode == "hard"), return_trace=True)
        return th, np.asarray(tr["kstar"], np.int32), np.asarray(tr["lse"], np.float64)
    th = orc.em_pm(tb.Yd[b], tb.Yp[b], tb.PsiD[b], tb.PsiP[b], tb.Xp[b], w.M, w.varn, w.itera, tb.theta0[b],
                   h_true=tb.h[b], partition_r=w.partition_r, weighted=(w.mode == "pm_beta"), genie_stop=False,
                   quirks=w.quirks, how="solve")
    return th, np.zeros((0,), np.int32), np.zeros((0,), np.float64)


def mint(name, w, B, trials):
    import sbce

    t0 = time.time()
    tb = sbce.workloads.make_batch(w, B)
    thetas, kstars, lses, nmses = [], [], [], []
    for b in trials:
        th, ks, ls = run_oracle(w, tb, b)
        thetas.append(th)
        kstars.append(ks)
        lses.append(ls)
        nmses.append(orc.nmse(th, tb.h[b]))
        print("  %s trial %d: nmse %.6e (start %.3e)  %.0fs" % (name, b, nmses[-1], orc.nmse(tb.theta0[b], tb.h[b]),
                                                                time.time() - t0), flush=True)
    out = dict(meta_kind=np.asarray("config"), meta_workload=np.asarray(w.key), meta_B=np.asarray(B),
               meta_trials=np.asarray(trials), meta_seed=np.asarray(w.seed), meta_desc=np.asarray(w.describe()),
               meta_digest=np.asarray(input_digest(tb, trials)),
               theta_ref=np.stack(thetas), kstar_ref=np.stack(kstars), lse_ref=np.stack(lses),
               nmse_ref=np.asarray(nmses))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print("wrote %s (%.0fs)" % (name, time.time() - t0), flush=True)


def main(argv):
    import sbce

    W = sbce.workloads.WORKLOADS
    only = set(argv[1:])
    want = lambda n: not only or n in only
    if want("headline"):
        mint("config_headline_b1184", W[2], 1184, [0, 592])
    if want("c3"):
        mint("config_3_n256", W[3], 2, [0, 1])
    if want("c4"):
        mint("config_4_8x8qpsk", W[4], 2, [0, 1])
    if want("c41"):
        mint("config_41_64qam_pm", W[41], 2, [0, 1])
    if want("c5"):
        mint("config_5_l2056", W[5], 1, [0])


if __name__ == "__main__":
    main(sys.argv)
