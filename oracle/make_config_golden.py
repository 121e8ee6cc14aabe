"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/config_*.npz: oracle results at the FULL sizes of the
BASELINE.json workloads (sbce.workloads), where the numpy oracle takes minutes per trial and cannot run
inside the GPU test suite.

    python oracle/make_config_golden.py [headline] [c3] [c4] [c41] [c5]

Each fixture stores, for a few trials of the batch `sbce.workloads.make_batch(w, B)` produces (numpy
Generator, seed in the workload), the oracle's theta after all iterations (oracle/em_numpy.py: em / em_pm --
the restatement pinned on the literal reference by tests/test_oracle_golden.py), its decisions, NMSE and
per-iteration log-sums, plus a fingerprint of the inputs (SHA-256 of the integer draws, projections of the
floating arrays) so that a test can prove it regenerated the same inputs before comparing.  The headline fixture uses bench.py's batch (B = 1184, rank 0) and checks the
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
    """SHA-256 over the integer-valued draws of the listed trials (data symbol indices, pilot symbols): proves
    that a regenerated batch sits on the same random stream.  The floating arrays are pinned by input_probes():
    they pass through BLAS (W @ h, pinv) and libm (exp), whose last bits differ between host CPUs."""
    h = hashlib.sha256()
    for b in trials:
        h.update(np.ascontiguousarray(tb.idx_d[b]).astype(np.int64).tobytes())
        h.update(np.ascontiguousarray(tb.Xp[b]).tobytes())
    return h.hexdigest()


PROBED = ("Yd", "Yp", "PsiD", "PsiP", "theta0", "h")


def input_probes(tb, trials):
    """(len(trials), 6, 2) complex: for every probed array a fixed pseudo-random projection and the L1 norm.
    A regenerated array must reproduce the projection to ~1e-11 of the L1 norm (rounding-level differences of
    BLAS / libm only)."""
    out = np.zeros((len(trials), len(PROBED), 2), np.complex128)
    for i, b in enumerate(trials):
        for j, k in enumerate(PROBED):
            a = np.ascontiguousarray(getattr(tb, k)[b]).reshape(-1)
            wts = np.cos(0.37 * np.arange(a.size) + 0.11 * j)          # deterministic weights, no RNG
            out[i, j, 0] = np.dot(wts, a)
            out[i, j, 1] = np.abs(a).sum()
    return out


def check_inputs(tb, trials, meta_digest, probes, rtol=1e-10):
    """Raises AssertionError unless `tb` reproduces the fixture's inputs (see input_digest / input_probes)."""
    assert input_digest(tb, trials) == str(meta_digest), "regenerated batch is on a different random stream"
    got = input_probes(tb, trials)
    err = np.abs(got[:, :, 0] - probes[:, :, 0]) / np.abs(probes[:, :, 1])
    assert err.max() < rtol, "regenerated floating inputs differ from the fixture's beyond rounding: %.2e" % err.max()


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
               meta_digest=np.asarray(input_digest(tb, trials)), probes=input_probes(tb, trials),
               theta_ref=np.stack(thetas), kstar_ref=np.stack(kstars), lse_ref=np.stack(lses),
               nmse_ref=np.asarray(nmses))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print("wrote %s (%.0fs)" % (name, time.time() - t0), flush=True)


def refingerprint(name, w, B, trials):
    """Rewrite the input fingerprint of an existing fixture (oracle outputs untouched)."""
    import sbce

    path = os.path.join(GOLDEN_DIR, name + ".npz")
    z = dict(np.load(path, allow_pickle=False))
    tb = sbce.workloads.make_batch(w, B)
    z["meta_digest"] = np.asarray(input_digest(tb, trials))
    z["probes"] = input_probes(tb, trials)
    np.savez_compressed(path, **z)
    print("refingerprinted", name, flush=True)


def main(argv):
    import sbce

    W = sbce.workloads.WORKLOADS
    if len(argv) > 1 and argv[1] == "--refingerprint":
        refingerprint("config_headline_b1184", W[2], 1184, [0, 592])
        refingerprint("config_3_n256", W[3], 2, [0, 1])
        refingerprint("config_4_8x8qpsk", W[4], 2, [0, 1])
        refingerprint("config_41_64qam_pm", W[41], 2, [0, 1])
        refingerprint("config_5_l2056", W[5], 1, [0])
        return
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
