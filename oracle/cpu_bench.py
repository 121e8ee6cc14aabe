"""TEST INFRASTRUCTURE ONLY -- times the numpy oracle (the CPU port of the
reference's hot path) on the host cores.  Used by bench.py's `cpu_baseline`
leg and by `bench.py --impl reference`; never by the product.

The literal reference is interpreted Python at O(T_d*K*n_rx*D^2) per iteration
(weeks per trial at the north-star size; oracle/time_literal_reference.py measures a
slice of it in the build container) and cannot travel to
the GPU box, so the baseline is kind="port": oracle/em_numpy.py, one process per
core, one BLAS thread each, on a BOUNDED sample -- `sample_iters` EM iterations of
one trial per worker, started from the reference's LS start (pinv, PM.py:147) like
the GPU arm -- scaled linearly to the full iteration count."""
from __future__ import annotations

import os
import time

import numpy as np


def _worker(args):
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        threadpool_limits = None
    # one BLAS thread per worker for the WHOLE worker: input generation (pinv, W @ Theta) with the default thread
    # count oversubscribes the cores 16-fold and took 45 s of wall time per sample before the clock even started
    if threadpool_limits is not None:
        with threadpool_limits(limits=1):
            return _worker_body(args)
    return _worker_body(args)


def _worker_body(args):
    (seed, w, iters) = args
    from oracle import em_numpy as orc

    N, n_tx, n_rx, M, T_p, T_d, varn = w["N"], w["n_tx"], w["n_rx"], w["M"], w["T_p"], w["T_d"], w["varn"]
    mode = w.get("mode", "soft")
    rs = np.random.RandomState(seed)
    # synthetic inputs with the shape and statistics of the workload (generation is not timed)
    Th = orc.channel_vector(n_tx, n_rx, N, 1.0, rs)
    cons = orc.qam_constellation(M)
    Xd = cons[rs.randint(0, M, (T_d, n_tx))]
    Xp = cons[rs.randint(0, M, (T_p, n_tx))]
    PsiD = np.ones((T_d, N + 1), np.complex128)
    PsiD[:, 1:] = np.exp(1j * rs.uniform(0, 2 * np.pi, (T_d, N)))
    PsiP = np.ones((T_p, N + 1), np.complex128)
    PsiP[:, 1:] = np.exp(1j * rs.uniform(0, 2 * np.pi, (T_p, N)))
    nz = lambda T: (rs.standard_normal((T, n_rx)) + 1j * rs.standard_normal((T, n_rx))) * np.sqrt(varn / 2)
    Wp = orc.design_rows(PsiP, Xp)
    Yp = Wp @ Th + nz(T_p)
    Yd = orc.design_rows(PsiD, Xd) @ Th + nz(T_d)
    theta0 = None if w.get("zero_start") else np.linalg.pinv(Wp) @ Yp      # h_initial, PM.py:147

    t0 = time.perf_counter()
    if mode in ("soft", "hard"):
        orc.em(Yd, Yp, PsiD, PsiP, Xp, M, varn, iters, theta0=theta0, hard=(mode == "hard"))
    elif mode in ("pm", "pm_beta"):
        orc.em_pm(Yd, Yp, PsiD, PsiP, Xp, M, varn, iters, theta0, h_true=Th, partition_r=w.get("partition_r", 0),
                  weighted=(mode == "pm_beta"), genie_stop=False, quirks=w.get("quirks", True), how="solve")
    else:
        orc.em_detector(Yd, Yp, PsiD, PsiP, Xp, M, varn, iters, theta0, kind=mode, h_true=Th, genie_stop=False,
                        quirks=w.get("quirks", True))
    return time.perf_counter() - t0


def time_sample(w, sample_iters=2, workers=None, seed=1234):
    """Run `workers` processes, each `sample_iters` EM iterations of one trial of workload dict `w`
    (keys N, n_tx, n_rx, M, T_p, T_d, varn, itera, mode[, partition_r, quirks, zero_start]).
    Returns dict(trials_per_s (scaled to w['itera'] iterations), cores, seconds, sample)."""
    import multiprocessing as mp

    itera = int(w["itera"])
    sample_iters = max(1, min(int(sample_iters), itera))
    ncpu = os.cpu_count() or 1
    workers = max(1, min(ncpu, 32) if workers is None else workers)
    args = [(seed + i, dict(w), sample_iters) for i in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        per = [_worker(args[0])]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            per = pool.map(_worker, args)
    wall = time.perf_counter() - t0
    slowest = max(per)
    # throughput of the sample: `workers` trial-slices finished in `slowest` seconds, each slice is
    # sample_iters/itera of a trial
    trials_per_s = workers * (sample_iters / float(itera)) / slowest
    return dict(trials_per_s=trials_per_s, cores=workers, seconds=wall, slowest_worker_s=slowest,
                sample_iters=sample_iters,
                sample="%d worker(s) x 1 trial x %d of %d EM iterations from the %s start (full T_d=%d, %d-QAM, "
                       "n_tx=%d, mode %s), scaled x%g"
                       % (workers, sample_iters, itera, "zero" if w.get("zero_start") else "LS", w["T_d"], w["M"],
                          w["n_tx"], w.get("mode", "soft"), itera / float(sample_iters)))


def main(argv=None):
    """CLI used by bench.py (a FRESH process: forking the GPU process with its pinned buffers and CUDA context
    costs tens of seconds per pool):  python -m oracle.cpu_bench '<workload json>' <sample_iters> [seed]"""
    import json
    import sys

    argv = sys.argv[1:] if argv is None else argv
    w = json.loads(argv[0])
    r = time_sample(w, sample_iters=int(argv[1]), workers=None, seed=int(argv[2]) if len(argv) > 2 else 1234)
    print(json.dumps(r))


if __name__ == "__main__":
    main()
