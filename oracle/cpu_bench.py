"""TEST INFRASTRUCTURE ONLY -- times the numpy oracle (the CPU port of the
reference's hot path) on the host cores.  Used by bench.py's `cpu_baseline`
leg and by `bench.py --impl reference`; never by the product.

The literal reference is interpreted Python at O(T_d*K*n_rx*D^2) per iteration
(hours per trial at the north-star size) and cannot travel to the GPU box, so
the baseline is kind="port": oracle/em_numpy.py, one process per core, one BLAS
thread each, on a BOUNDED sample (a fixed number of EM iterations of one trial
per worker), scaled linearly to the full iteration count."""
from __future__ import annotations

import os
import time

import numpy as np


def _worker(args):
    (seed, N, n_tx, n_rx, M, T_p, T_d, varn, iters, hard) = args
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        threadpool_limits = None
    from oracle import em_numpy as orc

    rs = np.random.RandomState(seed)
    # light-weight synthetic inputs (shape and statistics of the workload; generation is not timed)
    Th = orc.channel_vector(n_tx, n_rx, N, 1.0, rs)
    cons = orc.qam_constellation(M)
    Xd = cons[rs.randint(0, M, (T_d, n_tx))]
    Xp = cons[rs.randint(0, M, (T_p, n_tx))]
    PsiD = np.ones((T_d, N + 1), np.complex128)
    PsiD[:, 1:] = np.exp(1j * rs.uniform(0, 2 * np.pi, (T_d, N)))
    PsiP = np.ones((T_p, N + 1), np.complex128)
    PsiP[:, 1:] = np.exp(1j * rs.uniform(0, 2 * np.pi, (T_p, N)))
    nz = lambda T: (rs.standard_normal((T, n_rx)) + 1j * rs.standard_normal((T, n_rx))) * np.sqrt(varn / 2)
    Yp = orc.design_rows(PsiP, Xp) @ Th + nz(T_p)
    Yd = orc.design_rows(PsiD, Xd) @ Th + nz(T_d)
    theta0 = Th + 0.05 * (rs.standard_normal(Th.shape) + 1j * rs.standard_normal(Th.shape))

    def run():
        t0 = time.perf_counter()
        orc.em(Yd, Yp, PsiD, PsiP, Xp, M, varn, iters, theta0=theta0, hard=hard)
        return time.perf_counter() - t0

    if threadpool_limits is not None:
        with threadpool_limits(limits=1):
            return run()
    return run()


def time_sample(N, n_tx, n_rx, M, T_p, T_d, varn, itera, sample_iters=1, workers=None, hard=False, seed=1234):
    """Run `workers` processes, each `sample_iters` EM iterations of one trial.
    Returns dict(trials_per_s (scaled to `itera` iterations), cores, seconds, sample)."""
    import multiprocessing as mp

    ncpu = os.cpu_count() or 1
    workers = max(1, min(ncpu, 32) if workers is None else workers)
    args = [(seed + i, N, n_tx, n_rx, M, T_p, T_d, varn, sample_iters, hard) for i in range(workers)]
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
                sample="%d worker(s) x 1 trial x %d of %d EM iterations (full T_d=%d, K=%d), scaled x%g"
                       % (workers, sample_iters, itera, T_d, M ** n_tx, itera / float(sample_iters)))
