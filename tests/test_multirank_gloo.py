"""world_size-2 `gloo` tests of the multi-GPU plumbing on CPU: trial sharding with disjoint seeds and
the single sum all-reduce of the per-point accumulators.  The numeric runner injected here is the numpy
oracle (test infrastructure standing in for the CUDA library, which needs a GPU); what is under test is
drivers/dist: the 2-rank curves must equal the 1-rank curves."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def oracle_runner(prob, tb, device=0):
    from oracle import em_numpy as orc
    import sbce

    B = tb.Yd.shape[0]
    out = sbce.Result(theta=np.empty_like(tb.h), kstar=np.empty((B, prob.T_d), np.int32), nmse=np.empty(B),
                      status=np.zeros(B, np.int32))
    for b in range(B):
        th, tr = orc.em(tb.Yd[b], tb.Yp[b], tb.PsiD[b], tb.PsiP[b], tb.Xp[b], prob.M, float(tb.varn[b]), prob.itera,
                        theta0=None if prob.zero_start else tb.theta0[b], hard=(prob.mode == "hard"),
                        return_trace=True)
        out.theta[b], out.kstar[b], out.nmse[b] = th, tr["kstar"], orc.nmse(th, tb.h[b])
    return out


def _cfg():
    import sbce

    return sbce.SweepConfig(N=6, n_tx=2, n_rx=2, M=4, T_p=6, T_d=16, itera=2, monte_iter=7, varn=0.2, mode="hard",
                            seed=11, max_batch=3)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sbce

    r, w, _ = sbce.dist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    res = sbce.nmse_vs_tp(_cfg(), [6, 8], runner=oracle_runner)
    ser = sbce.ser_vs_snr(_cfg(), [5.0, 15.0], runner=oracle_runner)
    sbce.dist.barrier()
    if rank == 0:
        q.put((res, ser))
    import torch.distributed as dist

    dist.destroy_process_group()


def test_two_rank_sweep_equals_single_rank():
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    import sbce

    single = sbce.nmse_vs_tp(_cfg(), [6, 8], runner=oracle_runner, keep_per_trial=True)
    single_ser = sbce.ser_vs_snr(_cfg(), [5.0, 15.0], runner=oracle_runner)
    assert single["per_trial"].shape == (7, 2) and np.isfinite(single["per_trial"]).all()
    np.testing.assert_allclose(single["nmse"], single["per_trial"].mean(axis=0), rtol=1e-13)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, ser = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # disjoint shards, same seeds per trial -> identical sums up to summation order
    np.testing.assert_allclose(res["nmse"], single["nmse"], rtol=1e-12)
    assert list(res["n_trials"]) == [7.0, 7.0] and list(res["n_valid"]) == [7.0, 7.0]
    np.testing.assert_allclose(ser["ser"], single_ser["ser"], rtol=1e-12)
    np.testing.assert_allclose(ser["ser_as_coded"], single_ser["ser_as_coded"], rtol=1e-12)


def test_shard_trials_partitions_exactly():
    sys.path.insert(0, ROOT)
    import sbce

    for n in (0, 1, 7, 16, 1000003):
        for w in (1, 2, 3, 8):
            spans = [sbce.dist.shard_trials(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
