"""The five workloads BASELINE.json names (`configs[0..4]`), as concrete shapes this repository runs,
tests and benches.  Each entry states the reference script it scales, the estimator mode and a
well-posed operating point (T_p + T_d >= 1.3 L, SURVEY.md section 8d / section 7 hard part 1).

`bench.py --config K` times workload K; `tests/test_gpu_parity.py` checks each against the oracle
(live where the oracle finishes in seconds, otherwise against fixtures minted by
oracle/make_config_golden.py with the same `make_batch` inputs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import engine, signal_model


@dataclass(frozen=True)
class Workload:
    key: int
    name: str
    source: str                  # reference file:line the shape comes from
    N: int
    n_tx: int
    n_rx: int
    M: int
    T_p: int
    T_d: int
    itera: int
    varn: float
    mode: str = "soft"
    partition_r: float = 0.0
    quirks: bool = True
    zero_start: bool = False
    variant: str = "top_tp"      # signal_model pilot / data phase design family
    trials_per_step: int = 1184  # default batch of one bench step on one GPU
    seed: int = 20260
    note: str = ""

    @property
    def L(self):
        return (self.N + 1) * self.n_tx

    def problem(self, psip_shared=True, **over) -> engine.Problem:
        kw = dict(N=self.N, n_tx=self.n_tx, n_rx=self.n_rx, M=self.M, T_p=self.T_p, T_d=self.T_d, itera=self.itera,
                  mode=self.mode, partition_r=self.partition_r, quirks=self.quirks, zero_start=self.zero_start,
                  psip_shared=psip_shared)
        kw.update(over)
        return engine.Problem(**kw)

    def describe(self):
        return ("%s: N=%d RIS, %dx%d MIMO, %d-QAM, T_p=%d, T_d=%d, %d EM iterations, %s EM, %s start, varn=%g"
                % (self.name, self.N, self.n_tx, self.n_rx, self.M, self.T_p, self.T_d, self.itera, self.mode,
                   "zero" if self.zero_start else "LS", self.varn))

    def as_dict(self):
        return dict(N=self.N, n_tx=self.n_tx, n_rx=self.n_rx, M=self.M, T_p=self.T_p, T_d=self.T_d, itera=self.itera,
                    varn=self.varn, mode=self.mode)


def _td_for(L, T_p, floor=32):
    """Smallest multiple of 8 with T_p + T_d >= 1.3 L."""
    return max(floor, int(math.ceil((1.3 * L - T_p) / 8.0)) * 8)


WORKLOADS = {
    # configs[0]: the reference's own CPU-runnable case, at its well-posed sweep point (T_p = 40)
    1: Workload(1, "nmse_vs_tp point (as shipped)", "Proposed_method_NMSEvsTp.py:103-119", N=32, n_tx=2, n_rx=2, M=4,
                T_p=40, T_d=50, itera=10, varn=0.1, mode="soft", zero_start=True, variant="top_tp",
                trials_per_step=9472),
    # configs[1]: the north-star size (bench default; see bench.py's docstring for why T_p = 320)
    2: Workload(2, "nmse_vs_td point", "Proposed_method_NMSEvsTd.py:116-161", N=64, n_tx=4, n_rx=4, M=16, T_p=320,
                T_d=256, itera=10, varn=0.1, mode="soft", variant="top_tp", trials_per_step=1184),
    # configs[2]: IRS_elements.py at its largest N (2x2 QPSK, varn = 1, 5 iterations, T_p = 20 ceil(N/15))
    3: Workload(3, "nmse_vs_N point", "Proposed method/IRS_elements.py:353-430", N=256, n_tx=2, n_rx=2, M=4,
                T_p=20 * math.ceil(256 / 15), T_d=_td_for(257 * 2, 20 * math.ceil(256 / 15)), itera=5, varn=1.0,
                mode="soft", variant="top_tp", trials_per_step=592,
                note="M-step dominated: L = 514, Gram + Cholesky at growing size.  The script's own pilot design "
                     "(PM.py:120-124, period N < T_p) makes pinv() invert rounding noise (tests: "
                     "test_garbage_start_stays_finite); the top-level scripts' design keeps the LS start well-posed"),
    # configs[3]: 8x8 QPSK (K = 65536) hard-decision EM for SER; the 4x4 64-QAM partitioned leg is workload 41
    4: Workload(4, "detectors_vs_snr point, 8x8 QPSK", "Proposed method/PMvsMLvsZFvsMMSE.py:342-416", N=16, n_tx=8,
                n_rx=8, M=4, T_p=160, T_d=64, itera=3, varn=1.0, mode="hard", variant="top_tp", trials_per_step=1184),
    41: Workload(41, "detectors_vs_snr point, 4x4 64-QAM partitioned (4096 candidates)",
                 "Proposed method/all_detectorsvsTd.py:345-418", N=16, n_tx=4, n_rx=4, M=64, T_p=80, T_d=48, itera=2,
                 varn=0.1, mode="pm_beta", partition_r=6, quirks=False, variant="top_tp", trials_per_step=1184),
    # configs[4]: 8x8, N = 256, 16-QAM; 2^32 joint hypotheses per symbol -> only the partitioned estimator applies
    5: Workload(5, "large sweep point, 8x8 16-QAM partitioned", "BASELINE.json configs[4] / Proposed method/PM_beta.py:42-112",
                N=256, n_tx=8, n_rx=8, M=16, T_p=2080, T_d=640, itera=2, varn=0.1, mode="pm_beta", partition_r=4,
                quirks=False, variant="top_tp", trials_per_step=148,
                note="L = 2056: 67 MB normal matrix per trial, Cholesky 1.2e10 flops per iteration"),
}

HEADLINE = WORKLOADS[2]


def make_batch(w: Workload, B: int, seed=None, ls="pinv") -> signal_model.TrialBatch:
    """B synthetic trials of workload `w` from numpy's Generator (one stream for the whole batch), LS start by
    the reference's pinv (PM.py:147).  (The top-level scripts' pilot design repeats the all-ones phase column
    -- direct link and RIS element 0 -- so W_p is structurally rank deficient and only the min-norm pinv
    defines the start; large batches of long channels should use the on-device generator + LS start instead.)"""
    seed = w.seed if seed is None else seed
    return signal_model.generate_batch(w.N, w.n_tx, w.n_rx, w.M, w.T_p, w.T_d, w.varn, B, seed=seed, legacy=False,
                                       variant=w.variant, ls=ls)


def bench_seed(w: Workload, rank: int) -> int:
    """Disjoint generator seeds per rank (bench.py, weak scaling)."""
    return w.seed + 7919 * rank


def host_arrays(w: Workload, tb: signal_model.TrialBatch, psip_shared=True, pinned=False):
    """The input dictionary of one batched call in the layout `Workload.problem(psip_shared)` expects:
    the deterministic pilot design is passed once ([T_p][N+1]) when psip_shared, everything else per trial.
    pinned=True returns page-locked copies (through torch) for the end-to-end route."""
    if psip_shared and not (tb.PsiP == tb.PsiP[0]).all():
        raise ValueError("pilot design differs between trials: cannot share it")
    d = dict(Yd=tb.Yd, Yp=tb.Yp, PsiD=tb.PsiD, PsiP=np.ascontiguousarray(tb.PsiP[0]) if psip_shared else tb.PsiP,
             Xp=tb.Xp, theta0=tb.theta0, h_true=tb.h, varn=tb.varn)
    if pinned:
        import torch

        d = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in d.items()}
    return d
