#!/usr/bin/env python
"""GPU-side exploration: NMSE of the LS start and of the EM estimate at the north-star size for
several pilot lengths / pilot designs, to pick a bench operating point where EM actually helps."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbce

N, n_tx, n_rx, M, T_d, itera = 64, 4, 4, 16, 256, 10
B = 24
for varn in (0.1, 1.0):
    for variant in ("pm", "top_tp"):
        for T_p in (64, 128, 192, 260, 320):
            tb = sbce.signal_model.generate_batch(N, n_tx, n_rx, M, T_p, T_d, varn, B, seed=5, legacy=False, variant=variant)
            nm0 = np.array([sbce.nmse(tb.theta0[b], tb.h[b]) for b in range(B)])
            out = {}
            for mode in ("soft", "hard"):
                prob = sbce.Problem(N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d, itera=itera, mode=mode)
                t0 = time.time()
                res = sbce.run_host(prob, tb.Yd, tb.Yp, tb.PsiD, tb.PsiP, tb.Xp, tb.varn, theta0=tb.theta0, h_true=tb.h)
                ok = res.status == 0
                out[mode] = (float(np.median(res.nmse[ok])) if ok.any() else float("nan"), int((~ok).sum()), time.time() - t0)
            print("varn=%g variant=%-6s T_p=%3d  NMSE init median %.3e | soft %.3e (flag %d) | hard %.3e (flag %d)" %
                  (varn, variant, T_p, np.median(nm0), out["soft"][0], out["soft"][1], out["hard"][0], out["hard"][1]), flush=True)
