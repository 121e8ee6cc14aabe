#!/usr/bin/env python
"""How many hard decisions / posterior statistics change from one EM iteration to the next at the
bench operating point?  (motivates delta updates of the normal matrix and fixed-point early exit)"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbce

for varn in (0.1, 1.0, 3.16):
    N, n_tx, n_rx, M, T_p, T_d = 64, 4, 4, 16, 320, 256
    B = 64
    tb = sbce.signal_model.generate_batch(N, n_tx, n_rx, M, T_p, T_d, varn, B, seed=5, legacy=False, variant="top_tp")
    prob1 = sbce.Problem(N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d, itera=1)
    ses = sbce.DeviceSession(prob1, B)
    dev = ses.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Yd, Yp, PsiD, PsiP, Xp, vn, h = t(tb.Yd), t(tb.Yp), t(tb.PsiD), t(tb.PsiP), t(tb.Xp), t(tb.varn), t(tb.h)
    theta = t(tb.theta0)
    prev = None
    for l in range(10):
        m, R, ks, ls = ses.estep(Yd, PsiD, theta, vn)
        stats = torch.cat([torch.view_as_real(m).reshape(B, T_d, -1), torch.view_as_real(R).reshape(B, T_d, -1)], dim=2)
        if prev is not None:
            ch = (stats != prev).any(dim=2)                     # (B, T_d) bitwise change of any statistic
            per_trial = ch.sum(dim=1)
            print("varn=%g iter %d: changed symbols mean %.2f / %d (%.2f%%), trials with zero changes %d / %d" %
                  (varn, l, per_trial.float().mean().item(), T_d, 100 * ch.float().mean().item(), int((per_trial == 0).sum()), B), flush=True)
        prev = stats
        theta, st = ses.mstep(Yd, Yp, PsiD, PsiP, Xp, m, R)
    nm = ((theta - h).abs() ** 2).sum(dim=(1, 2)) / (h.abs() ** 2).sum(dim=(1, 2))
    print("   final NMSE mean %.3e" % nm.mean().item())
