"""TEST INFRASTRUCTURE ONLY -- mint golden vectors from the LITERAL reference.

Run in the build container (needs /root/reference, read-only):

    python oracle/make_golden.py            # writes tests/golden/*.npz

Every case seeds numpy's legacy global RNG, calls the reference's own input
generators in the reference drivers' call order, calls the reference's own
estimator, and stores (inputs in this repo's dense layout, reference outputs).
The reference functions are loaded unmodified by oracle/ref_harness.py.
Nothing here is imported by the product.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
TWO_PI = 2 * np.pi


def _nmse(theta_hat, h):
    d = np.asarray(theta_hat).reshape(-1) - np.asarray(h).reshape(-1)
    return float(np.vdot(d, d).real / np.vdot(h, h).real)


def _save(name, meta, arrays):
    os.makedirs(OUT, exist_ok=True)
    flat = {}
    for k, v in meta.items():
        flat["meta_" + k] = np.asarray(v)
    for k, v in arrays.items():
        if v is not None:
            flat[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **flat)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in arrays.items() if v is not None and np.ndim(v) == 0})


def _gen_inputs(ns, order, N, n_tx, n_rx, M, T_p, T_d, varn, has_hinit, insert_pilot_ones=False, qam3=False):
    """Drive the reference's generators in the reference driver's order.
    order 'pm'  : PM.py:174-183     order 'rev4': Proposed method/Proposed_method_NMSEvsTp.py:155-163"""
    h = ns["channelMatrix"](n_tx, n_rx, N, 1)
    out = ns["symbols"](n_tx, M, T_d)
    X_d, aps = out[0], out[1]
    if order == "rev4":
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
    else:
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
    if insert_pilot_ones:
        PsiTilde_tp = np.insert(PsiTilde_tp, 0, np.ones((1, T_p), dtype="complex128"), axis=0)
    PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
    if order != "rev4":
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
    rs = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h, varn, M)
    if has_hinit:
        Y_p, Y_d, Z_p, Z_d, h_initial = rs
    else:
        (Y_p, Y_d, Z_p, Z_d), h_initial = rs, None
    return dict(h=h, X_d=X_d, X_p=X_p, aps=aps, PsiTilde_tp=PsiTilde_tp, PsiTilde_td=PsiTilde_td,
                Y_p=Y_p, Y_d=Y_d, Z_p=Z_p, Z_d=Z_d, h_initial=h_initial)


def _dense(g, n_tx, n_rx):
    return rh.extract_arrays(g["Y_p"], g["Y_d"], g["Z_p"], g["X_p"], g["X_d"], g["PsiTilde_tp"],
                             g["PsiTilde_td"], g["h"], g["h_initial"], n_tx, n_rx)


def case_soft_rev4(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """Soft EM, LS start: `Proposed method/Proposed_method_NMSEvsTp.py:50-83`."""
    ns = rh.load_functions("Proposed method/Proposed_method_NMSEvsTp.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        g = _gen_inputs(ns, "rev4", N, n_tx, n_rx, M, T_p, T_d, varn, True)
        theta = ns["em"](g["Y_d"], g["Y_p"], T_d, T_p, g["Z_p"], g["PsiTilde_td"], g["aps"], M, varn, itera, g["h_initial"])
    d = _dense(g, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="soft", order="rev4", variant="pm", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d,
                itera=itera, varn=varn, src="Proposed method/Proposed_method_NMSEvsTp.py:em")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx),
             nmse_ref=_nmse(theta, g["h"]), nmse_init_ref=_nmse(g["h_initial"], g["h"]),
             norm_ref=float(np.linalg.norm(theta)))
    _save(name, meta, d)
    return d


def case_soft_top(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """Soft EM, zero start, float() weights: `Proposed_method_NMSEvsTp.py:43-69`
    (pilot phases exp(-j2pi t n/T_p) + inserted ones row, :77,:129; C-order h, :14)."""
    ns = rh.load_functions("Proposed_method_NMSEvsTp.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        g = _gen_inputs(ns, "pm", N, n_tx, n_rx, M, T_p, T_d, varn, False, insert_pilot_ones=True)
        theta = ns["em"](g["Y_d"], g["Y_p"], T_d, T_p, g["Z_p"], g["PsiTilde_td"], g["aps"], M, varn, itera)
    d = _dense(g, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="soft", order="pm", variant="top_tp", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d,
                itera=itera, varn=varn, src="Proposed_method_NMSEvsTp.py:em", zero_start=1, h_order="C")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, g["h"]))
    _save(name, meta, d)


def case_nodirect(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """No-direct-link layout: `Proposed method/direct vs non direct - T_pv s nmse.py`, both of its
    estimators on the driver's own data (:159-171): `em` (:82-106; N phase rows, L = N n_tx, zero start) and
    `em_direct_in` (:45-80; ones row inserted, L = (N+1) n_tx)."""
    ns = rh.load_functions("Proposed method/direct vs non direct - T_pv s nmse.py")
    np.random.seed(seed)
    with rh.quiet():
        h_direct_in, h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        X_d, aps = ns["symbols"](n_tx, M, T_d)
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        Y_p, Y_d, Z_p, Z_d = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h, varn, M)
        theta = ns["em"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, n_tx, N)
        Pp1 = np.insert(PsiTilde_tp, 0, np.ones((1, T_p), dtype="complex128"), axis=0)
        Pd1 = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
        Y_p1, Y_d1, Z_p1, Z_d1 = ns["receivedSignals"](T_p, T_d, Pp1, Pd1, n_rx, n_tx, X_d, X_p, h_direct_in, varn, M)
        theta1 = ns["em_direct_in"](Y_d1, Y_p1, T_d, T_p, Z_p1, Pd1, aps, M, varn, itera, N, n_tx)
    d = rh.extract_arrays(Y_p, Y_d, Z_p, X_p, X_d, PsiTilde_tp, PsiTilde_td, h, None, n_tx, n_rx)
    d1 = rh.extract_arrays(Y_p1, Y_d1, Z_p1, X_p, X_d, Pp1, Pd1, h_direct_in, None, n_tx, n_rx)
    meta = dict(kind="nodirect", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d, itera=itera, varn=varn,
                src="Proposed method/direct vs non direct - T_pv s nmse.py:em,em_direct_in", zero_start=1, h_order="C")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(N * n_tx, n_rx), nmse_ref=_nmse(theta, h))
    for k, v in d1.items():
        if v is not None:
            d["din_" + k] = v
    d.update(din_theta_ref=np.asarray(theta1, dtype=np.complex128).reshape((N + 1) * n_tx, n_rx),
             din_nmse_ref=_nmse(theta1, h_direct_in))
    _save(name, meta, d)


def case_parallel(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """Superimposed pilots: `Parallel/ParallelProtocol_Tp.py` em (:64-86) on the driver's own data
    (:119-128): channel, symbols, then per point irsMatrix, pilotSymbols, dataPilotSymbols, receivedSignals."""
    ns = rh.load_functions("Parallel/ParallelProtocol_Tp.py")
    np.random.seed(seed)
    with rh.quiet():
        h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        X_d, aps = ns["symbols"](n_tx, M, T_d)
        T = max(T_d, T_p)
        PsiTilde_t = ns["irsMatrix"](T, N)
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        X = ns["dataPilotSymbols"](n_tx, X_p, X_d)
        Y, Z = ns["receivedSignals"](T, PsiTilde_t, n_rx, n_tx, X, h, varn)
        theta = ns["em"](Y, T, Z, X_d, X_p, T_p, T_d, n_tx, PsiTilde_t, aps, M, varn, itera, N)
    Xoff = np.zeros((T, n_tx), dtype=np.complex128)
    Xoff[:T_p] = np.hstack(X_p).T
    L = (N + 1) * n_tx
    meta = dict(kind="parallel", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d, T=T, itera=itera, varn=varn,
                src="Parallel/ParallelProtocol_Tp.py:em", zero_start=1, h_order="C")
    d = dict(Y=np.hstack(Y).T.copy(), Psi=np.asarray(PsiTilde_t).T.copy(), Xoff=Xoff, Xp=np.hstack(X_p).T.copy(),
             Xd=np.hstack(X_d).T.copy(), h=np.asarray(h).reshape(L, n_rx).copy(),
             theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, h))
    _save(name, meta, d)


def case_hard_llf(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """Hard EM + as-coded LLF: `Proposed method/ML_detecctor.py:51-86`."""
    ns = rh.load_functions("Proposed method/ML_detecctor.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        g = _gen_inputs(ns, "rev4", N, n_tx, n_rx, M, T_p, T_d, varn, True)
        ns["Z_d"] = g["Z_d"]
        theta, llf = ns["em"](g["Y_d"], g["Y_p"], T_d, T_p, g["Z_p"], g["PsiTilde_td"], g["aps"], M, varn, itera, g["h_initial"])
    d = _dense(g, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="hard", order="rev4", variant="pm", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d,
                itera=itera, varn=varn, src="Proposed method/ML_detecctor.py:em")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, g["h"]),
             llf_ref=np.asarray(llf, dtype=np.float64).reshape(-1))
    _save(name, meta, d)


def case_loglik(name, seed, N, n_rx, M, T_p, T_d, itera, varn):
    """`Proposed method/Log_likelihood.py` (named in BASELINE.json north_star): hard EM + as-coded LLF with Z_d and
    n_tx as arguments (:45-85), on the script's own data (:156-166: pilot phases exp(-j2pi t n/N) over all N+1 rows).
    n_tx = 1 only: the script sizes its weight matrix with `M^n_tx` (XOR, quirk Q9), which holds K rows only then."""
    n_tx = 1
    ns = rh.load_functions("Proposed method/Log_likelihood.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        X_d, aps = ns["symbols"](n_tx, M, T_d)
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
        PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        Y_p, Y_d, Z_p, Z_d, h_initial = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h,
                                                              varn, M, N)
        theta, llf = ns["em"](Y_d, Y_p, T_d, T_p, Z_p, Z_d, PsiTilde_td, aps, M, varn, itera, h_initial, n_tx)
    d = rh.extract_arrays(Y_p, Y_d, Z_p, X_p, X_d, PsiTilde_tp, PsiTilde_td, h, h_initial, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="loglik", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d, itera=itera, varn=varn,
                src="Proposed method/Log_likelihood.py:em")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, h),
             llf_ref=np.asarray(llf, dtype=np.float64).reshape(-1))
    _save(name, meta, d)


def case_irs_elements(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """BASELINE.json config 3: `Proposed method/IRS_elements.py` em (:268-304: soft EM from the LS start with the
    trailing N argument, multiprecision weights, genie stop on the global h) on the driver's own data
    (:380-398: pilots and data symbols drawn once, then per N: channel, phases, ones row, received blocks)."""
    ns = rh.load_functions("Proposed method/IRS_elements.py", n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        X_d, aps, _cons = ns["symbols"](n_tx, M, T_d)
        h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
        PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
        Y_p, Y_d, Z_p, Z_d, h_initial = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h,
                                                              varn, M)
        ns["h"] = h
        theta = ns["em"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial, N)
    d = rh.extract_arrays(Y_p, Y_d, Z_p, X_p, X_d, PsiTilde_tp, PsiTilde_td, h, h_initial, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="irs", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d, itera=itera, varn=varn,
                src="Proposed method/IRS_elements.py:em")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, h))
    _save(name, meta, d)


def case_soft_td(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """`Proposed method/Proposed_method_NMSEvsTd.py` em (:44-76) in its driver's order (:141-150: channel, pilots,
    then per T_d symbols, phases, ones row, received blocks)."""
    ns = rh.load_functions("Proposed method/Proposed_method_NMSEvsTd.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        X_d, aps = ns["symbols"](n_tx, M, T_d)
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
        PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
        Y_p, Y_d, Z_p, Z_d, h_initial = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h,
                                                              varn, M)
        theta = ns["em"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial)
    d = rh.extract_arrays(Y_p, Y_d, Z_p, X_p, X_d, PsiTilde_tp, PsiTilde_td, h, h_initial, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="soft_td", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d, itera=itera, varn=varn,
                src="Proposed method/Proposed_method_NMSEvsTd.py:em")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, h))
    _save(name, meta, d)


def case_iterations_llf(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """`Proposed method/IterationsvsLLF.py` em (:44-84: SOFT EM + as-coded LLF, Z_d and n_tx as arguments) on the
    script's own data (:141-150: pilot phases exp(-j2pi t n/N) over N rows plus an inserted ones row)."""
    ns = rh.load_functions("Proposed method/IterationsvsLLF.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        X_d, aps = ns["symbols"](n_tx, M, T_d)
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
        PsiTilde_tp = np.insert(PsiTilde_tp, 0, np.ones((1, T_p), dtype="complex128"), axis=0)
        PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        Y_p, Y_d, Z_p, Z_d, h_initial = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h,
                                                              varn, M, N)
        theta, llf = ns["em"](Y_d, Y_p, T_d, T_p, Z_p, Z_d, PsiTilde_td, aps, M, varn, itera, h_initial, n_tx)
    d = rh.extract_arrays(Y_p, Y_d, Z_p, X_p, X_d, PsiTilde_tp, PsiTilde_td, h, h_initial, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="iter_llf", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d, itera=itera, varn=varn,
                src="Proposed method/IterationsvsLLF.py:em")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, h),
             llf_ref=np.asarray(llf, dtype=np.float64).reshape(-1))
    _save(name, meta, d)


def case_hard_ser(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """Hard EM returning last-iteration decisions + as-coded SER:
    `Proposed method/SER/log_max_SER.py:51-89,162`."""
    ns = rh.load_functions("Proposed method/SER/log_max_SER.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        # driver order log_max_SER.py:150-158: channel, symbols, irsMatrix, insert, pilots, receivedSignals
        g = _gen_inputs(ns, "pm", N, n_tx, n_rx, M, T_p, T_d, varn, True)
        ns["Z_d"] = g["Z_d"]
        theta, X_dest = ns["em"](g["Y_d"], g["Y_p"], T_d, T_p, g["Z_p"], g["PsiTilde_td"], g["aps"], M, varn, itera, g["h_initial"])
    ser = np.count_nonzero(np.array(g["X_d"]) - np.array(X_dest)) / (T_d * n_tx)
    d = _dense(g, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="hard", order="pm", variant="pm", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d,
                itera=itera, varn=varn, src="Proposed method/SER/log_max_SER.py:em")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, g["h"]),
             xdest_ref=np.vstack(X_dest), ser_ref=float(ser))
    _save(name, meta, d)


def case_pm(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn, partition_r, weighted):
    """Partitioned EM. weighted=False: `Proposed method/PM.py:47-116`;
    weighted=True: `Proposed method/PM_beta.py:42-112`."""
    if weighted:
        ns = rh.load_functions("Proposed method/PM_beta.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    else:
        ns = rh.load_functions("Proposed method/PM.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    with rh.quiet():
        h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        out = ns["symbols"](n_tx, M, T_d)  # PM.py returns 3 values, PM_beta.py 2
        X_d, qamCons = out[0], out[-1]
        aps = out[1] if len(out) == 3 else None
        ns["qamCons"] = qamCons
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
        PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        Y_p, Y_d, Z_p, Z_d, h_initial = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h, varn, M)
        if weighted:
            theta = ns["em_pm"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, M, varn, itera, h_initial, h, n_tx, partition_r, X_d, qamCons)
        else:
            theta = ns["em_pm"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial, h, n_tx, partition_r, X_d, qamCons)
    g = dict(h=h, X_d=X_d, X_p=X_p, PsiTilde_tp=PsiTilde_tp, PsiTilde_td=PsiTilde_td, Y_p=Y_p, Y_d=Y_d, Z_p=Z_p, h_initial=h_initial)
    d = _dense(g, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="pm_beta" if weighted else "pm", order="pm", variant="pm", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M,
                T_p=T_p, T_d=T_d, itera=itera, varn=varn, partition_r=partition_r,
                src="Proposed method/%s:em_pm" % ("PM_beta.py" if weighted else "PM.py"))
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, h))
    _save(name, meta, d)


def case_pm_ser(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn, partition_r):
    """`Proposed method/SER/PM_SER.py` em_pm (:55-140): the PM.py estimator started from a RANDOM theta drawn
    inside the function (:57, needs the globals varh and N), `solve` instead of `lstsq`, no genie stop.  The
    random start is captured by replaying the RNG state, so that it can be fed to the restatement."""
    ns = rh.load_functions("Proposed method/SER/PM_SER.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI, varh=1)
    np.random.seed(seed)
    with rh.quiet():
        h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        X_d, aps, qamCons = ns["symbols"](n_tx, M, T_d)
        ns["qamCons"] = qamCons
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
        PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        Y_p, Y_d, Z_p, Z_d, h_initial = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h, varn, M)
        state = np.random.get_state()
        theta_start = np.random.normal(loc=0, scale=np.sqrt(1 / 2), size=(n_rx * n_tx * (N + 1), 1 * 2)).view(np.complex128)
        np.random.set_state(state)
        theta, _xfinal = ns["em_pm"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial, h, n_tx,
                                     partition_r, X_d, qamCons)
    g = dict(h=h, X_d=X_d, X_p=X_p, PsiTilde_tp=PsiTilde_tp, PsiTilde_td=PsiTilde_td, Y_p=Y_p, Y_d=Y_d, Z_p=Z_p,
             h_initial=theta_start)
    d = _dense(g, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="pm_ser", order="pm", variant="pm", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d,
                itera=itera, varn=varn, partition_r=partition_r, src="Proposed method/SER/PM_SER.py:em_pm")
    d.update(theta_ref=np.asarray(theta, dtype=np.complex128).reshape(L, n_rx), nmse_ref=_nmse(theta, h))
    _save(name, meta, d)


def case_multi(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn, partition_r,
               script="Proposed method/PMvsMLvsZFvsMMSE.py"):
    """`Proposed method/PMvsMLvsZFvsMMSE.py`: em / em_ml / em_pm with the genie stop
    active (needs the module global h), same inputs for all (:54-292).  The same five estimators ship again in
    `Proposed method/all_detectorsvsTd.py` and `Proposed method/SNR/all_Detectors.py` (BASELINE.json config 4)."""
    ns = rh.load_functions(script, N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    out = {}
    with rh.quiet():
        h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        X_d, aps, qamCons = ns["symbols"](n_tx, M, T_d)
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        ns["qamCons"] = qamCons
        ns["h"] = h
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
        PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
        Y_p, Y_d, Z_p, Z_d, h_initial = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h, varn, M)
        ns["Z_d"] = Z_d
        out["theta_em"] = ns["em"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial)
        out["theta_ml"] = ns["em_ml"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial)
        out["theta_pm"] = ns["em_pm"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, M, varn, itera, h_initial, h, n_tx, partition_r, X_d, qamCons)
        out["theta_zf"] = ns["em_zf"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial, h)
        out["theta_mmse"] = ns["em_mmse"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial, h)
    g = dict(h=h, X_d=X_d, X_p=X_p, PsiTilde_tp=PsiTilde_tp, PsiTilde_td=PsiTilde_td, Y_p=Y_p, Y_d=Y_d, Z_p=Z_p, h_initial=h_initial)
    d = _dense(g, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="multi", order="multi", variant="pm", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d,
                itera=itera, varn=varn, partition_r=partition_r, src=script)
    for k, v in out.items():
        d[k + "_ref"] = np.asarray(v, dtype=np.complex128).reshape(L, n_rx)
        d["nmse_" + k[6:] + "_ref"] = _nmse(v, h)
    _save(name, meta, d)


def case_detectors(name, seed, N, n_tx, n_rx, M, T_p, T_d, itera, varn):
    """em_zf / em_mmse only (`Proposed method/PMvsMLvsZFvsMMSE.py:54-133`), genie stop active."""
    ns = rh.load_functions("Proposed method/PMvsMLvsZFvsMMSE.py", N=N, n_tx=n_tx, n_rx=n_rx, beta_max=TWO_PI)
    np.random.seed(seed)
    out = {}
    with rh.quiet():
        h = ns["channelMatrix"](n_tx, n_rx, N, 1)
        X_d, aps, qamCons = ns["symbols"](n_tx, M, T_d)
        X_p = ns["pilotSymbols"](n_tx, M, T_p)
        ns["h"] = h
        PsiTilde_tp, PsiTilde_td = ns["irsMatrix"](T_p, T_d, N, 0, 1)
        PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
        Y_p, Y_d, Z_p, Z_d, h_initial = ns["receivedSignals"](T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h, varn, M)
        ns["Z_d"] = Z_d
        out["theta_zf"] = ns["em_zf"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial, h)
        out["theta_mmse"] = ns["em_mmse"](Y_d, Y_p, T_d, T_p, Z_p, PsiTilde_td, aps, M, varn, itera, h_initial, h)
    g = dict(h=h, X_d=X_d, X_p=X_p, PsiTilde_tp=PsiTilde_tp, PsiTilde_td=PsiTilde_td, Y_p=Y_p, Y_d=Y_d, Z_p=Z_p, h_initial=h_initial)
    d = _dense(g, n_tx, n_rx)
    L = (N + 1) * n_tx
    meta = dict(kind="detectors", order="multi", variant="pm", seed=seed, N=N, n_tx=n_tx, n_rx=n_rx, M=M, T_p=T_p, T_d=T_d,
                itera=itera, varn=varn, src="Proposed method/PMvsMLvsZFvsMMSE.py:em_zf,em_mmse")
    for k, v in out.items():
        d[k + "_ref"] = np.asarray(v, dtype=np.complex128).reshape(L, n_rx)
    _save(name, meta, d)


def case_script(name, relpath, seed):
    """A whole top-level script, unmodified, after np.random.seed(seed)."""
    t0 = time.time()
    g = rh.run_script(relpath, seed)
    meta = dict(kind="script", seed=seed, src=relpath, N=g["N"], n_tx=g["n_tx"], n_rx=g["n_rx"], M=g["M_symbols"],
                itera=g["itera"], varn=g["varn"], T_p=np.asarray(g["T_p"]), T_d=np.asarray(g["T_d"]))
    _save(name, meta, dict(mse_ref=np.asarray(g["mse"], dtype=np.float64), seconds=time.time() - t0))


def main(argv):
    if not rh.reference_available():
        raise SystemExit("reference checkout not found; goldens can only be minted in the build container")
    only = set(argv[1:])

    def want(n):
        return not only or n in only

    # known answers of BASELINE.md section 3.2 are reproduced by these two:
    if want("soft_rev4_s1234"):
        d = case_soft_rev4("soft_rev4_s1234", 1234, 8, 2, 2, 4, 12, 40, 3, 0.1)
        print("  BASELINE.md expects NMSE(LS)=0.35400441315416953 NMSE(EM)=0.1116629209833915 ; got",
              d["nmse_init_ref"], d["nmse_ref"])
    if want("hard_llf_s3"):
        case_hard_llf("hard_llf_s3", 3, 8, 2, 2, 4, 12, 40, 3, 0.1)
    if want("soft_top_s5"):
        case_soft_top("soft_top_s5", 5, 6, 2, 2, 4, 16, 30, 4, 0.1)
    if want("soft_top_16qam_s11"):
        case_soft_top("soft_top_16qam_s11", 11, 4, 2, 2, 16, 12, 16, 2, 0.3)
    if want("soft_rev4_1x4_s21"):
        case_soft_rev4("soft_rev4_1x4_s21", 21, 6, 1, 4, 4, 8, 20, 3, 0.2)
    if want("soft_rev4_3x2_s8"):
        case_soft_rev4("soft_rev4_3x2_s8", 8, 3, 3, 2, 4, 14, 16, 2, 0.6)
    if want("hard_ser_s7"):
        case_hard_ser("hard_ser_s7", 7, 6, 2, 2, 4, 10, 24, 3, 1.0)
    if want("pm_s10"):
        case_pm("pm_s10", 10, 6, 2, 2, 4, 20, 20, 4, 0.1, 0, False)
    if want("pm_3x3_s10"):
        case_pm("pm_3x3_s10", 10, 4, 3, 3, 4, 24, 16, 3, 0.1, 2, False)
    if want("pm_beta_s12"):
        case_pm("pm_beta_s12", 12, 6, 2, 2, 4, 20, 20, 3, 0.1, 1, True)
    if want("pm_beta_3x3_s13"):
        case_pm("pm_beta_3x3_s13", 13, 4, 3, 3, 4, 24, 16, 3, 0.5, 2, True)
    if want("multi_s3"):
        case_multi("multi_s3", 3, 8, 2, 2, 4, 12, 40, 3, 0.1, 1)
    if want("multi_td_s4"):
        case_multi("multi_td_s4", 4, 6, 2, 2, 4, 10, 24, 3, 0.2, 1, script="Proposed method/all_detectorsvsTd.py")
    if want("multi_snr_s5"):
        case_multi("multi_snr_s5", 5, 6, 2, 2, 4, 10, 24, 3, 10.0 / 10 ** (12 / 10.0), 1,
                   script="Proposed method/SNR/all_Detectors.py")
    if want("det_16qam_s31"):
        case_detectors("det_16qam_s31", 31, 6, 2, 2, 16, 6, 24, 4, 0.3)
    if want("det_3x3_s32"):
        case_detectors("det_3x3_s32", 32, 5, 3, 3, 4, 5, 20, 4, 0.5)
    if want("det_2x4_s33"):
        case_detectors("det_2x4_s33", 33, 7, 2, 4, 4, 7, 30, 5, 1.0)
    if want("nodirect_s41"):
        case_nodirect("nodirect_s41", 41, 6, 2, 2, 4, 10, 20, 3, 0.1)
    if want("nodirect_1x4_s42"):
        case_nodirect("nodirect_1x4_s42", 42, 8, 1, 4, 16, 12, 24, 3, 0.2)
    if want("parallel_1x4_s51"):
        case_parallel("parallel_1x4_s51", 51, 5, 1, 4, 16, 8, 24, 4, 0.1)
    if want("parallel_2x2_s52"):
        case_parallel("parallel_2x2_s52", 52, 4, 2, 2, 4, 30, 20, 3, 0.2)
    if want("pm_ser_s91"):
        case_pm_ser("pm_ser_s91", 91, 6, 2, 2, 4, 20, 20, 3, 0.5, 1)
    if want("soft_td_s81"):
        case_soft_td("soft_td_s81", 81, 6, 2, 2, 4, 12, 28, 3, 0.1)
    if want("iter_llf_s82"):
        case_iterations_llf("iter_llf_s82", 82, 6, 2, 2, 4, 12, 24, 4, 0.1)
    if want("irs_elements_s71"):
        case_irs_elements("irs_elements_s71", 71, 7, 2, 2, 4, 20, 24, 5, 1.0)
    if want("loglik_s61"):
        case_loglik("loglik_s61", 61, 6, 4, 4, 10, 24, 4, 0.1)
    if want("script_top_td_s0"):
        case_script("script_top_td_s0", "Proposed_method_NMSEvsTd.py", 0)
    if want("script_top_tp_s0"):
        case_script("script_top_tp_s0", "Proposed_method_NMSEvsTp.py", 0)


if __name__ == "__main__":
    main(sys.argv)
