"""CPU tests of the host-side layer: generators, adapters, C-ABI surface, loud failure without a
GPU.  No compute call reaches the CUDA library here."""
import ctypes
import os
import re

import numpy as np
import pytest

import sbce
from oracle import em_numpy as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_of_the_header():
    if not os.path.exists(sbce._lib.LIB_PATH):
        import __graft_entry__ as ge

        ge.build()
    lib = sbce._lib.load()
    header = open(os.path.join(ROOT, "include", "sbce.h")).read()
    declared = set(re.findall(r"\b(sbce_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(sbce._lib.EXPORTS)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert lib.sbce_version() == 100
    assert lib.sbce_error_string(-3).decode().startswith("unsupported")


def test_python_constants_match_header():
    """Every SBCE_MODE_* / SBCE_FLAG_* / SBCE_ST_* value of include/sbce.h equals its twin in _lib.py."""
    header = open(os.path.join(ROOT, "include", "sbce.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+SBCE_((?:MODE|FLAG|ST|PILOTS|PHASES)_[A-Z_]+)\s+(\d+)u?", header)}
    L = sbce._lib
    for name, val in defs.items():
        if name.startswith("PILOTS_"):
            assert L.PILOTS[{"PILOTS_PM": "pm", "PILOTS_TOP": "top"}[name]] == val
        elif name.startswith("PHASES_"):
            assert L.PHASES_KIND[{"PHASES_RANDOM": "random", "PHASES_DFT": "dft"}[name]] == val
        else:
            assert getattr(L, name) == val, name
    assert {"FLAG_SUPERIMPOSED", "FLAG_PSIP_SHARED", "MODE_MMSE", "ST_NOT_PD"} <= set(defs)


def test_struct_layout_matches_header():
    assert ctypes.sizeof(sbce._lib.Cfg) == 16 * 4
    assert ctypes.sizeof(sbce._lib.Io) == 16 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(sbce._lib.Gen) == 48 and sbce._lib.Gen.varh.offset == 24   # sbce_gen (checked against g++)


def test_workspace_query_and_argument_errors_need_no_gpu():
    lib = sbce._lib.load()
    prob = sbce.Problem(N=64, n_tx=4, n_rx=4, M=16, T_p=64, T_d=256, itera=10)
    one = sbce.engine.workspace_bytes(prob, 1)
    many = sbce.engine.workspace_bytes(prob, 100)
    assert 2.0e6 < one < 3.5e6 and 95 * one < many < 105 * one
    assert sbce.engine.workspace_bytes(sbce.Problem(N=8, n_tx=8, n_rx=8, M=4, T_p=8, T_d=8, itera=1), 1) > 0
    for bad in (sbce.Problem(N=8, n_tx=9, n_rx=2, M=4, T_p=8, T_d=8, itera=1),          # more than 8 streams
                sbce.Problem(N=8, n_tx=8, n_rx=8, M=16, T_p=8, T_d=8, itera=1),         # 2^32 joint hypotheses
                sbce.Problem(N=8, n_tx=3, n_rx=5, M=4, T_p=8, T_d=8, itera=1)):
        with pytest.raises(sbce.SbceError):
            sbce.engine.workspace_bytes(bad, 1)
    # ... which the partitioned mode handles (BASELINE.json config 5)
    assert sbce.engine.workspace_bytes(sbce.Problem(N=8, n_tx=8, n_rx=8, M=16, T_p=8, T_d=8, itera=1, mode="pm_beta",
                                                    partition_r=4), 1) > 0
    bad = sbce.Problem(N=8, n_tx=2, n_rx=2, M=8, T_p=8, T_d=8, itera=1)
    with pytest.raises(sbce.SbceError):
        sbce.engine.workspace_bytes(bad, 1)


def test_no_cpu_fallback():
    """Without a CUDA device the estimator must raise, not fall back to the oracle."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    tb = sbce.signal_model.generate_batch(6, 2, 2, 4, 6, 12, 0.1, 2, seed=1)
    prob = sbce.Problem(N=6, n_tx=2, n_rx=2, M=4, T_p=6, T_d=12, itera=2)
    with pytest.raises(sbce.SbceError):
        sbce.run_host(prob, tb.Yd, tb.Yp, tb.PsiD, tb.PsiP, tb.Xp, tb.varn, theta0=tb.theta0)
    import inspect

    for mod in (sbce.engine, sbce.estimators, sbce.drivers, sbce.signal_model, sbce.qam, sbce.dist, sbce._lib):
        src = inspect.getsource(mod)
        assert "import oracle" not in src and "from oracle" not in src, mod.__name__


@pytest.mark.parametrize("order,variant", [("pm", "pm"), ("rev4", "pm"), ("pm", "top_tp"), ("pm", "top_td")])
def test_product_generator_equals_oracle_generator(order, variant):
    """signal_model (product, host side) restates the reference generators independently of the oracle;
    both must draw byte-identical inputs from the same legacy RNG stream."""
    kw = dict(N=7, n_tx=2, n_rx=3, M=16, T_p=7, T_d=11, varn=0.3)
    a = sbce.signal_model.generate_trial(rs=np.random.RandomState(99), order=order, variant=variant, **kw)
    b = orc.gen_trial(rs=np.random.RandomState(99), order=order, variant=variant, **kw)
    for k in ("h", "Xd", "Xp", "PsiP", "PsiD", "Yp", "Yd"):
        assert np.array_equal(a[k], b[k]), k
    assert np.allclose(a["theta0"], b["theta0"], rtol=1e-12, atol=1e-14)


def test_reference_named_generators_follow_reference_call_order():
    """channelMatrix / symbols / irsMatrix / pilotSymbols / receivedSignals with the global RNG."""
    N, n_tx, n_rx, M, T_p, T_d, varn = 6, 2, 2, 4, 6, 9, 0.1
    np.random.seed(5)
    h = sbce.channelMatrix(n_tx, n_rx, N, 1)
    X_d, table, cons = sbce.symbols(n_tx, M, T_d)
    PsiTilde_tp, PsiTilde_td = sbce.irsMatrix(T_p, T_d, N, 0, 1)
    PsiTilde_td = np.insert(PsiTilde_td, 0, np.ones((1, T_d), dtype="complex128"), axis=0)
    X_p = sbce.pilotSymbols(n_tx, M, T_p)
    Y_p, Y_d, Z_p, Z_d, h_initial = sbce.receivedSignals(T_p, T_d, PsiTilde_tp, PsiTilde_td, n_rx, n_tx, X_d, X_p, h,
                                                         varn, M)
    ref = orc.gen_trial(N, n_tx, n_rx, M, T_p, T_d, varn, rs=np.random.RandomState(5), order="pm")
    assert np.array_equal(h.reshape(-1, n_rx), ref["h"])
    assert np.array_equal(np.hstack(X_d).T, ref["Xd"]) and np.array_equal(np.hstack(X_p).T, ref["Xp"])
    assert np.array_equal(PsiTilde_td.T, ref["PsiD"]) and np.array_equal(PsiTilde_tp.T, ref["PsiP"])
    assert np.array_equal(np.hstack(Y_d).T, ref["Yd"]) and np.array_equal(np.hstack(Y_p).T, ref["Yp"])
    assert table.shape == (M ** n_tx, n_tx) and np.array_equal(table, orc.hypothesis_table(cons, n_tx))
    # the implicit design rows behave like the reference's dense Kronecker matrices
    dense = np.kron(np.kron(PsiTilde_tp[:, 2][None, :], X_p[2].T), np.eye(n_rx, dtype="complex128"))
    assert np.array_equal(np.asarray(Z_p[2]), dense) and Z_p[2].shape == dense.shape
    assert np.array_equal(Z_p[2][0, 0::n_rx], dense[0, 0::n_rx])
    assert h_initial.shape == ((N + 1) * n_tx * n_rx, 1)


def test_pilot_factor_recovery_from_dense_Zp():
    """The adapter rebuilds (psi~_t, x_t) from row 0 of the reference's dense Z_p[t]."""
    from importlib import import_module

    est = sbce.estimators
    t = orc.gen_trial(5, 3, 2, 16, 9, 4, 0.1, seed=3, order="pm")
    Wp = orc.design_rows(t["PsiP"], t["Xp"])
    Z_p = [np.kron(Wp[i][None, :], np.eye(2, dtype=np.complex128)) for i in range(9)]
    PsiP, Xp = est._pilot_factors(Z_p, 2, 3, 6)
    assert np.allclose(orc.design_rows(PsiP, Xp), Wp, rtol=1e-15, atol=1e-15)
    PsiP2, Xp2 = est._pilot_factors(Z_p, 2, 3, 6, PsiTilde_tp=t["PsiP"].T, X_p=[x.reshape(-1, 1) for x in t["Xp"]])
    assert np.array_equal(PsiP2, t["PsiP"]) and np.array_equal(Xp2, t["Xp"])


def test_qam_module_matches_oracle():
    for M in (4, 16, 64):
        assert np.array_equal(sbce.qam.constellation(M), orc.qam_constellation(M))
    k = np.arange(4 ** 3)
    assert np.array_equal(sbce.qam.digits_of(k, 4, 3), orc.hypothesis_digits(k, 4, 3))
    tab = orc.hypothesis_table(orc.qam_constellation(16), 2)
    assert np.array_equal(sbce.qam.constellation_from_table(tab, 16), orc.qam_constellation(16))
    with pytest.raises(ValueError):
        sbce.qam.constellation(8)


def test_problem_flags_and_partition():
    p = sbce.Problem(N=8, n_tx=3, n_rx=3, M=4, T_p=8, T_d=8, itera=2, mode="pm_beta", genie_stop=True, quirks=True,
                     partition_r=2)
    c = p.cfg(5)
    assert (c.mode, c.flags, c.partition_p1, c.batch) == (3, 1 | 2, 2, 5)   # int(2/log2 4)+1 = 2 (PM.py:74-75)
    assert sbce.Problem(N=8, n_tx=2, n_rx=2, M=16, T_p=1, T_d=1, itera=1, partition_r=1).p1 == 1
    assert sbce.Problem(N=8, n_tx=2, n_rx=2, M=4, T_p=1, T_d=1, itera=1, zero_start=True, psi_shared=True).cfg(1).flags == 2 | 4 | 8
    assert np.allclose(sbce.drivers.snr_to_varn([-5, 20]), [10 / 10 ** -0.5, 0.1])   # all_Detectors.py:350-354


def test_metrics_helpers():
    X_d = [np.array([[1 + 1j], [1 - 1j]]), np.array([[-1 + 1j], [-1 + 1j]])]
    X_hat = [x.T.copy() for x in X_d]
    # perfect detection still reports 2 cross-stream mismatches / (T_d n_tx = 4) under the reference's broadcast (Q8)
    assert sbce.ser_as_coded(X_d, X_hat) == 0.5
    assert sbce.ser_true(X_d, X_hat) == 0.0
    assert sbce.ser_as_coded(X_d, X_hat) == orc.ser_as_coded(np.hstack(X_d).T, np.vstack(X_hat))


def test_structured_product_helpers():
    from scipy import linalg

    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 5)) + 1j * rng.standard_normal((3, 5))
    b = rng.standard_normal((4, 5)) + 1j * rng.standard_normal((4, 5))
    assert np.array_equal(sbce.signal_model.khatri_rao(a, b), linalg.khatri_rao(a, b))
    m, n = 3, 4
    A = rng.standard_normal((m, n))
    w = sbce.signal_model.commutation_permutation(m, n)
    K = np.eye(m * n)[w, :]                       # what commutation_matrix.py:3-8 builds densely
    assert np.array_equal(K @ A.flatten(order="F"), A.T.flatten(order="F"))
    assert np.array_equal(A.flatten(order="F")[w], A.T.flatten(order="F"))
    # the index map reproduces the dense Kronecker design row
    n_tx, n_rx, N1 = 2, 3, 4
    psi = rng.standard_normal(N1) + 1j * rng.standard_normal(N1)
    x = rng.standard_normal(n_tx) + 1j * rng.standard_normal(n_tx)
    Z = np.kron(np.kron(psi[None, :], x[None, :]), np.eye(n_rx))
    for npr in range(N1):
        for j in range(n_tx):
            for r in range(n_rx):
                assert np.isclose(Z[r, sbce.signal_model.design_index(npr, j, r, n_tx, n_rx)], psi[npr] * x[j], rtol=1e-15, atol=0)


def test_problem_flags_and_shapes_for_the_layout_variants():
    L = sbce._lib
    p = sbce.Problem(N=6, n_tx=2, n_rx=3, M=16, T_p=0, T_d=12, itera=2, superimposed=True, zero_start=True)
    c = p.cfg(5)
    assert c.flags & L.FLAG_SUPERIMPOSED and c.flags & L.FLAG_ZERO_START and not c.flags & L.FLAG_PSI_SHARED
    assert p.shapes(5)["Xp"] == (5, 12, 2) and p.shapes(5)["Yp"] == (5, 0, 3)          # offsets ride in Xp
    p = sbce.Problem(N=6, n_tx=2, n_rx=3, M=16, T_p=8, T_d=12, itera=2, psip_shared=True)
    sh = p.shapes(5)
    assert sh["PsiP"] == (8, 7) and sh["PsiD"] == (5, 12, 7) and p.cfg(5).flags & L.FLAG_PSIP_SHARED
    p = sbce.Problem(N=6, n_tx=2, n_rx=3, M=16, T_p=8, T_d=12, itera=2, psi_shared=True)
    assert p.shapes(5)["PsiP"] == (8, 7) and p.shapes(5)["PsiD"] == (12, 7)
    p = sbce.Problem(N=6, n_tx=2, n_rx=3, M=16, T_p=8, T_d=12, itera=2, mode="zf", zf_stop_guard=True, genie_stop=True)
    assert p.cfg(1).flags & L.FLAG_ZF_STOP_GUARD and p.cfg(1).flags & L.FLAG_GENIE_STOP and p.cfg(1).mode == L.MODE_ZF
    # superimposed pilots need T_p = 0 and a tree mode: rejected by the library's argument check (no GPU needed)
    for bad in (sbce.Problem(N=6, n_tx=2, n_rx=3, M=16, T_p=4, T_d=12, itera=2, superimposed=True),
                sbce.Problem(N=6, n_tx=2, n_rx=3, M=16, T_p=0, T_d=12, itera=2, superimposed=True, mode="pm_beta")):
        with pytest.raises(sbce.SbceError):
            sbce.engine.workspace_bytes(bad, 1)


def test_pilot_design_rows_are_recovered_from_the_dense_Z_p():
    """The reference hands the estimator dense Kronecker matrices Z_p[t] = psi_t^T (x) x_t^T (x) I; the adapter reads
    w_t = psi_t (x) x_t from row 0 and splits it back into (psi, x) -- exact up to a common scalar."""
    from importlib import import_module

    est = import_module(sbce.__name__ + ".estimators")
    rng = np.random.default_rng(3)
    n_tx, n_rx, N1, T_p = 3, 2, 5, 7
    psi = np.exp(2j * np.pi * rng.random((T_p, N1)))
    psi[:, -1] = 0.0                                         # a switched-off element (quirk Q3)
    x = sbce.qam.constellation(16)[rng.integers(0, 16, (T_p, n_tx))]
    Z_p = [np.kron(np.kron(psi[t][None, :], x[t][None, :]), np.eye(n_rx)) for t in range(T_p)]
    PsiP, Xp = est._pilot_factors(Z_p, n_rx, n_tx, N1)
    W = (PsiP[:, :, None] * Xp[:, None, :]).reshape(T_p, -1)
    W_ref = (psi[:, :, None] * x[:, None, :]).reshape(T_p, -1)
    assert np.abs(W - W_ref).max() < 1e-13


def test_em_parallel_validates_shapes_before_touching_the_device():
    cons = sbce.qam.constellation(4)
    table = np.array([[c] for c in cons])
    Y = [np.zeros((2, 1), complex) for _ in range(5)]
    with pytest.raises(ValueError):
        sbce.em_parallel(Y, 6, None, [], [], 0, 5, 1, np.ones((4, 5), complex), table, 4, 0.1, 2, 3)   # T mismatch
    with pytest.raises(ValueError):
        sbce.em_parallel(Y, 5, None, [], [], 0, 5, 1, np.ones((3, 5), complex), table, 4, 0.1, 2, 3)   # N+1 rows expected


def test_finish_reports_nan_ser_for_points_without_decisions():
    """drivers._finish: points whose estimator takes no joint decision (PM modes) have no symbol counts; the SER
    curves must read NaN there instead of 0, the NMSE stays the mean over valid trials (single rank: no collective)."""
    from sbce.drivers import PointResult, _finish

    a = PointResult(nmse_sum=0.6, n_valid=3.0, n_flagged=1.0, sym_err=5.0, sym_total=50.0, ser_coded_sum=0.9, n_trials=4.0)
    b = PointResult(nmse_sum=0.2, n_valid=4.0, n_flagged=0.0, n_trials=4.0)
    out = _finish([10, 20], [a, b], device=None)
    np.testing.assert_allclose(out["nmse"], [0.2, 0.05])
    assert out["ser"][0] == 0.1 and np.isnan(out["ser"][1])
    assert abs(out["ser_as_coded"][0] - 0.225) < 1e-15 and np.isnan(out["ser_as_coded"][1])
    assert list(out["n_flagged"]) == [1.0, 0.0]
