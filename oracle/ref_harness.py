"""TEST INFRASTRUCTURE ONLY -- loader for the *literal* reference functions.

This module executes the unmodified reference scripts' function definitions
(from /root/reference, read-only) so that golden vectors can be minted from the
reference itself.  It only works in the build container (the GPU box has no
/root/reference); nothing in the product, in `-m gpu` tests, in smoke() or in
bench.py imports it.  `oracle/make_golden.py` is its only caller besides the
`not gpu` tests that re-validate the numpy restatement when the reference is
present.

How the reference is made importable without modifying it:
  * numpy aliases removed in numpy>=1.24 are restored (QAM.py:320-321 uses
    np.int / np.complex);
  * `matplotlib`, `matplotlib.pyplot` and `gmpy2` are stubbed
    (gmpy2.exp -> mpmath.exp: both are 53-bit mantissa, unbounded exponent);
  * a script is parsed with `ast`, only FunctionDef / Import / ImportFrom nodes
    are kept (the Monte-Carlo loops that run at import time are dropped) and
    the result is exec'd into a namespace pre-seeded with the module globals the
    functions silently capture (N, n_tx, n_rx, beta_max, h, Z_d, qamCons ...).
"""
from __future__ import annotations

import ast
import contextlib
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
_PM_DIR = os.path.join(REFERENCE_ROOT, "Proposed method")


def reference_available() -> bool:
    return os.path.isdir(_PM_DIR)


def _install_shims() -> None:
    for name, typ in (("int", int), ("float", float), ("complex", complex)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")

        def _noop(*a, **k):
            return None

        for fn in ("plot", "grid", "ylabel", "xlabel", "xticks", "yticks", "yscale",
                   "xscale", "title", "legend", "show", "figure", "savefig", "semilogy"):
            setattr(plt, fn, _noop)
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "gmpy2" not in sys.modules:
        import mpmath

        gp = types.ModuleType("gmpy2")
        gp.exp = mpmath.exp
        sys.modules["gmpy2"] = gp
    if _PM_DIR not in sys.path:
        sys.path.insert(0, _PM_DIR)


def load_functions(relpath: str, **implicit_globals):
    """Return a namespace dict holding the functions defined in the reference
    script `relpath` (relative to /root/reference).  `implicit_globals` are the
    module-level names the functions read (e.g. N=8, n_tx=2, beta_max=2*pi)."""
    if not reference_available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _install_shims()
    path = os.path.join(REFERENCE_ROOT, relpath)
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.Import, ast.ImportFrom))]
    mod = ast.Module(body=keep, type_ignores=[])
    ns: dict = {"__name__": "ref_" + os.path.basename(relpath)}
    ns.update(implicit_globals)
    exec(compile(mod, path, "exec"), ns)
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints every iteration; swallow it."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


def run_script(relpath: str, seed: int):
    """Run a whole reference script unmodified (runpy) after np.random.seed(seed);
    returns its globals (e.g. `mse`)."""
    import runpy

    _install_shims()
    np.random.seed(seed)
    with quiet():
        return runpy.run_path(os.path.join(REFERENCE_ROOT, relpath))


# ---------------------------------------------------------------------------
# Helpers that turn the reference's Python-list objects into dense arrays in
# the layout the numpy oracle and the CUDA library use.
# ---------------------------------------------------------------------------

def extract_arrays(Y_p, Y_d, Z_p, X_p, X_d, PsiTilde_tp, PsiTilde_td, h, h_initial, n_tx, n_rx):
    """Yd (T_d,n_rx), Yp (T_p,n_rx), Xp (T_p,n_tx), Xd (T_d,n_tx),
    PsiP (T_p,N+1), PsiD (T_d,N+1), theta0 (L,n_rx), h (L,n_rx), Wp (T_p,L)."""
    Yd = np.hstack(Y_d).T.copy()
    Yp = np.hstack(Y_p).T.copy()
    Xp = np.hstack(X_p).T.copy()
    Xd = np.hstack(X_d).T.copy()
    PsiP = np.asarray(PsiTilde_tp).T.copy()
    PsiD = np.asarray(PsiTilde_td).T.copy()
    L = PsiD.shape[1] * n_tx
    h = np.asarray(h).reshape(L, n_rx).copy()
    theta0 = None if h_initial is None else np.asarray(h_initial, dtype=np.complex128).reshape(L, n_rx).copy()
    Wp = np.stack([np.asarray(Z_p[t])[0, 0::n_rx] for t in range(len(Z_p))])
    return dict(Yd=Yd, Yp=Yp, Xp=Xp, Xd=Xd, PsiP=PsiP, PsiD=PsiD, theta0=theta0, h=h, Wp=Wp)
