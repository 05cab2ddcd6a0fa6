"""ctypes front end of oracle/quadprog_gi.c with the call signature of ``quadprog.solve_qp``
(SURVEY.md A.8).  CPU ORACLE -- test infrastructure, not product code; parity unpinned."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libquadprog_gi.so")
    src = os.path.join(_HERE, "quadprog_gi.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libquadprog_gi.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        _LIB.qp_solve_gi.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, dp, ip, ip, ip, dp]
        _LIB.qp_solve_gi.restype = ctypes.c_int
    return _LIB


def solve_qp(G, a, C=None, b=None, meq=0, factorized=False):
    """minimise 1/2 x^T G x - a^T x  s.t.  C^T x >= b.  Returns (x, f, xu, iterations, lagrangian, iact)."""
    if factorized:
        raise NotImplementedError("factorized=True is not used by tph")
    G = np.asfortranarray(G, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = G.shape[0]
    if C is None:
        C = np.zeros((n, 0))
        b = np.zeros(0)
    C = np.asfortranarray(C, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    m = C.shape[1]
    x = np.zeros(n)
    lagr = np.zeros(max(m, 1))
    iact = np.zeros(max(m, 1), dtype=np.int32)
    nact = ctypes.c_int(0)
    iters = np.zeros(2, dtype=np.int32)
    fval = ctypes.c_double(0.0)
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int)
    st = _lib().qp_solve_gi(n, m, int(meq), G.ctypes.data_as(dp), a.ctypes.data_as(dp), C.ctypes.data_as(dp),
                            b.ctypes.data_as(dp), x.ctypes.data_as(dp), lagr.ctypes.data_as(dp),
                            iact.ctypes.data_as(ip), ctypes.byref(nact), iters.ctypes.data_as(ip), ctypes.byref(fval))
    if st == 1:
        raise ValueError("constraints are inconsistent, no solution")
    if st == 2:
        raise ValueError("matrix G is not positive definite")
    if st == 3:
        raise RuntimeError("oracle Goldfarb-Idnani solver hit its iteration cap")
    # unconstrained minimiser xu = G^{-1} a
    xu = np.linalg.solve(G, a)
    return x, fval.value, xu, iters, lagr[:m], iact[:nact.value]
