"""ctypes wrapper of oracle/banded_cpu.c -- the banded O(N b^2) minimum-curvature QP on one host core (CPU BASELINE /
CHECKER: test infrastructure, only tests/ and bench.py's cpu_baseline legs may use it; see the C file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbanded_cpu.so")
_LIB = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "banded_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=native", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.banded_mincurv_solve.restype = ctypes.c_int
        _LIB.banded_mincurv_solve.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_double] * 2 + [ctypes.c_void_p]
    return _LIB


def spline_h(reftrack: np.ndarray) -> np.ndarray:
    """Parameter scales of tph.calc_splines(use_dist_scaling=True) for a closed track: the segment lengths."""
    p = np.vstack((reftrack[:, :2], reftrack[0, :2]))
    return np.sqrt(np.sum(np.diff(p, axis=0) ** 2, axis=1))


def opt_min_curv_banded(reftrack: np.ndarray, normvec: np.ndarray, w_veh: float, f_scale: float = 2.0, h: np.ndarray = None):
    """Box-only QP of tph.opt_min_curv by the banded algorithm.  Returns (alpha, interior-point iterations)."""
    rt = np.ascontiguousarray(reftrack, dtype=np.float64)
    nv = np.ascontiguousarray(normvec, dtype=np.float64)
    hh = np.ascontiguousarray(spline_h(rt) if h is None else h, dtype=np.float64)
    n = rt.shape[0]
    alpha = np.empty(n)
    it = _lib().banded_mincurv_solve(n, rt.ctypes.data, nv.ctypes.data, hh.ctypes.data, float(w_veh), float(f_scale), alpha.ctypes.data)
    if it == 0:
        raise RuntimeError("Problem not solvable, track might be too small to run with current safety distance!")
    if it < 0:
        raise ValueError("banded_mincurv_solve failed with code %d" % it)
    return alpha, it
