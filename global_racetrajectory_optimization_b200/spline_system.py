"""Stand-in for the dense 4N x 4N spline system matrix that tph.calc_splines returns as ``M`` and that
/root/reference/main_globaltraj.py:264-284 passes back into opt_min_curv / iqp_handler as ``A``.

The matrix only encodes N neighbour scalings: ``scaling_i = -A[4i+2, 4i+5]`` for
i < N-1 and ``scaling_{N-1} = A[4N-2, 1]``.  The CUDA path never forms it; this object carries the
per-segment parameter scales ``h`` (scaling_i = h_i / h_{i+1}) and materialises the dense matrix only when
somebody really converts it to an array."""
from __future__ import annotations

import numpy as np


class SplineSystem:
    def __init__(self, h: np.ndarray):
        self.h = np.ascontiguousarray(h, dtype=np.float64)

    @property
    def no_splines(self) -> int:
        return int(self.h.size)

    @property
    def shape(self):
        return (4 * self.no_splines, 4 * self.no_splines)

    @property
    def scaling(self) -> np.ndarray:
        return self.h / np.roll(self.h, -1)

    def __array__(self, dtype=None, copy=None):
        n = self.no_splines
        sc = self.scaling
        M = np.zeros((4 * n, 4 * n))
        for i in range(n):
            j = 4 * i
            M[j, j] = 1.0
            M[j + 1, j:j + 4] = 1.0
            if i < n - 1:
                M[j + 2, j + 1:j + 4] = (1.0, 2.0, 3.0)
                M[j + 2, j + 5] = -sc[i]
                M[j + 3, j + 2:j + 4] = (2.0, 6.0)
                M[j + 3, j + 6] = -2.0 * sc[i] ** 2
        M[-2, 1] = sc[-1]
        M[-2, -3:] = (-1.0, -2.0, -3.0)
        M[-1, 2] = 2.0 * sc[-1] ** 2
        M[-1, -2:] = (-2.0, -6.0)
        return M if dtype is None else M.astype(dtype)


def h_from_system(A, no_points: int) -> np.ndarray:
    """Per-segment parameter scales from whatever the caller passed as ``A``."""
    if isinstance(A, SplineSystem):
        if A.no_splines != no_points:
            raise RuntimeError("Spline equation system matrix A has wrong dimensions!")
        return A.h
    A = np.asarray(A)
    if A.ndim != 2 or A.shape[0] != 4 * no_points or A.shape[1] != 4 * no_points:
        raise RuntimeError("Spline equation system matrix A has wrong dimensions!")
    n = no_points
    idx = np.arange(n - 1)
    sc = np.empty(n)
    sc[:-1] = -A[4 * idx + 2, 4 * idx + 5]
    sc[-1] = A[4 * n - 2, 1]
    h = np.ones(n)
    h[1:] = 1.0 / np.cumprod(sc[:-1])     # h_{i+1} = h_i / scaling_i  (only ratios matter)
    return h
