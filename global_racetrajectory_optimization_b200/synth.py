"""Deterministic synthetic closed-track generator.

Produces reftracks ``[x, y, w_tr_right, w_tr_left]`` with N equidistant points on a closed,
star-shaped centre line r(theta) = R (1 + sum_k a_k cos(k theta + phi_k)) whose curvature range is
Berlin-like (max |kappa_ref| around 0.05-0.1 1/m) and whose normals never cross
(max(w * |kappa|) < 0.8 is asserted).  The same function feeds the GPU path, the CPU oracle and
the bench, so that "identical inputs" means bit-identical float64 arrays.
"""
from __future__ import annotations

import numpy as np

STEP_M = 3.0          # point spacing (racecar.ini stepsize_reg, /root/reference/params/racecar.ini:14)
W_LO, W_HI = 3.0, 7.5
LAMBDA_MIN_M = 70.0   # shortest wavelength of the radius function -> K_max = N * step / LAMBDA_MIN_M
AMP = 0.6             # radial amplitude scale a_k = AMP * U(0.3,1) / k^1.5 (max |kappa_ref| ~ 0.05-0.1 1/m)


def _closed_track(n_points: int, rng: np.random.Generator, step: float, amp: float):
    R = n_points * step / (2.0 * np.pi)
    k_max = max(4, int(round(n_points * step / LAMBDA_MIN_M)))
    ks = np.arange(2, k_max + 1)
    # harmonic k perturbs the curvature by ~ a_k k^2 / R; with k_max ~ N the mix keeps the same
    # physical feature size (>= LAMBDA_MIN_M) and a similar curvature range at every N
    a = amp * rng.uniform(0.3, 1.0, ks.size) / ks ** 1.5
    phi = rng.uniform(0.0, 2.0 * np.pi, ks.size)
    n_over = 16 * n_points
    th = np.linspace(0.0, 2.0 * np.pi, n_over, endpoint=False)
    r = R * (1.0 + (a[:, None] * np.cos(ks[:, None] * th[None, :] + phi[:, None])).sum(axis=0))
    xo, yo = r * np.cos(th), r * np.sin(th)
    # arc-length resampling to n_points equidistant points
    dx = np.diff(np.append(xo, xo[0]))
    dy = np.diff(np.append(yo, yo[0]))
    s = np.concatenate(([0.0], np.cumsum(np.hypot(dx, dy))))
    s_new = np.linspace(0.0, s[-1], n_points, endpoint=False)
    x = np.interp(s_new, s, np.append(xo, xo[0]))
    y = np.interp(s_new, s, np.append(yo, yo[0]))
    return x, y, s_new / s[-1]


def _widths(u: np.ndarray, rng: np.random.Generator):
    out = []
    for _ in range(2):
        c = rng.uniform(-1.0, 1.0, 4)
        p = rng.uniform(0.0, 2.0 * np.pi, 4)
        w = sum(c[k] * np.cos(2.0 * np.pi * (k + 1) * u + p[k]) for k in range(4))
        w = (w - w.min()) / max(w.max() - w.min(), 1e-12)
        out.append(W_LO + (W_HI - W_LO) * w)
    return out


def make_track(seed: int, n_points: int, step: float = STEP_M, amp: float = AMP) -> np.ndarray:
    """One reftrack [n_points, 4] (float64), counter-clockwise, equidistant points."""
    rng = np.random.default_rng(seed)
    x, y, u = _closed_track(n_points, rng, step, amp)
    w_r, w_l = _widths(u, rng)
    # keep the normals from crossing inside the track: w * |kappa| <= 0.7 everywhere
    w_cap = 0.7 / np.maximum(np.abs(discrete_curvature(np.column_stack((x, y)))), 1e-9)
    w_r, w_l = np.minimum(w_r, w_cap), np.minimum(w_l, w_cap)
    return np.column_stack((x, y, w_r, w_l))


def make_batch(seed0: int, batch: int, n_points: int, step: float = STEP_M, amp: float = AMP) -> np.ndarray:
    """[batch, n_points, 4]; track i uses seed seed0 + i."""
    return np.stack([make_track(seed0 + i, n_points, step, amp) for i in range(batch)])


def jitter_widths(reftrack: np.ndarray, seed: int, rel: float = 0.1) -> np.ndarray:
    """Smooth multiplicative width jitter w <- w (1 + rel * g(s)), |g| <= 1 (configs C2/C4)."""
    rng = np.random.default_rng(seed)
    n = reftrack.shape[0]
    u = np.arange(n) / n
    out = reftrack.copy()
    for col in (2, 3):
        c = rng.uniform(-1.0, 1.0, 3)
        p = rng.uniform(0.0, 2.0 * np.pi, 3)
        g = sum(c[k] * np.cos(2.0 * np.pi * (k + 1) * u + p[k]) for k in range(3)) / 3.0
        out[:, col] = reftrack[:, col] * (1.0 + rel * g)
    return out


def discrete_curvature(xy: np.ndarray) -> np.ndarray:
    """Menger curvature of the closed polygon (diagnostic for the generator's asserts)."""
    p0, p1, p2 = np.roll(xy, 1, axis=0), xy, np.roll(xy, -1, axis=0)
    a = np.linalg.norm(p1 - p0, axis=1)
    b = np.linalg.norm(p2 - p1, axis=1)
    c = np.linalg.norm(p2 - p0, axis=1)
    cross = (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p1[:, 1] - p0[:, 1]) * (p2[:, 0] - p0[:, 0])
    return 2.0 * cross / (a * b * c)


# ----------------------------------------------------------------------------------------------
# host mirror of csrc/synth.cu (variants generated on the device from a 64-bit seed per variant)
# ----------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def _u01(h: int) -> float:
    return (h >> 11) * (1.0 / 9007199254740992.0)


def jitter_widths_hash(reftrack: np.ndarray, seed: int, rel: float = 0.1) -> np.ndarray:
    """The variant mc_jitter_widths_batch generates on the device for this seed (identical up to the last ulp of cos)."""
    n = reftrack.shape[0]
    u = np.arange(n) / n
    out = reftrack.copy()
    for s in range(2):
        g = np.zeros(n)
        for k in range(3):
            h0 = _splitmix64((seed * 6 + 2 * (3 * s + k)) & _M64)
            h1 = _splitmix64((seed * 6 + 2 * (3 * s + k) + 1 + 0x5851F42D4C957F2D) & _M64)
            g += (2.0 * _u01(h0) - 1.0) * np.cos(2.0 * np.pi * (k + 1) * u + 2.0 * np.pi * _u01(h1))
        out[:, 2 + s] = reftrack[:, 2 + s] * (1.0 + rel * (g * (1.0 / 3.0)))
    return out
