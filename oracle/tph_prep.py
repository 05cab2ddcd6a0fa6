"""CPU ORACLE (test infrastructure, NOT product code) -- restatement of ``tph.check_normals_crossing`` (tph 0.76), the
validity check of the prepared track at /root/reference/helper_funcs_glob/src/prep_track.py:57-59.
PARITY UNPINNED like the other tph restatements (package not vendored, no reference tests); follows the published
algorithm: for every point the normals of the +-horizon neighbours are intersected with its own normal by a 2 x 2 linear
solve; a crossing counts if both intersection parameters lie inside the track widths."""
from __future__ import annotations

import numpy as np


def check_normals_crossing(track: np.ndarray, normvec_normalized: np.ndarray, horizon: int = 10) -> bool:
    no_points = track.shape[0]
    if horizon >= no_points:
        raise RuntimeError("Horizon of %i points is too large for a track with %i points, reduce horizon!"
                           % (horizon, no_points))
    elif horizon >= no_points / 2:
        print("WARNING: Horizon of %i points makes no sense for a track with %i points, reduce horizon!"
              % (horizon, no_points))
    les_mat = np.zeros((2, 2))
    idx_list = list(range(0, no_points))
    idx_list = idx_list[-horizon:] + idx_list + idx_list[:horizon]
    for idx in range(no_points):
        idx_neighbours = idx_list[idx:idx + 2 * horizon + 1]
        del idx_neighbours[horizon]
        idx_neighbours = np.array(idx_neighbours)
        # normals (almost) parallel to the current one cannot cross it
        # (tph: np.cross of 2-D vectors, written out because numpy >= 2.0 deprecates that form)
        n_cur, n_nb = normvec_normalized[idx], normvec_normalized[idx_neighbours]
        is_collinear_b = np.isclose(n_cur[0] * n_nb[:, 1] - n_cur[1] * n_nb[:, 0], 0.0)
        idx_neighbours_rel = idx_neighbours[np.nonzero(np.invert(is_collinear_b))[0]]
        for idx_comp in list(idx_neighbours_rel):
            # p_1 + lambda_1 n_1 = p_2 + lambda_2 n_2
            const = track[idx_comp, :2] - track[idx, :2]
            les_mat[:, 0] = normvec_normalized[idx]
            les_mat[:, 1] = -normvec_normalized[idx_comp]
            lambdas = np.linalg.solve(les_mat, const)
            if -track[idx, 3] <= lambdas[0] <= track[idx, 2] and -track[idx_comp, 3] <= lambdas[1] <= track[idx_comp, 2]:
                return True
    return False


# ----------------------------------------------------------------------------------------------------------------------
# spline_approximation with a Reinsch smoothing spline (checker of csrc/prep_track.cu)
# ----------------------------------------------------------------------------------------------------------------------
# tph.spline_approximation (/root/reference/helper_funcs_glob/src/prep_track.py:39-45) smooths the centre line with
# scipy.interpolate.splprep(k=3, s=s_reg, per=1): FITPACK's adaptive-knot smoothing spline.  The device path keeps every
# statement of tph.spline_approximation except that one call, which it replaces by the periodic cubic smoothing spline of
# Reinsch with the SAME residual budget (sum of squared distances of the data points to the curve, x and y together,
# equal to s_reg; one smoothing parameter for both coordinates, like splprep) and all data points as knots -- an O(N)
# cyclic pentadiagonal problem instead of FITPACK's knot search.  This file states that algorithm in dense numpy;
# tests/test_gpu_prep.py compares the kernel with it (1e-8) and reports the distance to the scipy/FITPACK route.
def _interp_track_cl(track, stepsize):
    """helper_funcs_glob.src.interp_track on the closed track; returns the closed array (last point = first)."""
    track_cl = np.vstack((track, track[0]))
    el = np.sqrt(np.sum(np.power(np.diff(track_cl[:, :2], axis=0), 2), axis=1))
    dists = np.insert(np.cumsum(el), 0, 0.0)
    n = int(np.ceil(dists[-1] / stepsize)) + 1
    di = np.linspace(0.0, dists[-1], n)
    out = np.zeros((n, track_cl.shape[1]))
    for c in range(track_cl.shape[1]):
        out[:, c] = np.interp(di, dists, track_cl[:, c])
    return out, dists


def reinsch_periodic(u, period, xy, s):
    """Periodic cubic smoothing spline through the knots u[0..n-1] (period `period`): minimises
    sum |xy_i - f(u_i)|^2 + lam * int |f''|^2 with lam such that the residual sum equals s (Reinsch 1967).
    Returns (f [n,2], gamma [n,2] = f'' at the knots, lam)."""
    n = u.size
    h = np.diff(np.append(u, u[0] + period))
    Q = np.zeros((n, n))
    R = np.zeros((n, n))
    for i in range(n):
        im, ip = (i - 1) % n, (i + 1) % n
        # (Q^T y)_i = (y_ip - y_i) / h_i - (y_i - y_im) / h_im
        Q[ip, i] += 1.0 / h[i]
        Q[i, i] += -1.0 / h[i] - 1.0 / h[im]
        Q[im, i] += 1.0 / h[im]
        R[i, i] = (h[im] + h[i]) / 3.0
        R[i, ip] += h[i] / 6.0
        R[ip, i] += h[i] / 6.0
    QtQ, Qty = Q.T @ Q, Q.T @ xy

    def resid(lam):
        gam = np.linalg.solve(R + lam * QtQ, Qty)
        r = lam * (Q @ gam)
        return float(np.sum(r * r)), gam, r

    # F(lam) is increasing from 0 (interpolation) to the residual of the best constant; bracket, then bisect on log(lam)
    lo, hi = 1e-12, 1.0
    while resid(hi)[0] < s and hi < 1e30:
        lo, hi = hi, hi * 16.0
    while resid(lo)[0] > s and lo > 1e-300:
        lo, hi = lo / 16.0, lo
    for _ in range(200):
        mid = np.sqrt(lo * hi)
        if resid(mid)[0] < s:
            lo = mid
        else:
            hi = mid
        if hi / lo < 1.0 + 1e-14:
            break
    lam = np.sqrt(lo * hi)
    _, gam, r = resid(lam)
    return xy - r, gam, lam


def eval_periodic_cubic(u, period, f, gam, t):
    """Value of the cubic spline with knots u, values f and second derivatives gam at the parameters t (in [0, period])."""
    n = u.size
    ue = np.append(u, u[0] + period)
    t = np.asarray(t, dtype=float)
    j = np.clip(np.searchsorted(ue, t, side="right") - 1, 0, n - 1)
    jp = (j + 1) % n
    h = ue[j + 1] - ue[j]
    a, b = (ue[j + 1] - t) / h, (t - ue[j]) / h
    return (a[:, None] * f[j] + b[:, None] * f[jp]
            + ((a ** 3 - a) * h * h / 6.0)[:, None] * gam[j] + ((b ** 3 - b) * h * h / 6.0)[:, None] * gam[jp])


def spline_approximation_reinsch(track, k_reg=3, s_reg=10, stepsize_prep=1.0, stepsize_reg=3.0, min_width=None):
    """tph.spline_approximation statement by statement, with the Reinsch smoothing spline in place of splprep, followed by
    prep_track's min-width inflation (prep_track.py:89-98).  Returns reftrack_interp [n, 4]."""
    assert k_reg == 3
    track = np.asarray(track, dtype=float)
    track_interp_cl, dists_cl = _interp_track_cl(track, stepsize_prep)
    pts = track_interp_cl[:-1, :2]
    # splprep's default parametrisation: cumulative chord length of the (closed) data points, normalised to [0, 1]
    ch = np.sqrt(np.sum(np.diff(track_interp_cl[:, :2], axis=0) ** 2, axis=1))
    ucl = np.insert(np.cumsum(ch), 0, 0.0)
    u = ucl[:-1] / ucl[-1]
    f, gam, lam = reinsch_periodic(u, 1.0, pts, float(s_reg))
    n_len = int(np.ceil(dists_cl[-1])) * 4
    tmp = eval_periodic_cubic(u, 1.0, f, gam, np.linspace(0.0, 1.0, n_len))
    length = float(np.sum(np.sqrt(np.sum(np.diff(tmp, axis=0) ** 2, axis=1))))
    n_reg_cl = int(np.ceil(length / stepsize_reg)) + 1
    tq = np.linspace(0.0, 1.0, n_reg_cl)
    path = eval_periodic_cubic(u, 1.0, f, gam, tq)[:-1]
    # closest point on the curve for every point of the (closed) original track: golden-section + parabolic refinement
    from scipy import optimize
    track_cl = np.vstack((track, track[0]))
    n_cl = track_cl.shape[0]
    t_close, d_close, p_close = np.zeros(n_cl), np.zeros(n_cl), np.zeros((n_cl, 2))
    for i in range(n_cl):
        fun = lambda t: float(np.sum((eval_periodic_cubic(u, 1.0, f, gam, [t % 1.0])[0] - track_cl[i, :2]) ** 2))
        t0 = dists_cl[i] / dists_cl[-1]
        span = 4.0 * stepsize_prep / dists_cl[-1]
        cand = t0 + np.linspace(-span, span, 33)
        tb = cand[int(np.argmin([fun(c) for c in cand]))]
        res = optimize.minimize_scalar(fun, bounds=(tb - span / 16.0, tb + span / 16.0), method="bounded", options=dict(xatol=1e-13))
        t_close[i] = res.x
        p_close[i] = eval_periodic_cubic(u, 1.0, f, gam, [res.x % 1.0])[0]
        d_close[i] = np.sqrt(fun(res.x))
    t_close[0], t_close[-1] = 0.0, 1.0
    sides = np.array([np.sign((track_cl[i + 1, 0] - track_cl[i, 0]) * (p_close[i, 1] - track_cl[i, 1])
                              - (track_cl[i + 1, 1] - track_cl[i, 1]) * (p_close[i, 0] - track_cl[i, 0])) for i in range(n_cl - 1)])
    sides_cl = np.hstack((sides, sides[0]))
    w_r = track_cl[:, 2] + sides_cl * d_close
    w_l = track_cl[:, 3] - sides_cl * d_close
    w_r_s = np.interp(tq, t_close, w_r)[:-1]
    w_l_s = np.interp(tq, t_close, w_l)[:-1]
    out = np.column_stack((path, w_r_s, w_l_s))
    if min_width is not None:
        cur = out[:, 2] + out[:, 3]
        add = np.where(cur < min_width, (min_width - cur) / 2.0, 0.0)
        out[:, 2] += add
        out[:, 3] += add
    return out
