"""CPU tests of the oracle itself (-m "not gpu"): the dense tph restatement and the Goldfarb-Idnani
solver are pinned against the committed golden vectors (regression), against an independent exact
solver (scipy BVLS), against the structural identities of SURVEY.md 8c and against analytic cases.
PARITY UNPINNED vs the real tph/quadprog (absent offline) -- these tests are the substitute pinning."""
import numpy as np
import pytest
from scipy.optimize import lsq_linear

from conftest import rel_max
from oracle import quadprog_gi
from oracle import tph_dense as T
from global_racetrajectory_optimization_b200 import synth

SMALL = ["synth128", "synth200", "handling"]


@pytest.mark.parametrize("name", SMALL + ["synth333"])
def test_oracle_reproduces_golden(golden, name):
    g = golden(name)
    rt = g["reftrack"]
    path = np.vstack((rt[:, :2], rt[0, :2]))
    cx, cy, A, nv = T.calc_splines(path)
    assert np.allclose(cx, g["coeffs_x"], rtol=0, atol=1e-9)
    assert np.allclose(nv, g["normvec"], rtol=0, atol=1e-12)
    alpha, cerr = T.opt_min_curv(rt, nv, A, float(g["kappa_bound"]), float(g["w_veh"]))
    assert rel_max(alpha, g["alpha_mincurv"]) < 1e-7
    assert abs(cerr - float(g["curv_error_max"])) < 1e-7
    assert rel_max(T.opt_shortest_path(rt, nv, float(g["w_veh"])), g["alpha_shpath"]) < 1e-9


@pytest.mark.parametrize("name", SMALL)
def test_structural_identities(golden, name):
    g = golden(name)
    rt, nv = g["reftrack"], g["normvec"]
    path = np.vstack((rt[:, :2], rt[0, :2]))
    _, _, A, _ = T.calc_splines(path)
    qp = T.assemble_min_curv(rt, nv, A, float(g["kappa_bound"]), float(g["w_veh"]))
    E, k = qp["E_kappa"], qp["k_kappa_ref"]
    assert np.abs(qp["H"] - E.T @ E).max() <= 1e-12 * np.abs(qp["H"]).max()          # H = E^T E
    assert np.abs(qp["f"] - T.F_SCALE * E.T @ k).max() <= 1e-11 * np.abs(qp["f"]).max()   # f = F_SCALE E^T k_ref
    # scalings encoded in A
    n = rt.shape[0]
    sc = np.array([-A[4 * i + 2, 4 * i + 5] for i in range(n - 1)] + [A[4 * n - 2, 1]])
    assert np.allclose(sc, g["el_lengths"] / np.roll(g["el_lengths"], -1), rtol=1e-13)
    # moment form of the spline: a1 = dp - h^2 (2 m_i + m_{i+1}) / 6 with 2 a2 = h^2 m
    h = g["el_lengths"]
    m = 2.0 * g["coeffs_x"][:, 2] / h ** 2
    a1 = (np.roll(rt[:, 0], -1) - rt[:, 0]) - h ** 2 * (2 * m + np.roll(m, -1)) / 6
    assert np.abs(a1 - g["coeffs_x"][:, 1]).max() < 1e-10


@pytest.mark.parametrize("name", SMALL + ["berlin", "synth1000"])
def test_golden_alpha_satisfies_kkt(golden, name):
    """KKT residuals of the stored optimum, using only the stored band of H (no dense re-assembly)."""
    g = golden(name)
    rt = g["reftrack"]
    n = rt.shape[0]
    a = g["alpha_mincurv_boxonly"]
    HB = g["H_band"]
    Ha = HB[:, 0] * a
    for k in range(1, 33):
        Ha += HB[:, k] * np.roll(a, -k) + np.roll(HB[:, k] * a, k)
    grad = Ha + g["f"]
    ub = rt[:, 2] - float(g["w_veh"]) / 2
    lb = -(rt[:, 3] - float(g["w_veh"]) / 2)
    assert np.all(a <= ub + 1e-9) and np.all(a >= lb - 1e-9)
    scale = np.abs(g["f"]).max()
    at_ub, at_lb = np.abs(a - ub) < 1e-8, np.abs(a - lb) < 1e-8
    free = ~(at_ub | at_lb)
    assert np.abs(grad[free]).max() <= 2e-6 * scale          # stationarity on the free set (band-truncation level)
    assert np.all(grad[at_ub] <= 1e-6 * scale)               # multipliers have the right sign
    assert np.all(grad[at_lb] >= -1e-6 * scale)
    assert not bool(g["kappa_rows_active"])
    assert rel_max(g["alpha_mincurv"], a) < 1e-9


def test_goldfarb_idnani_vs_bvls_and_kkt():
    rng = np.random.default_rng(5)
    n = 70
    E = rng.standard_normal((n, n))
    k = rng.standard_normal(n)
    H, f = E.T @ E, 2 * E.T @ k
    lb, ub = -rng.uniform(0.1, 1, n), rng.uniform(0.1, 1, n)
    G = np.vstack((np.eye(n), -np.eye(n)))
    h = np.append(ub, -lb)
    x, fval, xu, iters, lagr, iact = quadprog_gi.solve_qp(H, -f, -G.T, -h, 0)
    ref = lsq_linear(E, -2 * k, bounds=(lb, ub), method="bvls", tol=1e-15).x
    assert np.abs(x - ref).max() < 1e-10
    r = H @ x + f + G.T @ lagr
    assert np.abs(r).max() < 1e-9 and lagr.min() >= 0
    assert abs(fval - (0.5 * x @ H @ x + f @ x)) < 1e-9
    # general inequality rows, infeasible problem, non-PD matrix
    with pytest.raises(ValueError, match="inconsistent"):
        quadprog_gi.solve_qp(np.eye(2), np.zeros(2), np.array([[1.0, -1.0], [0.0, 0.0]]), np.array([1.0, 1.0]))
    with pytest.raises(ValueError, match="positive definite"):
        quadprog_gi.solve_qp(np.array([[1.0, 2.0], [2.0, 1.0]]), np.zeros(2))


def _circle_track(n, radius, w):
    th = np.linspace(0, 2 * np.pi, n, endpoint=False)
    return np.column_stack((radius * np.cos(th), radius * np.sin(th), np.full(n, w), np.full(n, w)))


def test_circle_is_symmetric():
    """Circle with symmetric widths: alpha is constant by symmetry.  With the first derivatives frozen
    (tph's linearisation) the linearised curvature of a uniformly shifted circle is (R + alpha) / R^2, so the
    min-curvature QP runs to the INNER bound, like the shortest path (the IQP loop exists to repair this)."""
    n, R, w, wv = 96, 60.0, 4.0, 2.0
    rt = _circle_track(n, R, w)
    path = np.vstack((rt[:, :2], rt[0, :2]))
    cx, cy, A, nv = T.calc_splines(path)
    # counter-clockwise circle: right-pointing normals point outwards
    assert np.allclose(nv, rt[:, :2] / R, atol=1e-6)
    alpha, _ = T.opt_min_curv(rt, nv, A, 0.12, wv)
    assert np.ptp(alpha) < 1e-7 and abs(alpha[0] + (w - wv / 2)) < 1e-7
    a_sp = T.opt_shortest_path(rt, nv, wv)
    assert np.ptp(a_sp) < 1e-9 and abs(a_sp[0] + (w - wv / 2)) < 1e-9
    rl = T.create_raceline(rt[:, :2], nv, alpha, 2.0)
    psi, kappa = T.calc_head_curv_an(rl[2], rl[3], rl[4], rl[5])
    assert np.allclose(kappa, 1.0 / (R + alpha[0]), rtol=1e-3)   # cubic-spline circle
    assert np.all(psi >= -np.pi) and np.all(psi < np.pi)


def test_invariances():
    rt = synth.make_track(21, 120)
    path = np.vstack((rt[:, :2], rt[0, :2]))
    _, _, A, nv = T.calc_splines(path)
    a0, _ = T.opt_min_curv(rt, nv, A, 0.12, 2.0)
    # rotation + translation leave alpha unchanged
    c, s = np.cos(0.7), np.sin(0.7)
    Rm = np.array([[c, -s], [s, c]])
    rt2 = rt.copy()
    rt2[:, :2] = rt[:, :2] @ Rm.T + np.array([13.0, -4.0])
    p2 = np.vstack((rt2[:, :2], rt2[0, :2]))
    _, _, A2, nv2 = T.calc_splines(p2)
    a2, _ = T.opt_min_curv(rt2, nv2, A2, 0.12, 2.0)
    assert rel_max(a2, a0) < 1e-6
    # reversing the direction of travel swaps left/right: alpha changes sign
    idx = np.r_[0, np.arange(rt.shape[0] - 1, 0, -1)]
    rt3 = rt[idx][:, [0, 1, 3, 2]]
    p3 = np.vstack((rt3[:, :2], rt3[0, :2]))
    _, _, A3, nv3 = T.calc_splines(p3)
    a3, _ = T.opt_min_curv(rt3, nv3, A3, 0.12, 2.0)
    assert rel_max(-a3[idx], a0) < 1e-4    # same curve, different spline knots' orientation -> close, not identical


def test_too_narrow_raises():
    rt = synth.make_track(2, 100)
    path = np.vstack((rt[:, :2], rt[0, :2]))
    _, _, A, nv = T.calc_splines(path)
    with pytest.raises(RuntimeError, match="Problem not solvable"):
        T.opt_min_curv(rt, nv, A, 0.12, 20.0)


def test_iqp_history_matches_golden(golden):
    g = golden("synth128")
    rt, nv = g["reftrack"], g["normvec"]
    path = np.vstack((rt[:, :2], rt[0, :2]))
    _, _, A, _ = T.calc_splines(path)
    hist = []
    a, rt_new, nv_new = T.iqp_handler(rt, nv, A, float(g["kappa_bound"]), float(g["w_veh"]), False, False, 3.0, 3, 0.01,
                                      history=hist)
    assert [h["n"] for h in hist] == list(g["iqp_n"])
    assert rel_max(a, g["iqp_alpha"]) < 1e-6
    assert rt_new.shape == g["iqp_reftrack"].shape and np.abs(rt_new - g["iqp_reftrack"]).max() < 1e-6
    assert np.all(rt[:, 2:] == g["reftrack"][:, 2:])       # the oracle does not mutate the caller's widths


# ---- the banded CPU port of the path (oracle/banded_cpu.c: bench.py's fair CPU baseline) against the dense oracle ----
@pytest.mark.parametrize("name", ["berlin", "handling", "synth200", "synth500", "synth1000"])
def test_banded_cpu_port_matches_the_dense_oracle(golden, name):
    from oracle import banded_cpu as BC
    g = golden(name)
    alpha, iters = BC.opt_min_curv_banded(g["reftrack"], g["normvec"], float(g["w_veh"]), f_scale=T.F_SCALE)
    ref = g["alpha_mincurv_boxonly"]              # box-only QP: the dense oracle's Goldfarb-Idnani solve without curvature rows
    assert 5 <= iters <= 30
    assert np.abs(alpha - ref).max() <= 1e-7 * np.abs(ref).max()


def test_banded_cpu_port_reports_a_track_that_is_too_narrow(golden):
    from oracle import banded_cpu as BC
    g = golden("synth200")
    with pytest.raises(RuntimeError, match="Problem not solvable"):
        BC.opt_min_curv_banded(g["reftrack"], g["normvec"], 20.0)
