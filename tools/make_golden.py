"""Generates tests/golden/*.npz -- inputs and CPU-oracle outputs of the hot path.

Run in the build container (needs /root/reference for the track CSVs; the tests only read the .npz):

    python tools/make_golden.py [--only berlin,synth200,...]

Each fixture holds the reftrack fed to the path and what the dense oracle (oracle/tph_dense.py +
oracle/quadprog_gi.c, PARITY UNPINNED -- see their headers) returns for
calc_splines / opt_min_curv / opt_shortest_path / create_raceline / calc_head_curv_an / iqp_handler.
The pre-processing that turns a track CSV into a reftrack (tph.spline_approximation, FITPACK based,
/root/reference/helper_funcs_glob/src/prep_track.py:39-45) is restated here with scipy only to obtain
realistic inputs; parity is defined on identical reftrack input, so it is outside the parity loop.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time

import numpy as np
from scipy import interpolate, optimize, spatial

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tph_dense as T  # noqa: E402
from global_racetrajectory_optimization_b200 import synth  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")

# /root/reference/params/racecar.ini: stepsize_opts (:13-15), reg_smooth_opts (:21-22), curvlim (:49),
# optim_opts_mincurv (:72-74), optim_opts_shortest_path (:65)
PARS = dict(stepsize_prep=1.0, stepsize_reg=3.0, stepsize_interp_after_opt=2.0, k_reg=3, s_reg=10,
            curvlim=0.12, width_opt=3.4, iqp_iters_min=3, iqp_curverror_allowed=0.01)


def import_track(name: str) -> np.ndarray:
    """4-column branch of /root/reference/helper_funcs_glob/src/import_track.py:29-45."""
    csv = np.loadtxt(os.path.join(REF, "inputs", "tracks", name + ".csv"), comments="#", delimiter=",")
    return np.column_stack((csv[:, 0:2], csv[:, 2], csv[:, 3]))


def interp_track(track: np.ndarray, stepsize: float) -> np.ndarray:
    track_cl = np.vstack((track, track[0]))
    el = np.sqrt(np.sum(np.power(np.diff(track_cl[:, :2], axis=0), 2), axis=1))
    dists = np.insert(np.cumsum(el), 0, 0.0)
    n = math.ceil(dists[-1] / stepsize) + 1
    di = np.linspace(0.0, dists[-1], n)
    out = np.zeros((n, track_cl.shape[1]))
    for c in range(track_cl.shape[1]):
        out[:, c] = np.interp(di, dists, track_cl[:, c])
    return out[:-1]


def _side_of_line(a, b, z):
    return np.sign((b[0] - a[0]) * (z[1] - a[1]) - (b[1] - a[1]) * (z[0] - a[0]))


def spline_approximation(track, k_reg=3, s_reg=10, stepsize_prep=1.0, stepsize_reg=3.0):
    """scipy restatement of tph.spline_approximation (input generation only)."""
    track_interp = interp_track(track, stepsize_prep)
    track_interp_cl = np.vstack((track_interp, track_interp[0]))
    track_cl = np.vstack((track, track[0]))
    n_cl = track_cl.shape[0]
    el = np.sqrt(np.sum(np.power(np.diff(track_cl[:, :2], axis=0), 2), axis=1))
    dists = np.insert(np.cumsum(el), 0, 0.0)
    tck, _ = interpolate.splprep([track_interp_cl[:, 0], track_interp_cl[:, 1]], k=k_reg, s=s_reg, per=1)[:2]
    n_len = math.ceil(dists[-1]) * 4
    tmp = np.array(interpolate.splev(np.linspace(0.0, 1.0, n_len), tck)).T
    length = np.sum(np.sqrt(np.sum(np.power(np.diff(tmp, axis=0), 2), axis=1)))
    n_reg_cl = math.ceil(length / stepsize_reg) + 1
    path = np.array(interpolate.splev(np.linspace(0.0, 1.0, n_reg_cl), tck)).T[:-1]
    # closest points on the spline (dense sampling + local refinement instead of per-point Nelder-Mead)
    tt = np.linspace(0.0, 1.0, 20 * n_len)
    dense = np.array(interpolate.splev(tt, tck)).T
    tree = spatial.cKDTree(dense)
    _, idx = tree.query(track_cl[:, :2])
    t_close = np.zeros(n_cl)
    d_close = np.zeros(n_cl)
    p_close = np.zeros((n_cl, 2))
    for i in range(n_cl):
        fun = lambda t: float(np.sum((np.array(interpolate.splev(t, tck)).ravel() - track_cl[i, :2]) ** 2))
        lo, hi = tt[max(idx[i] - 2, 0)], tt[min(idx[i] + 2, tt.size - 1)]
        t_close[i] = optimize.minimize_scalar(fun, bounds=(lo, hi), method="bounded", options=dict(xatol=1e-10)).x
        p_close[i] = np.array(interpolate.splev(t_close[i], tck)).ravel()
        d_close[i] = math.sqrt(fun(t_close[i]))
    t_close[0], t_close[-1] = 0.0, 1.0
    sides = np.array([_side_of_line(track_cl[i, :2], track_cl[i + 1, :2], p_close[i]) for i in range(n_cl - 1)])
    sides_cl = np.hstack((sides, sides[0]))
    w_r = track_cl[:, 2] + sides_cl * d_close
    w_l = track_cl[:, 3] - sides_cl * d_close
    order = np.argsort(t_close, kind="stable")
    tq = np.linspace(0.0, 1.0, n_reg_cl)
    w_r_s = np.interp(tq, t_close[order], w_r[order])
    w_l_s = np.interp(tq, t_close[order], w_l[order])
    return np.column_stack((path, w_r_s[:-1], w_l_s[:-1]))


def berlin_n(n: int) -> np.ndarray:
    """The smoothed Berlin reftrack re-sampled (linear in arc length, like interp_track) to n equidistant points."""
    rt = spline_approximation(import_track("berlin_2018"), PARS["k_reg"], PARS["s_reg"], PARS["stepsize_prep"],
                              PARS["stepsize_reg"])
    cl = np.vstack((rt, rt[0]))
    s = np.insert(np.cumsum(np.sqrt(np.sum(np.diff(cl[:, :2], axis=0) ** 2, axis=1))), 0, 0.0)
    sq = np.linspace(0.0, s[-1], n, endpoint=False)
    return np.column_stack([np.interp(sq, s, cl[:, c]) for c in range(4)])


def oracle_case(name: str, reftrack: np.ndarray, w_veh: float, kappa_bound: float, with_iqp: bool) -> dict:
    t0 = time.time()
    path_cl = np.vstack((reftrack[:, :2], reftrack[0, :2]))
    cx, cy, A, nv = T.calc_splines(path=path_cl)
    n = reftrack.shape[0]
    scaling = np.array([-A[4 * i + 2, 4 * i + 5] for i in range(n - 1)] + [A[4 * n - 2, 1]])
    el = np.sqrt(np.sum(np.diff(path_cl, axis=0) ** 2, axis=1))
    out = dict(reftrack=reftrack, w_veh=w_veh, kappa_bound=kappa_bound, coeffs_x=cx, coeffs_y=cy, normvec=nv,
               scaling=scaling, el_lengths=el)
    qp = T.assemble_min_curv(reftrack, nv, A, kappa_bound, w_veh)
    from oracle.quadprog_gi import solve_qp
    # box-only QP (what the GPU phase 1 solves) and the full QP with curvature rows (what tph solves)
    G, h = qp["G"], qp["h"]
    alpha_full = solve_qp(qp["H"], -qp["f"], -G.T, -h, 0)[0]
    alpha_box = solve_qp(qp["H"], -qp["f"], -G[:2 * n].T, -h[:2 * n], 0)[0]
    klin = qp["k_kappa_ref"] + qp["E_kappa"] @ alpha_full
    out.update(alpha_mincurv=alpha_full, alpha_mincurv_boxonly=alpha_box,
               curv_error_max=T.curv_error_max_from(qp, alpha_full), kappa_lin=klin, k_kappa_ref=qp["k_kappa_ref"],
               kappa_rows_active=bool(np.abs(alpha_full - alpha_box).max() > 1e-9 * max(1.0, np.abs(alpha_full).max())),
               x_prime=qp["x_prime"], y_prime=qp["y_prime"], f=qp["f"],
               H_band=np.stack([np.array([qp["H"][i, (i + k) % n] for k in range(33)]) for i in range(n)]))
    out["alpha_shpath"] = T.opt_shortest_path(reftrack, nv, w_veh)
    rl = T.create_raceline(reftrack[:, :2], nv, alpha_full, PARS["stepsize_interp_after_opt"])
    psi, kappa = T.calc_head_curv_an(rl[2], rl[3], rl[4], rl[5])
    out.update(rl_raceline_interp=rl[0], rl_coeffs_x=rl[2], rl_coeffs_y=rl[3], rl_spline_inds=rl[4], rl_t_values=rl[5],
               rl_s=rl[6], rl_spline_lengths=rl[7], rl_el_lengths=rl[8], rl_psi=psi, rl_kappa=kappa)
    if with_iqp:
        hist = []
        a_i, rt_i, nv_i = T.iqp_handler(reftrack, nv, A, kappa_bound, w_veh, False, False, PARS["stepsize_reg"],
                                        PARS["iqp_iters_min"], PARS["iqp_curverror_allowed"], history=hist)
        out.update(iqp_alpha=a_i, iqp_reftrack=rt_i, iqp_normvec=nv_i,
                   iqp_n=np.array([h_["n"] for h_ in hist]), iqp_curv_error=np.array([h_["curv_error_max"] for h_ in hist]))
    print(f"  {name}: N={n} max|alpha|={np.abs(alpha_full).max():.3f} kappa-rows-active={out['kappa_rows_active']} "
          f"max|klin|={np.abs(klin).max():.4f} curv_err={out['curv_error_max']:.4f} ({time.time() - t0:.1f}s)", flush=True)
    return out


CASES = {
    # name: (builder, w_veh, kappa_bound, with_iqp)
    "berlin": (lambda: spline_approximation(import_track("berlin_2018"), PARS["k_reg"], PARS["s_reg"],
                                            PARS["stepsize_prep"], PARS["stepsize_reg"]), PARS["width_opt"], PARS["curvlim"], True),
    "handling": (lambda: spline_approximation(import_track("handling_track"), PARS["k_reg"], PARS["s_reg"],
                                              PARS["stepsize_prep"], PARS["stepsize_reg"]), 2.0, PARS["curvlim"], True),
    "modena": (lambda: spline_approximation(import_track("modena_2019"), PARS["k_reg"], PARS["s_reg"],
                                            PARS["stepsize_prep"], PARS["stepsize_reg"]), 2.0, PARS["curvlim"], False),
    "synth128": (lambda: synth.make_track(11, 128), 2.0, 0.12, True),
    "synth200": (lambda: synth.make_track(12, 200), 2.0, 0.12, False),
    "synth333": (lambda: synth.make_track(13, 333), 2.0, 0.12, True),
    "synth500": (lambda: synth.make_track(14, 500), 2.0, 0.12, False),
    "synth500_narrow": (lambda: synth.make_track(15, 500), 5.2, 0.12, False),
    "synth1000": (lambda: synth.make_track(16, 1000), 2.0, 0.12, False),
    # BASELINE.json configs[3] (C4) size: N = 2000 (dense oracle: ~minutes, 8000 x 8000 systems)
    "synth2000": (lambda: synth.make_track(17, 2000), 2.0, 0.12, False),
    # BASELINE.json configs[1] (C2): Berlin centre line re-sampled to N = 500 points, smooth width jitter,
    # vehicle widths from the veh_width grid 1.6 ... 3.4 m
    "berlin500_jitter_a": (lambda: synth.jitter_widths(berlin_n(500), 41), 1.6, PARS["curvlim"], False),
    "berlin500_jitter_b": (lambda: synth.jitter_widths(berlin_n(500), 42), 3.4, PARS["curvlim"], False),
    # curvature rows |k_ref + E alpha| <= kappa_bound active at the optimum (tight kappa_bound)
    "synth160_kappa": (lambda: synth.make_track(3, 160), 2.0, 0.02, False),
    "synth333_kappa": (lambda: synth.make_track(13, 333), 2.0, 0.03, False),
}


def velprofile_golden():
    """tests/golden/velprofile.npz: the reference's ggv / ax_max_machines tables and vehicle parameters
    (/root/reference/inputs/veh_dyn_info/*.csv, /root/reference/params/racecar.ini:44-57) with what
    oracle/tph_velprofile.py returns on the raceline kappa / el_lengths of the committed fixtures: the velocity,
    acceleration and time profiles of the stock run, and a small lap-time matrix
    (/root/reference/main_globaltraj.py:442-496 with a coarser grid)."""
    from oracle import tph_velprofile as VP
    ggv, axm = VP.import_veh_dyn_info(os.path.join(REF, "inputs", "veh_dyn_info", "ggv.csv"),
                                      os.path.join(REF, "inputs", "veh_dyn_info", "ax_max_machines.csv"))
    veh = dict(v_max=70.0, mass=1200.0, dragcoeff=0.75, dyn_model_exp=1.0)
    out = dict(ggv=ggv, ax_max_machines=axm, **veh)
    scales = np.linspace(0.3, 1.0, 4)
    speeds = np.linspace(100.0, 150.0, 3) / 3.6
    out.update(ltm_scales=scales, ltm_top_speeds=speeds)
    for name in ("berlin", "handling", "modena", "synth333", "synth1000"):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        k, el = g["rl_kappa"], g["rl_el_lengths"]
        vx = VP.calc_vel_profile(ggv=ggv, ax_max_machines=axm, v_max=veh["v_max"], kappa=k, el_lengths=el, closed=True,
                                 filt_window=None, dyn_model_exp=veh["dyn_model_exp"], drag_coeff=veh["dragcoeff"],
                                 m_veh=veh["mass"])
        ax = VP.calc_ax_profile(vx_profile=np.append(vx, vx[0]), el_lengths=el, eq_length_output=False)
        t = VP.calc_t_profile(vx_profile=vx, ax_profile=ax, el_lengths=el)
        out.update({name + "_vx": vx, name + "_ax": ax, name + "_t": t})
        out[name + "_ltm"] = VP.lap_time_matrix(ggv, axm, k, el, scales, speeds, veh["dyn_model_exp"], veh["dragcoeff"],
                                                veh["mass"])
        print(f"  velprofile {name}: lap time {t[-1]:.3f} s, vx {vx.min():.2f}..{vx.max():.2f} m/s", flush=True)
    np.savez_compressed(os.path.join(GOLD, "velprofile.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    if args.only == "velprofile":
        velprofile_golden()
        return
    names = [s for s in args.only.split(",") if s] or list(CASES)
    for name in names:
        build, w_veh, kb, with_iqp = CASES[name]
        rt = build()
        try:
            out = oracle_case(name, rt, w_veh, kb, with_iqp)
        except RuntimeError as e:
            print(f"  {name}: oracle raised {e!r}")
            continue
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


if __name__ == "__main__":
    main()
