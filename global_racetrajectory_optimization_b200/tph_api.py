"""Single-track call surface of the hot path, mirroring trajectory_planning_helpers 0.76 as used by
/root/reference/main_globaltraj.py (keyword names, defaults, return arities, exception types):

    calc_splines       prep_track.py:48-51, main_globaltraj.py:568
    opt_min_curv       main_globaltraj.py:264-271, :344-350
    iqp_handler        main_globaltraj.py:273-284
    opt_shortest_path  main_globaltraj.py:286-290
    create_raceline    main_globaltraj.py:371-376
    calc_head_curv_an  main_globaltraj.py:383-387
    calc_vel_profile   main_globaltraj.py:400-410, :469-479   (the stage after the path)
    calc_ax_profile    main_globaltraj.py:413-416, :482-485
    calc_t_profile     main_globaltraj.py:419-421, :488-490
    import_veh_dyn_info main_globaltraj.py:211-213           (host-side CSV reader, no compute)
    check_normals_crossing prep_track.py:57-59               

numpy in / numpy out; every call runs the CUDA kernels through the C-ABI with a batch of one
(no CPU fallback: without the extension or a GPU these functions raise)."""
from __future__ import annotations

import numpy as np
import torch

from . import batch as _b
from .spline_system import SplineSystem, h_from_system


def _dev():
    _b._require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _up(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(_dev()).unsqueeze(0)


def _raise_status(st: int):
    if st == 0:
        return
    if st == 4:     # still violated after the curvature-row phase: the rows cannot be met inside the track
        raise ValueError("constraints are inconsistent, no solution")
    if st == 1:
        raise RuntimeError(_b.STATUS_TEXT[1])
    if st == 3:
        raise ValueError("matrix G is not positive definite")
    if st == 2:
        raise RuntimeError("interior-point iteration cap reached without convergence")
    raise NotImplementedError(_b.STATUS_TEXT.get(st, f"solver status {st}"))


# ------------------------------------------------------------------------------------------------
def calc_splines(path: np.ndarray, el_lengths: np.ndarray = None, psi_s: float = None, psi_e: float = None,
                 use_dist_scaling: bool = True) -> tuple:
    """tph.calc_splines.calc_splines -> (coeffs_x, coeffs_y, M, normvec_normalized).
    ``M`` is a SplineSystem (lazy stand-in of the dense 4N x 4N matrix)."""
    path = np.asarray(path, dtype=np.float64)
    closed = bool(np.all(np.isclose(path[0], path[-1])) and psi_s is None)
    if not closed and (psi_s is None or psi_e is None):
        raise RuntimeError("Headings must be provided for unclosed spline calculation!")
    if el_lengths is not None and path.shape[0] != el_lengths.size + 1:
        raise RuntimeError("el_lengths input must be one element smaller than path input!")
    if not closed:
        raise NotImplementedError("open-path splines (psi_s/psi_e) are outside the B200 hot path; "
                                  "main_globaltraj.py only ever passes closed paths")
    pts = path[:-1, :2]
    el = _up(el_lengths) if (el_lengths is not None and use_dist_scaling) else None
    cx, cy, nv, h = _b.calc_splines_batch(_up(pts), el_lengths=el, use_dist_scaling=use_dist_scaling)
    return (cx[0].cpu().numpy(), cy[0].cpu().numpy(), SplineSystem(h[0].cpu().numpy()), nv[0].cpu().numpy())


def opt_min_curv(reftrack: np.ndarray, normvectors: np.ndarray, A, kappa_bound: float, w_veh: float,
                 print_debug: bool = False, plot_debug: bool = False, closed: bool = True, psi_s: float = None,
                 psi_e: float = None, fix_s: bool = False, fix_e: bool = False) -> tuple:
    """tph.opt_min_curv.opt_min_curv -> (alpha_mincurv, curv_error_max)."""
    reftrack = np.asarray(reftrack, dtype=np.float64)
    normvectors = np.asarray(normvectors, dtype=np.float64)
    no_points = reftrack.shape[0]
    if no_points != normvectors.shape[0]:
        raise RuntimeError("Array size of reftrack should be the same as normvectors!")
    if not closed:
        raise NotImplementedError("open tracks (closed=False) are outside the B200 hot path")
    h = h_from_system(A, no_points)
    res = _b.opt_min_curv_batch(_up(reftrack), _up(normvectors), _up(h), kappa_bound, float(w_veh))
    _raise_status(int(res["status"][0].item()))
    alpha = res["alpha"][0].cpu().numpy()
    if print_debug:
        print("IPM iterations opt_min_curv: %i" % int(res["iters"][0].item()))
    return alpha, float(res["curv_error_max"][0].item())


def opt_shortest_path(reftrack: np.ndarray, normvectors: np.ndarray, w_veh: float, print_debug: bool = False) -> np.ndarray:
    """tph.opt_shortest_path.opt_shortest_path -> alpha_shpath."""
    reftrack = np.asarray(reftrack, dtype=np.float64)
    normvectors = np.asarray(normvectors, dtype=np.float64)
    if reftrack.shape[0] != normvectors.shape[0]:
        raise RuntimeError("Array size of reftrack should be the same as normvectors!")
    res = _b.opt_shortest_path_batch(_up(reftrack), _up(normvectors), float(w_veh))
    _raise_status(int(res["status"][0].item()))
    return res["alpha"][0].cpu().numpy()


def create_raceline(refline: np.ndarray, normvectors: np.ndarray, alpha: np.ndarray, stepsize_interp: float) -> tuple:
    """tph.create_raceline.create_raceline -> 9-tuple (raceline_interp, A_raceline, coeffs_x_raceline,
    coeffs_y_raceline, spline_inds_raceline_interp, t_values_raceline_interp, s_raceline_interp,
    spline_lengths_raceline, el_lengths_raceline_interp_cl)."""
    refline = np.asarray(refline, dtype=np.float64)
    n = refline.shape[0]
    out = _b.create_raceline_batch(_up(refline[:, :2]), _up(normvectors), _up(alpha), float(stepsize_interp),
                                   with_head_curv=False)
    no = int(out["n_out"][0].item())
    if no <= 0:
        raise RuntimeError("create_raceline: resampling buffer too small")
    g = lambda k, m: out[k][0, :m].cpu().numpy()
    return (g("raceline_interp", no), SplineSystem(np.ones(n)), g("coeffs_x", n), g("coeffs_y", n),
            g("spline_inds", no).astype(int), g("t_values", no), g("s_interp", no), g("spline_lengths", n),
            g("el_lengths_interp", no))


def calc_head_curv_an(coeffs_x: np.ndarray, coeffs_y: np.ndarray, ind_spls: np.ndarray, t_spls: np.ndarray,
                      calc_curv: bool = True, calc_dcurv: bool = False) -> tuple:
    """tph.calc_head_curv_an.calc_head_curv_an -> (psi, kappa[, dkappa])."""
    coeffs_x = np.asarray(coeffs_x, dtype=np.float64)
    coeffs_y = np.asarray(coeffs_y, dtype=np.float64)
    ind_spls = np.asarray(ind_spls)
    t_spls = np.asarray(t_spls, dtype=np.float64)
    if coeffs_x.shape[0] != coeffs_y.shape[0]:
        raise ValueError("Coefficient matrices must have the same length!")
    if ind_spls.size != t_spls.size:
        raise ValueError("ind_spls and t_spls must have the same length!")
    if not calc_curv and calc_dcurv:
        raise ValueError("dkappa cannot be calculated without kappa!")
    psi, kappa, dkappa = _b.calc_head_curv_batch(_up(coeffs_x), _up(coeffs_y), _up(ind_spls, torch.int32), _up(t_spls),
                                                 calc_curv=calc_curv, calc_dcurv=calc_dcurv)
    psi = psi[0].cpu().numpy()
    kap = kappa[0].cpu().numpy() if calc_curv else 0.0
    if calc_dcurv:
        return psi, kap, dkappa[0].cpu().numpy()
    return psi, kap


def iqp_handler(reftrack: np.ndarray, normvectors: np.ndarray, A, kappa_bound: float, w_veh: float,
                print_debug: bool, plot_debug: bool, stepsize_interp: float, iters_min: int = 3,
                curv_error_allowed: float = 0.01) -> tuple:
    """tph.iqp_handler.iqp_handler -> (alpha_mincurv_tmp, reftrack_tmp, normvectors_tmp) of the last
    iteration.  Unlike tph the caller's reftrack is not modified in place (tph aliases it on iteration 1)."""
    reftrack = np.asarray(reftrack, dtype=np.float64)
    normvectors = np.asarray(normvectors, dtype=np.float64)
    h = h_from_system(A, reftrack.shape[0])
    res = _b.iqp_batch(_up(reftrack), _up(normvectors), _up(h), kappa_bound, float(w_veh), float(stepsize_interp),
                       iters_min=int(iters_min), curv_error_allowed=float(curv_error_allowed))
    _raise_status(int(res["status"][0].item()))
    n = int(res["n_pts"][0].item())
    if print_debug:
        print("Minimum curvature IQP: %i iterations, curv_error_max: %.4frad/m"
              % (int(res["outer_iters"][0].item()), float(res["curv_error_max"][0].item())))
    return (res["alpha"][0, :n].cpu().numpy(), res["reftrack"][0, :n].cpu().numpy(), res["normvec"][0, :n].cpu().numpy())


# ------------------------------------------------------------------------------------------------
# velocity-profile stage
# ------------------------------------------------------------------------------------------------
def import_veh_dyn_info(ggv_import_path: str = None, ax_max_machines_import_path: str = None) -> tuple:
    """tph.import_veh_dyn_info.import_veh_dyn_info -> (ggv [k, 3], ax_max_machines [m, 2]); host-side file reader
    with tph's plausibility checks."""
    ggv = None
    if ggv_import_path is not None:
        with open(ggv_import_path, "rb") as fh:
            ggv = np.loadtxt(fh, comments="#", delimiter=",")
        if ggv.ndim == 1:
            ggv = np.expand_dims(ggv, 0)
        if ggv.shape[1] != 3:
            raise RuntimeError("ggv diagram must consist of the three columns [vx, ax_max, ay_max]!")
        if np.any(ggv[:, 0] < 0.0) or np.any(ggv[:, 1:] > 50.0) or np.any(ggv[:, 1] < 0.0) or np.any(ggv[:, 2] < 0.0):
            raise RuntimeError("ggv seems unreasonable!")
    ax_max_machines = None
    if ax_max_machines_import_path is not None:
        with open(ax_max_machines_import_path, "rb") as fh:
            ax_max_machines = np.loadtxt(fh, comments="#", delimiter=",")
        if ax_max_machines.ndim == 1:
            ax_max_machines = np.expand_dims(ax_max_machines, 0)
        if ax_max_machines.shape[1] != 2:
            raise RuntimeError("ax_max_machines must consist of the two columns [vx, ax_max_machines]!")
        if np.any(ax_max_machines[:, 0] < 0.0) or np.any(ax_max_machines[:, 1] > 20.0) or np.any(ax_max_machines[:, 1] < 0.0):
            raise RuntimeError("ax_max_machines seems unreasonable!")
    return ggv, ax_max_machines


def calc_vel_profile(ax_max_machines: np.ndarray, kappa: np.ndarray, el_lengths: np.ndarray, closed: bool,
                     drag_coeff: float, m_veh: float, ggv: np.ndarray = None, loc_gg: np.ndarray = None,
                     v_max: float = None, dyn_model_exp: float = 1.0, mu: np.ndarray = None, v_start: float = None,
                     v_end: float = None, filt_window: int = None) -> np.ndarray:
    """tph.calc_vel_profile.calc_vel_profile -> vx_profile (closed tracks, ggv branch)."""
    if (ggv is not None or mu is not None) and loc_gg is not None:
        raise RuntimeError("Either ggv and optionally mu OR loc_gg must be supplied, not both (or all) of them!")
    if ggv is None and loc_gg is None:
        raise RuntimeError("Either ggv or loc_gg must be supplied!")
    if loc_gg is not None:
        raise NotImplementedError("loc_gg is outside the B200 path; main_globaltraj.py passes a ggv diagram")
    kappa = np.asarray(kappa, dtype=np.float64)
    el_lengths = np.asarray(el_lengths, dtype=np.float64)
    if mu is not None and kappa.size != np.asarray(mu).size:
        raise RuntimeError("kappa and mu must have the same length!")
    if closed and kappa.size != el_lengths.size:
        raise RuntimeError("kappa and el_lengths must have the same length if closed!")
    if not closed and kappa.size != el_lengths.size + 1:
        raise RuntimeError("kappa must have the length of el_lengths + 1 if unclosed!")
    if not closed and v_start is None:
        raise RuntimeError("v_start must be provided for the unclosed case!")
    if not closed:
        raise NotImplementedError("open tracks (closed=False) are outside the B200 path; main_globaltraj.py passes closed=True")
    ggv = np.asarray(ggv, dtype=np.float64)
    ax_max_machines = np.asarray(ax_max_machines, dtype=np.float64)
    if ax_max_machines.ndim != 2 or ax_max_machines.shape[1] != 2:
        raise RuntimeError("ax_max_machines must consist of the two columns [vx, ax_max_machines]!")
    if ggv.ndim != 2 or ggv.shape[1] != 3:
        raise RuntimeError("ggv diagram must consist of the three columns [vx, ax_max, ay_max]!")
    if v_max is None:
        v_max = min(ggv[-1, 0], ax_max_machines[-1, 0])
    if not 1.0 <= dyn_model_exp <= 2.0:
        print("WARNING: Exponent for the vehicle dynamics model should be in the range [1.0, 2.0]!")
    res = _b.vel_profile_batch(_up(kappa), _up(el_lengths), ggv, ax_max_machines, float(v_max), float(drag_coeff),
                               float(m_veh), dyn_model_exp=float(dyn_model_exp), filt_window=filt_window,
                               mu=_up(mu) if mu is not None else None)
    return res["vx"][0, 0].cpu().numpy()


def calc_ax_profile(vx_profile: np.ndarray, el_lengths: np.ndarray, eq_length_output: bool = False) -> np.ndarray:
    """tph.calc_ax_profile.calc_ax_profile -> ax_profile."""
    vx_profile = np.asarray(vx_profile, dtype=np.float64)
    el_lengths = np.asarray(el_lengths, dtype=np.float64)
    if vx_profile.size != el_lengths.size + 1:
        raise RuntimeError("Array size of vx_profile should be 1 element bigger than el_lengths!")
    ax, _ = _b.calc_ax_t_profile_batch(_up(vx_profile), _up(el_lengths), want_t=False)
    ax = ax[0].cpu().numpy()
    if eq_length_output:
        return np.append(ax, 0.0)
    return ax


def calc_t_profile(vx_profile: np.ndarray, el_lengths: np.ndarray, t_start: float = 0.0,
                   ax_profile: np.ndarray = None) -> np.ndarray:
    """tph.calc_t_profile.calc_t_profile -> t_profile (el_lengths.size + 1 entries)."""
    vx_profile = np.asarray(vx_profile, dtype=np.float64)
    el_lengths = np.asarray(el_lengths, dtype=np.float64)
    if vx_profile.size < el_lengths.size:
        raise RuntimeError("vx_profile and el_lenghts must have at least the same length!")
    if ax_profile is not None and np.asarray(ax_profile).size < el_lengths.size:
        raise RuntimeError("ax_profile and el_lenghts must have at least the same length!")
    n = el_lengths.size
    if ax_profile is None:
        if vx_profile.size < n + 1:       # tph's calc_ax_profile call raises in this case
            raise RuntimeError("Array size of vx_profile should be 1 element bigger than el_lengths!")
        _, t = _b.calc_ax_t_profile_batch(_up(vx_profile[:n + 1]), _up(el_lengths), t_start=float(t_start))
    else:
        ax_in = np.asarray(ax_profile, dtype=np.float64)[:n]
        _, t = _b.calc_ax_t_profile_batch(_up(vx_profile[:n]), _up(el_lengths), ax_in=_up(ax_in), t_start=float(t_start))
    return t[0].cpu().numpy()


def check_normals_crossing(track: np.ndarray, normvec_normalized: np.ndarray, horizon: int = 10) -> bool:
    """tph.check_normals_crossing.check_normals_crossing -> True if normals cross inside the track."""
    track = np.asarray(track, dtype=np.float64)
    normvec_normalized = np.asarray(normvec_normalized, dtype=np.float64)
    no_points = track.shape[0]
    if horizon >= no_points:
        raise RuntimeError("Horizon of %i points is too large for a track with %i points, reduce horizon!"
                           % (horizon, no_points))
    elif horizon >= no_points / 2:
        print("WARNING: Horizon of %i points makes no sense for a track with %i points, reduce horizon!"
              % (horizon, no_points))
    return bool(_b.check_normals_crossing_batch(_up(track), _up(normvec_normalized), int(horizon))[0].item())


def spline_approximation(track: np.ndarray, k_reg: int = 3, s_reg: int = 10, stepsize_prep: float = 1.0,
                         stepsize_reg: float = 3.0, debug: bool = False) -> np.ndarray:
    """tph.spline_approximation.spline_approximation -> track_reg [n, 4] (unclosed), call site
    /root/reference/helper_funcs_glob/src/prep_track.py:39-45.  Same statements on the device, with the Reinsch smoothing
    spline (residual budget s_reg) in place of scipy's FITPACK splprep (csrc/prep_track.cu)."""
    track = np.asarray(track, dtype=np.float64)
    if track.ndim != 2 or track.shape[1] != 4:
        raise ValueError("track must be an [n, 4] array [x, y, w_tr_right, w_tr_left]")
    out, n_out, lam = _b.spline_approximation_batch(_up(track), k_reg=int(k_reg), s_reg=float(s_reg),
                                                    stepsize_prep=float(stepsize_prep), stepsize_reg=float(stepsize_reg))
    res = out[0, :int(n_out[0].item())].cpu().numpy()
    if debug:
        print("Spline approximation: smoothing parameter %.3e, %i points" % (float(lam[0].item()), res.shape[0]))
    return res
