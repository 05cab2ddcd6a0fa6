"""CPU ORACLE (test infrastructure, NOT product code) -- restatement of the velocity-profile stage of
``trajectory_planning_helpers==0.76`` that follows the minimum-curvature hot path:

* calc_vel_profile -- /root/reference/main_globaltraj.py:400-410 and the lap-time matrix sweep :469-479
* calc_ax_profile  -- /root/reference/main_globaltraj.py:413-416, :482-485
* calc_t_profile   -- /root/reference/main_globaltraj.py:419-421, :488-490
* import_veh_dyn_info -- /root/reference/main_globaltraj.py:211-213 (ggv.csv / ax_max_machines.csv readers)
* conv_filt        -- inside calc_vel_profile (``filt_window``; /root/reference/params/racecar.ini:56-57)

PARITY UNPINNED, like oracle/tph_dense.py: the package is not vendored under /root/reference nor
installable offline and the reference ships no tests or golden vectors.  The restatement follows the
published algorithm of tph 0.76 (forward/backward solver over the doubled lap with the ggv diagram,
machine limit and drag) statement by statement -- scalar ``math`` calls where tph uses them, numpy
where tph uses numpy -- so that the arithmetic (operation order, no fused multiply-add) is the one
the CUDA kernel mirrors with explicit round-to-nearest intrinsics.

One detail of the closed-track solver cannot be confirmed offline and is isolated as the constant
``DECEL_LAP_SLICE_UPPER``: after the backward (deceleration) pass over the doubled lap tph keeps
``vx_profile_double[no_points:]`` (as recalled), i.e. the half of the doubled array the backward
pass visits FIRST.  csrc/vel_profile_core.cuh carries the same constant (VP_DECEL_LAP_SLICE_UPPER).

Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
from __future__ import annotations

import math

import numpy as np

DECEL_LAP_SLICE_UPPER = True


# ----------------------------------------------------------------------------------------------
def import_veh_dyn_info(ggv_import_path=None, ax_max_machines_import_path=None):
    """tph.import_veh_dyn_info: (ggv [k,3] = v, ax_max, ay_max ; ax_max_machines [m,2] = v, ax_max_machines)."""
    ggv = None
    if ggv_import_path is not None:
        with open(ggv_import_path, "rb") as fh:
            ggv = np.loadtxt(fh, comments="#", delimiter=",")
        if ggv.ndim == 1:
            ggv = np.expand_dims(ggv, 0)
        if ggv.shape[1] != 3:
            raise RuntimeError("ggv diagram must consist of the three columns [vx, ax_max, ay_max]!")
        invalid_1 = ggv[:, 0] < 0.0
        invalid_2 = ggv[:, 1:] > 50.0
        invalid_3 = ggv[:, 1] < 0.0
        invalid_4 = ggv[:, 2] < 0.0
        if np.any(invalid_1) or np.any(invalid_2) or np.any(invalid_3) or np.any(invalid_4):
            raise RuntimeError("ggv seems unreasonable!")
    ax_max_machines = None
    if ax_max_machines_import_path is not None:
        with open(ax_max_machines_import_path, "rb") as fh:
            ax_max_machines = np.loadtxt(fh, comments="#", delimiter=",")
        if ax_max_machines.ndim == 1:
            ax_max_machines = np.expand_dims(ax_max_machines, 0)
        if ax_max_machines.shape[1] != 2:
            raise RuntimeError("ax_max_machines must consist of the two columns [vx, ax_max_machines]!")
        invalid_1 = ax_max_machines[:, 0] < 0.0
        invalid_2 = ax_max_machines[:, 1] > 20.0
        invalid_3 = ax_max_machines[:, 1] < 0.0
        if np.any(invalid_1) or np.any(invalid_2) or np.any(invalid_3):
            raise RuntimeError("ax_max_machines seems unreasonable!")
    return ggv, ax_max_machines


def conv_filt(signal, filt_window, closed):
    """tph.conv_filt: moving average of odd width; cyclic on closed signals."""
    if not filt_window % 2 == 1:
        raise RuntimeError("Window width of moving average filter must be odd!")
    w_window_half = int((filt_window - 1) / 2)
    if w_window_half == 0:
        return np.copy(signal)
    if closed:
        signal_tmp = np.concatenate((signal[-w_window_half:], signal, signal[:w_window_half]), axis=0)
        signal_filt = np.convolve(signal_tmp, np.ones(filt_window) / float(filt_window),
                                  mode="same")[w_window_half:-w_window_half]
    else:
        signal_filt = np.copy(signal)
        signal_filt[w_window_half:-w_window_half] = np.convolve(
            signal, np.ones(filt_window) / float(filt_window), mode="same")[w_window_half:-w_window_half]
    return signal_filt


# ----------------------------------------------------------------------------------------------
def calc_ax_poss(vx_start, radius, ggv, mu, dyn_model_exp, drag_coeff, m_veh, ax_max_machines=None,
                 mode="accel_forw"):
    """Longitudinal acceleration the vehicle can use at one point (tyres via the ggv diagram and the
    generalised friction ellipse, machine limit in forward acceleration, drag)."""
    if mode not in ("accel_forw", "decel_forw", "decel_backw"):
        raise RuntimeError("Unknown operation mode for calc_ax_poss!")
    if mode == "accel_forw" and ax_max_machines is None:
        raise RuntimeError("ax_max_machines is required if operation mode is accel_forw!")
    if ggv.ndim != 2 or ggv.shape[1] != 3:
        raise RuntimeError("ggv must have two dimensions and three columns [vx, ax_max, ay_max]!")
    # tyre potential
    ax_max_tires = mu * np.interp(vx_start, ggv[:, 0], ggv[:, 1])
    ay_max_tires = mu * np.interp(vx_start, ggv[:, 0], ggv[:, 2])
    ay_used = math.pow(vx_start, 2) / radius
    radicand = 1.0 - math.pow(ay_used / ay_max_tires, dyn_model_exp)
    if radicand > 0.0:
        ax_avail_tires = ax_max_tires * math.pow(radicand, 1.0 / dyn_model_exp)
    else:
        ax_avail_tires = 0.0
    # machine limit
    if mode == "accel_forw":
        ax_max_machines_tmp = np.interp(vx_start, ax_max_machines[:, 0], ax_max_machines[:, 1])
        ax_avail_vehicle = min(ax_avail_tires, ax_max_machines_tmp)
    else:
        ax_avail_vehicle = ax_avail_tires
    # drag: reduces the possible acceleration going forward, helps the deceleration computed backwards
    ax_drag = -math.pow(vx_start, 2) * drag_coeff / m_veh
    if mode in ("accel_forw", "decel_forw"):
        ax_final = ax_avail_vehicle + ax_drag
    else:
        ax_final = ax_avail_vehicle - ax_drag
    return ax_final


def _solver_fb_acc_profile(ggv, ax_max_machines, v_max, radii, el_lengths, mu, vx_profile, dyn_model_exp,
                           drag_coeff, m_veh, backwards=False):
    no_points = vx_profile.size
    if backwards:
        radii_mod = np.flipud(radii)
        el_lengths_mod = np.flipud(el_lengths)
        mu_mod = np.flipud(mu)
        vx_profile = np.flipud(vx_profile)
        mode = "decel_backw"
    else:
        radii_mod = radii
        el_lengths_mod = el_lengths
        mu_mod = mu
        mode = "accel_forw"
    vx_profile = np.copy(vx_profile)

    # first point of every phase in which the profile allows positive longitudinal acceleration
    vx_diffs = np.diff(vx_profile)
    acc_inds = np.where(vx_diffs > 0.0)[0]
    if acc_inds.size != 0:
        acc_inds_diffs = np.diff(acc_inds)
        acc_inds_diffs = np.insert(acc_inds_diffs, 0, 2)
        acc_inds_rel = acc_inds[acc_inds_diffs > 1]
    else:
        acc_inds_rel = []
    acc_inds_rel = list(acc_inds_rel)

    while acc_inds_rel:
        i = acc_inds_rel.pop(0)
        while i < no_points - 1:
            ax_possible_cur = calc_ax_poss(vx_start=vx_profile[i], radius=radii_mod[i], ggv=ggv, mu=mu_mod[i],
                                           dyn_model_exp=dyn_model_exp, drag_coeff=drag_coeff, m_veh=m_veh,
                                           ax_max_machines=ax_max_machines, mode=mode)
            vx_possible_next = math.sqrt(math.pow(vx_profile[i], 2) + 2 * ax_possible_cur * el_lengths_mod[i])
            if backwards:
                # the acceleration found at point i need not be feasible at i + 1 (different speed, radius, mu)
                for _ in range(1):
                    ax_possible_next = calc_ax_poss(vx_start=vx_possible_next, radius=radii_mod[i + 1], ggv=ggv,
                                                    mu=mu_mod[i + 1], dyn_model_exp=dyn_model_exp,
                                                    drag_coeff=drag_coeff, m_veh=m_veh,
                                                    ax_max_machines=ax_max_machines, mode=mode)
                    vx_tmp = math.sqrt(math.pow(vx_profile[i], 2) + 2 * ax_possible_next * el_lengths_mod[i])
                    if vx_tmp < vx_possible_next:
                        vx_possible_next = vx_tmp
                    else:
                        break
            if vx_possible_next < vx_profile[i + 1]:
                vx_profile[i + 1] = vx_possible_next
            i += 1
            if vx_possible_next > v_max or (acc_inds_rel and i >= acc_inds_rel[0]):
                break

    if backwards:
        vx_profile = np.flipud(vx_profile)
    return vx_profile


def _solver_fb_closed(ggv, ax_max_machines, v_max, radii, el_lengths, mu, dyn_model_exp, drag_coeff, m_veh):
    no_points = radii.size
    mu_mean = np.mean(mu)
    ay_max_global = mu_mean * np.amin(ggv[:, 2])
    with np.errstate(invalid="ignore", divide="ignore"):
        vx_profile = np.sqrt(ay_max_global * radii)
        converged = False
        for _ in range(100):
            vx_profile_prev_iteration = vx_profile
            ay_max_curr = mu * np.interp(vx_profile, ggv[:, 0], ggv[:, 2])
            vx_profile = np.sqrt(np.multiply(ay_max_curr, radii))
            if np.max(np.abs(vx_profile / vx_profile_prev_iteration - 1.0)) < 0.005:
                converged = True
                break
    del converged        # tph only prints a warning
    vx_profile[vx_profile > v_max] = v_max

    # the track is closed: run two laps so that the lap transition sees the right speeds
    vx_profile_double = np.concatenate((vx_profile, vx_profile), axis=0)
    radii_double = np.concatenate((radii, radii), axis=0)
    el_lengths_double = np.concatenate((el_lengths, el_lengths), axis=0)
    mu_double = np.concatenate((mu, mu), axis=0)

    vx_profile_double = _solver_fb_acc_profile(ggv, ax_max_machines, v_max, radii_double, el_lengths_double,
                                               mu_double, vx_profile_double, dyn_model_exp, drag_coeff, m_veh,
                                               backwards=False)
    # second lap of the acceleration profile
    vx_profile_double = np.concatenate((vx_profile_double[no_points:], vx_profile_double[no_points:]), axis=0)
    vx_profile_double = _solver_fb_acc_profile(ggv, ax_max_machines, v_max, radii_double, el_lengths_double,
                                               mu_double, vx_profile_double, dyn_model_exp, drag_coeff, m_veh,
                                               backwards=True)
    if DECEL_LAP_SLICE_UPPER:
        return vx_profile_double[no_points:]
    return vx_profile_double[:no_points]


def calc_vel_profile(ax_max_machines, kappa, el_lengths, closed, drag_coeff, m_veh, ggv=None, loc_gg=None,
                     v_max=None, dyn_model_exp=1.0, mu=None, v_start=None, v_end=None, filt_window=None):
    """tph.calc_vel_profile.calc_vel_profile (ggv branch, closed tracks)."""
    if (ggv is not None or mu is not None) and loc_gg is not None:
        raise RuntimeError("Either ggv and optionally mu OR loc_gg must be supplied, not both (or all) of them!")
    if ggv is None and loc_gg is None:
        raise RuntimeError("Either ggv or loc_gg must be supplied!")
    if loc_gg is not None:
        raise NotImplementedError("loc_gg is not on the reference's call path")
    kappa = np.asarray(kappa, dtype=float)
    el_lengths = np.asarray(el_lengths, dtype=float)
    if mu is not None and kappa.size != mu.size:
        raise RuntimeError("kappa and mu must have the same length!")
    if closed and kappa.size != el_lengths.size:
        raise RuntimeError("kappa and el_lengths must have the same length if closed!")
    if not closed and kappa.size != el_lengths.size + 1:
        raise RuntimeError("kappa must have the length of el_lengths + 1 if unclosed!")
    if not closed and v_start is None:
        raise RuntimeError("v_start must be provided for the unclosed case!")
    if not closed:
        raise NotImplementedError("open tracks are not on the reference's call path")
    if ax_max_machines.shape[1] != 2:
        raise RuntimeError("ax_max_machines must consist of the two columns [vx, ax_max_machines]!")
    if ggv.shape[1] != 3:
        raise RuntimeError("ggv diagram must consist of the three columns [vx, ax_max, ay_max]!")
    if v_max is None:
        v_max = min(ggv[-1, 0], ax_max_machines[-1, 0])
    if ax_max_machines[-1, 0] < v_max:
        raise RuntimeError("ax_max_machines has to cover the entire velocity range of the car (i.e. >= v_max)!")
    if ggv[-1, 0] < v_max:
        raise RuntimeError("ggv has to cover the entire velocity range of the car (i.e. >= v_max)!")
    if mu is None:
        mu = np.ones(kappa.size)
    else:
        mu = np.asarray(mu, dtype=float)
    # curvature has a sign in our convention: radii are absolute values
    radii = np.abs(np.divide(1.0, kappa, out=np.full(kappa.size, np.inf), where=kappa != 0.0))
    vx_profile = _solver_fb_closed(ggv, ax_max_machines, v_max, radii, el_lengths, mu, dyn_model_exp, drag_coeff,
                                   m_veh)
    if filt_window is not None:
        vx_profile = conv_filt(vx_profile, filt_window, closed)
    return vx_profile


def calc_ax_profile(vx_profile, el_lengths, eq_length_output=False):
    """tph.calc_ax_profile: ax_i = (v_{i+1}^2 - v_i^2) / (2 el_i)."""
    if vx_profile.size != el_lengths.size + 1:
        raise RuntimeError("Array size of vx_profile should be 1 element bigger than el_lengths!")
    if eq_length_output:
        ax_profile = np.zeros(vx_profile.size)
        ax_profile[:-1] = (np.power(vx_profile[1:], 2) - np.power(vx_profile[:-1], 2)) / (2 * el_lengths)
    else:
        ax_profile = (np.power(vx_profile[1:], 2) - np.power(vx_profile[:-1], 2)) / (2 * el_lengths)
    return ax_profile


def calc_t_profile(vx_profile, el_lengths, t_start=0.0, ax_profile=None):
    """tph.calc_t_profile: time stamps from el = v t + a t^2 / 2 per segment."""
    if vx_profile.size < el_lengths.size:
        raise RuntimeError("vx_profile and el_lenghts must have at least the same length!")
    if ax_profile is not None and ax_profile.size < el_lengths.size:
        raise RuntimeError("ax_profile and el_lenghts must have at least the same length!")
    if ax_profile is None:
        ax_profile = calc_ax_profile(vx_profile=vx_profile, el_lengths=el_lengths, eq_length_output=False)
    no_points = el_lengths.size
    t_steps = np.zeros(no_points)
    for i in range(no_points):
        if not math.isclose(ax_profile[i], 0.0):
            t_steps[i] = (-vx_profile[i] + math.sqrt((math.pow(vx_profile[i], 2)
                                                      + 2 * ax_profile[i] * el_lengths[i]))) / ax_profile[i]
        else:
            t_steps[i] = el_lengths[i] / vx_profile[i]
    return np.insert(np.cumsum(t_steps), 0, 0.0) + t_start


def lap_time_matrix(ggv, ax_max_machines, kappa, el_lengths, ggv_scales, top_speeds, dyn_model_exp, drag_coeff, m_veh,
                    filt_window=None):
    """The sweep of /root/reference/main_globaltraj.py:442-496 (without the file output): rows = top speeds
    [m/s], columns = ggv scales; returns the [len(top_speeds), len(ggv_scales)] matrix of lap times."""
    out = np.zeros((len(top_speeds), len(ggv_scales)))
    for i, top_speed in enumerate(top_speeds):
        for j, ggv_scale in enumerate(ggv_scales):
            ggv_mod = np.copy(ggv)
            ggv_mod[:, 1:] *= ggv_scale
            vx = calc_vel_profile(ggv=ggv_mod, ax_max_machines=ax_max_machines, v_max=top_speed, kappa=kappa,
                                  el_lengths=el_lengths, dyn_model_exp=dyn_model_exp, filt_window=filt_window,
                                  closed=True, drag_coeff=drag_coeff, m_veh=m_veh)
            vx_cl = np.append(vx, vx[0])
            ax = calc_ax_profile(vx_profile=vx_cl, el_lengths=el_lengths, eq_length_output=False)
            t = calc_t_profile(vx_profile=vx, ax_profile=ax, el_lengths=el_lengths)
            out[i, j] = t[-1]
    return out
