"""GPU parity tests (-m gpu) of the velocity-profile stage through the C-ABI:
mc_vel_profile_batch / mc_calc_ax_t_profile_batch vs the committed golden vectors (tests/golden/velprofile.npz,
made by tools/make_golden.py from oracle/tph_velprofile.py -- PARITY UNPINNED) and vs the oracle on seeded inputs.

Tolerances: the kernel evaluates tph's statements in tph's order with round-to-nearest products and sums (no FMA
contraction), so vx / ax / t agree with the numpy oracle to rounding: 1e-11 relative on vx and t, 1e-9 on ax (a
difference of squares), with dyn_model_exp != 1 (device pow vs libm pow) and the moving-average filter included."""
import numpy as np
import pytest
import torch

import global_racetrajectory_optimization_b200 as tph
from global_racetrajectory_optimization_b200 import batch as B_
from oracle import tph_dense as T
from oracle import tph_velprofile as VP

pytestmark = pytest.mark.gpu

NAMES = ["berlin", "handling", "modena", "synth333", "synth1000"]
VEH = dict(drag_coeff=0.75, m_veh=1200.0)


def _stack(rows, n_max, fill=0.0):
    out = np.full((len(rows), n_max), fill)
    for i, r in enumerate(rows):
        out[i, :r.size] = r
    return out


def _inputs(golden, names):
    gs = [golden(n) for n in names]
    npts = np.array([g["rl_kappa"].size for g in gs], dtype=np.int32)
    n_max = int(npts.max()) + 7
    dev = torch.device("cuda")
    kappa = torch.tensor(_stack([g["rl_kappa"] for g in gs], n_max), device=dev)
    el = torch.tensor(_stack([g["rl_el_lengths"] for g in gs], n_max, fill=1.0), device=dev)
    return gs, npts, kappa, el, torch.tensor(npts, device=dev)


def test_profiles_match_golden_on_a_ragged_batch(golden):
    v = golden("velprofile")
    gs, npts, kappa, el, nd = _inputs(golden, NAMES)
    res = B_.vel_profile_batch(kappa, el, v["ggv"], v["ax_max_machines"], float(v["v_max"]), float(v["dragcoeff"]),
                               float(v["mass"]), dyn_model_exp=float(v["dyn_model_exp"]), n_pts=nd)
    assert int((res["status"] != 0).sum()) == 0
    for i, name in enumerate(NAMES):
        n = int(npts[i])
        vx, ax, t = (res[k][i, 0].cpu().numpy() for k in ("vx", "ax", "t"))
        gvx, gax, gt = v[name + "_vx"], v[name + "_ax"], v[name + "_t"]
        assert np.abs(vx[:n] - gvx).max() <= 1e-11 * gvx.max()
        assert np.abs(ax[:n] - gax).max() <= 1e-9 * np.abs(gax).max()
        assert np.abs(t[:n + 1] - gt).max() <= 1e-11 * gt[-1]
        assert float(res["laptime"][i, 0]) == t[n]
        assert np.all(vx[n:] == 0.0) and np.all(t[n + 1:] == 0.0)          # padding untouched


def test_lap_time_matrix_matches_golden(golden):
    v = golden("velprofile")
    names = ["handling", "synth333", "berlin"]
    gs, npts, kappa, el, nd = _inputs(golden, names)
    ltm = B_.lap_time_matrix_batch(kappa, el, v["ggv"], v["ax_max_machines"], v["ltm_scales"], v["ltm_top_speeds"],
                                   float(v["dragcoeff"]), float(v["mass"]), n_pts=nd).cpu().numpy()
    assert ltm.shape == (3, v["ltm_top_speeds"].size, v["ltm_scales"].size)
    for i, name in enumerate(names):
        want = v[name + "_ltm"]
        assert np.abs(ltm[i] - want).max() <= 1e-11 * want.max()
    # physics: a weaker car (smaller ggv scale / lower top speed) is never faster
    assert np.all(np.diff(ltm, axis=2) < 0.0) and np.all(np.diff(ltm, axis=1) <= 1e-9)


def test_variants_mu_filter_and_exponent_vs_oracle(golden):
    v = golden("velprofile")
    names = ["handling", "synth333"]
    gs, npts, kappa, el, nd = _inputs(golden, names)
    rng = np.random.default_rng(7)
    mu_rows = [0.8 + 0.3 * rng.random(int(n)) for n in npts]
    mu = torch.tensor(_stack(mu_rows, kappa.shape[1], fill=1.0), device=kappa.device)
    scales, speeds = [1.0, 0.7, 0.45], [60.0, 33.0, 41.5]
    res = B_.vel_profile_batch(kappa, el, v["ggv"], v["ax_max_machines"], speeds, dyn_model_exp=1.6, filt_window=5, mu=mu,
                               n_pts=nd, ggv_scales=scales, **VEH)
    assert int((res["status"] != 0).sum()) == 0
    for i, g in enumerate(gs):
        n = int(npts[i])
        for j in range(3):
            ggv = v["ggv"].copy()
            ggv[:, 1:] *= scales[j]
            k, e = g["rl_kappa"], g["rl_el_lengths"]
            vx = VP.calc_vel_profile(ggv=ggv, ax_max_machines=v["ax_max_machines"], v_max=speeds[j], kappa=k, el_lengths=e,
                                     closed=True, dyn_model_exp=1.6, filt_window=5, mu=mu_rows[i], **VEH)
            ax = VP.calc_ax_profile(np.append(vx, vx[0]), e)
            t = VP.calc_t_profile(vx, e, ax_profile=ax)
            assert np.abs(res["vx"][i, j, :n].cpu().numpy() - vx).max() <= 1e-11 * vx.max()
            assert np.abs(res["ax"][i, j, :n].cpu().numpy() - ax).max() <= 1e-9 * np.abs(ax).max()
            assert abs(float(res["laptime"][i, j]) - t[-1]) <= 1e-11 * t[-1]


def test_tph_surface_and_error_behaviour(golden):
    v, g = golden("velprofile"), golden("handling")
    k, el = g["rl_kappa"], g["rl_el_lengths"]
    ggv, mach = v["ggv"], v["ax_max_machines"]
    # the call sequence of /root/reference/main_globaltraj.py:400-421
    vx = tph.calc_vel_profile.calc_vel_profile(ggv=ggv, ax_max_machines=mach, v_max=70.0, kappa=k, el_lengths=el,
                                               closed=True, filt_window=None, dyn_model_exp=1.0, drag_coeff=0.75,
                                               m_veh=1200.0)
    vx_cl = np.append(vx, vx[0])
    ax = tph.calc_ax_profile.calc_ax_profile(vx_profile=vx_cl, el_lengths=el, eq_length_output=False)
    t = tph.calc_t_profile.calc_t_profile(vx_profile=vx, ax_profile=ax, el_lengths=el)
    assert vx.shape == k.shape and ax.shape == k.shape and t.shape == (k.size + 1,)
    assert np.abs(vx - v["handling_vx"]).max() <= 1e-11 * vx.max()
    assert np.abs(ax - v["handling_ax"]).max() <= 1e-9 * np.abs(ax).max()
    assert np.abs(t - v["handling_t"]).max() <= 1e-11 * t[-1]
    # other argument forms of tph
    assert np.array_equal(tph.calc_ax_profile.calc_ax_profile(vx_cl, el, eq_length_output=True), np.append(ax, 0.0))
    t2 = tph.calc_t_profile.calc_t_profile(vx_profile=vx_cl, el_lengths=el, t_start=3.0)
    assert np.abs(t2 - (VP.calc_t_profile(vx_cl, el, t_start=3.0))).max() <= 1e-11 * t2[-1]
    # tph's exceptions
    with pytest.raises(RuntimeError, match="same length if closed"):
        tph.calc_vel_profile.calc_vel_profile(ggv=ggv, ax_max_machines=mach, v_max=70.0, kappa=k, el_lengths=el[:-1],
                                              closed=True, drag_coeff=0.75, m_veh=1200.0)
    with pytest.raises(RuntimeError, match="entire velocity range"):
        tph.calc_vel_profile.calc_vel_profile(ggv=ggv, ax_max_machines=mach, v_max=90.0, kappa=k, el_lengths=el,
                                              closed=True, drag_coeff=0.75, m_veh=1200.0)
    with pytest.raises(RuntimeError, match="must be odd"):
        tph.calc_vel_profile.calc_vel_profile(ggv=ggv, ax_max_machines=mach, v_max=70.0, kappa=k, el_lengths=el,
                                              closed=True, drag_coeff=0.75, m_veh=1200.0, filt_window=4)
    with pytest.raises(RuntimeError, match="Either ggv or loc_gg"):
        tph.calc_vel_profile.calc_vel_profile(ax_max_machines=mach, kappa=k, el_lengths=el, closed=True, drag_coeff=0.75,
                                              m_veh=1200.0)
    with pytest.raises(RuntimeError, match="1 element bigger"):
        tph.calc_ax_profile.calc_ax_profile(vx, el)
    with pytest.raises(NotImplementedError):
        tph.calc_vel_profile.calc_vel_profile(ggv=ggv, ax_max_machines=mach, v_max=70.0, kappa=np.append(k, 0.0),
                                              el_lengths=el, closed=False, v_start=10.0, drag_coeff=0.75, m_veh=1200.0)


def test_reftrack_to_lap_time_on_device_vs_oracle_chain(golden):
    """reftrack -> splines -> min-curvature QP -> raceline (kappa, el) -> velocity profile -> lap time, all batched on
    the device, against the same chain through the CPU oracle for one track; plus batch consistency."""
    v = golden("velprofile")
    g = golden("synth200")
    dev = torch.device("cuda")
    rt = g["reftrack"]
    rtd = torch.tensor(np.stack([rt, rt, rt]), device=dev)
    cx, cy, nv, h = B_.calc_splines_batch(rtd)
    res = B_.opt_min_curv_batch(rtd, nv, h, float(g["kappa_bound"]), float(g["w_veh"]))
    rl = B_.create_raceline_batch(rtd, nv, res["alpha"], 2.0)
    n_out = rl["n_out"]
    vp = B_.vel_profile_batch(rl["kappa"], rl["el_lengths_interp"], v["ggv"], v["ax_max_machines"], 70.0, n_pts=n_out, **VEH)
    assert int((vp["status"] != 0).sum()) == 0
    lap = vp["laptime"][:, 0].cpu().numpy()
    assert lap[0] == lap[1] == lap[2]                                   # identical tracks -> bit-identical lap times
    path = np.vstack((rt[:, :2], rt[0, :2]))
    _, _, A, onv = T.calc_splines(path)
    oalpha, _ = T.opt_min_curv(rt, onv, A, float(g["kappa_bound"]), float(g["w_veh"]))
    orl = T.create_raceline(rt[:, :2], onv, oalpha, 2.0)
    _, okappa = T.calc_head_curv_an(orl[2], orl[3], orl[4], orl[5])
    ovx = VP.calc_vel_profile(ggv=v["ggv"], ax_max_machines=v["ax_max_machines"], v_max=70.0, kappa=okappa,
                              el_lengths=orl[8], closed=True, **VEH)
    oax = VP.calc_ax_profile(np.append(ovx, ovx[0]), orl[8])
    ot = VP.calc_t_profile(ovx, orl[8], ax_profile=oax)
    m = int(n_out[0])
    assert m == okappa.size
    assert np.abs(vp["vx"][0, 0, :m].cpu().numpy() - ovx).max() <= 1e-6 * ovx.max()
    assert abs(lap[0] - ot[-1]) <= 1e-7 * ot[-1]
