"""Parity-pinning kit: turns "parity unpinned" into a pinned oracle the day the real packages are importable.

    pip install trajectory-planning-helpers==0.76 quadprog        # (needs an index; not possible in the build container)
    python tools/pin_against_tph.py --write                        # on any machine, CPU only
    python -m pytest tests/ -q                                     # re-runs every parity test against the REAL goldens
    python -m pytest tests/test_real_tph.py -q                     # oracle and CUDA path against the live packages

What it does (no /root/reference needed: the raw track files travel inside tests/golden/refback_import_track.npz):
 1. imports the REAL ``trajectory_planning_helpers`` and ``quadprog`` (the packages the reference calls,
    /root/reference/requirements.txt:3, /root/reference/main_globaltraj.py:264-290,371-387);
 2. prepares Berlin / Modena / handling track / rounded rectangle through the statements of the stock ``prep_track``
    (/root/reference/helper_funcs_glob/src/prep_track.py:39-51) with the real ``tph.spline_approximation``;
 3. determines the two constants that cannot be confirmed offline:
       f_scale            in {1, 2}  -- which scaling of the linear term reproduces the real ``tph.opt_min_curv``
       decel_slice_upper  in {0, 1}  -- which half of the doubled lap the real ``tph.calc_vel_profile`` keeps
 4. compares every function of oracle/ with the real package on those tracks and on the synthetic fixtures and prints
    the table;
 5. with --write: regenerates every fixture of tests/golden/ from the REAL packages into tests/golden_real/ (same keys)
    together with pin.json {f_scale, decel_slice_upper, versions, errors}.  tests/conftest.py picks that directory up:
    the oracle constants, the run-time parameters of the C-ABI (mc_mincurv_solve_batch_ex / mc_vel_profile_batch_ex via
    batch.F_SCALE / batch.VP_DECEL_SLICE_UPPER) and the golden vectors of all parity tests switch to what the real
    packages say -- no rebuild.

Exit status: 0 pinned (oracle reproduces the real packages), 1 the restatement differs beyond the two constants
(the table says where), 2 the packages are not importable.
"""
from __future__ import annotations

import argparse
import io
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

# /root/reference/params/racecar.ini: stepsize_opts (:13-15), reg_smooth_opts (:21-22), curvlim (:49), veh width (:46)
PARS = dict(stepsize_prep=1.0, stepsize_reg=3.0, stepsize_interp_after_opt=2.0, k_reg=3, s_reg=10, curvlim=0.12,
            width_opt=3.4, iqp_iters_min=3, iqp_curverror_allowed=0.01)
TOL = 1e-6          # oracle vs real package, max|diff| / max|ref|


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    if a.shape != b.shape:
        return float("inf")
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def raw_tracks() -> dict:
    """The reference's inputs/tracks/*.csv as shipped inside the committed fixture (bytes of the files)."""
    d = np.load(os.path.join(GOLD, "refback_import_track.npz"))
    out = {}
    for key in d.files:
        if key.endswith("_csv"):
            csv = np.loadtxt(io.BytesIO(d[key].tobytes()), comments="#", delimiter=",")
            out[key[:-4]] = csv[:, :4]
    return out


def prep_track(tph, track_imp):
    """The statements of the stock prep_track (prep_track.py:39-51) with the given tph."""
    rt = tph.spline_approximation.spline_approximation(track=track_imp, k_reg=PARS["k_reg"], s_reg=PARS["s_reg"],
                                                       stepsize_prep=PARS["stepsize_prep"],
                                                       stepsize_reg=PARS["stepsize_reg"], debug=False)
    path_cl = np.vstack((rt[:, :2], rt[0, :2]))
    cx, cy, A, nv = tph.calc_splines.calc_splines(path=path_cl)
    return rt, nv, A, cx, cy


def run_path(tph, rt, w_veh, kappa_bound, with_iqp, vel=None):
    """Everything the reference calls on its mincurv branch, with the given tph-like package.  Returns a dict with the
    keys of tests/golden/<name>.npz."""
    path_cl = np.vstack((rt[:, :2], rt[0, :2]))
    cx, cy, A, nv = tph.calc_splines.calc_splines(path=path_cl)
    n = rt.shape[0]
    A = np.asarray(A)
    scaling = np.array([-A[4 * i + 2, 4 * i + 5] for i in range(n - 1)] + [A[4 * n - 2, 1]])
    out = dict(reftrack=rt, w_veh=w_veh, kappa_bound=kappa_bound, coeffs_x=cx, coeffs_y=cy, normvec=nv, scaling=scaling,
               el_lengths=np.sqrt(np.sum(np.diff(path_cl, axis=0) ** 2, axis=1)))
    alpha, cerr = tph.opt_min_curv.opt_min_curv(reftrack=rt, normvectors=nv, A=A, kappa_bound=kappa_bound, w_veh=w_veh,
                                                print_debug=False, plot_debug=False)
    out.update(alpha_mincurv=alpha, curv_error_max=cerr)
    out["alpha_shpath"] = tph.opt_shortest_path.opt_shortest_path(reftrack=rt, normvectors=nv, w_veh=w_veh, print_debug=False)
    rl = tph.create_raceline.create_raceline(refline=rt[:, :2], normvectors=nv, alpha=alpha,
                                             stepsize_interp=PARS["stepsize_interp_after_opt"])
    psi, kappa = tph.calc_head_curv_an.calc_head_curv_an(coeffs_x=rl[2], coeffs_y=rl[3], ind_spls=rl[4], t_spls=rl[5])
    out.update(rl_raceline_interp=rl[0], rl_coeffs_x=rl[2], rl_coeffs_y=rl[3], rl_spline_inds=rl[4], rl_t_values=rl[5],
               rl_s=rl[6], rl_spline_lengths=rl[7], rl_el_lengths=rl[8], rl_psi=psi, rl_kappa=kappa)
    if with_iqp:
        a_i, rt_i, nv_i = tph.iqp_handler.iqp_handler(reftrack=rt.copy(), normvectors=nv, A=A, kappa_bound=kappa_bound,
                                                      w_veh=w_veh, print_debug=False, plot_debug=False,
                                                      stepsize_interp=PARS["stepsize_reg"], iters_min=PARS["iqp_iters_min"],
                                                      curv_error_allowed=PARS["iqp_curverror_allowed"])
        out.update(iqp_alpha=a_i, iqp_reftrack=rt_i, iqp_normvec=nv_i)
    if vel is not None:
        vx = tph.calc_vel_profile.calc_vel_profile(ggv=vel["ggv"], ax_max_machines=vel["ax_max_machines"], v_max=vel["v_max"],
                                                   kappa=kappa, el_lengths=rl[8], closed=True, filt_window=None,
                                                   dyn_model_exp=vel["dyn_model_exp"], drag_coeff=vel["dragcoeff"],
                                                   m_veh=vel["mass"])
        ax = tph.calc_ax_profile.calc_ax_profile(vx_profile=np.append(vx, vx[0]), el_lengths=rl[8], eq_length_output=False)
        t = tph.calc_t_profile.calc_t_profile(vx_profile=vx, ax_profile=ax, el_lengths=rl[8])
        out.update(vx=vx, ax=ax, t=t)
    return out


class OraclePackage:
    """oracle/ presented with the module layout of trajectory_planning_helpers (tph.<module>.<function>)."""

    def __init__(self):
        import types
        from oracle import tph_dense as T, tph_velprofile as VP
        self.T, self.VP = T, VP
        for name in ("calc_splines", "opt_min_curv", "opt_shortest_path", "create_raceline", "calc_head_curv_an", "iqp_handler"):
            setattr(self, name, types.SimpleNamespace(**{name: getattr(T, name)}))
        for name in ("calc_vel_profile", "calc_ax_profile", "calc_t_profile"):
            setattr(self, name, types.SimpleNamespace(**{name: getattr(VP, name)}))


def compare(real: dict, orc: dict) -> dict:
    keys = [k for k in real if k in orc and k not in ("reftrack", "w_veh", "kappa_bound")]
    return {k: rel(orc[k], real[k]) for k in keys}


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--write", action="store_true", help="write tests/golden_real/ (goldens from the real packages + pin.json)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden_real"))
    ap.add_argument("--quick", action="store_true", help="Berlin and the handling track only (no Modena, no synthetic fixtures)")
    args = ap.parse_args()
    try:
        import trajectory_planning_helpers as tph
        import quadprog  # noqa: F401  (tph.opt_min_curv imports it itself; fail early with a clear message)
    except Exception as e:  # pragma: no cover - depends on the environment
        print(f"pin_against_tph: the real packages are not importable here ({type(e).__name__}: {e}).\n"
              "  pip install trajectory-planning-helpers==0.76 quadprog   and run this tool again.", file=sys.stderr)
        return 2
    orc = OraclePackage()
    T, VP = orc.T, orc.VP
    vel = {k: v for k, v in np.load(os.path.join(GOLD, "velprofile.npz")).items()}
    vel = dict(ggv=vel["ggv"], ax_max_machines=vel["ax_max_machines"], v_max=float(vel["v_max"]), mass=float(vel["mass"]),
               dragcoeff=float(vel["dragcoeff"]), dyn_model_exp=float(vel["dyn_model_exp"]))
    raws = raw_tracks()
    report = {"versions": {"trajectory_planning_helpers": getattr(tph, "__version__", "?"),
                           "quadprog": getattr(sys.modules.get("quadprog"), "__version__", "?"), "numpy": np.__version__}}

    # ---- 1. the two constants ----
    rt_b, nv_b, A_b, _, _ = prep_track(tph, raws["berlin_2018"])
    a_real = tph.opt_min_curv.opt_min_curv(reftrack=rt_b, normvectors=nv_b, A=A_b, kappa_bound=PARS["curvlim"],
                                           w_veh=PARS["width_opt"], print_debug=False, plot_debug=False)[0]
    errs = {}
    for fs in (1.0, 2.0):
        T.F_SCALE = fs
        errs[fs] = rel(T.opt_min_curv(rt_b, nv_b, np.asarray(A_b), PARS["curvlim"], PARS["width_opt"])[0], a_real)
    f_scale = min(errs, key=errs.get)
    T.F_SCALE = f_scale
    report["f_scale"] = {"chosen": f_scale, "alpha_rel_err": {str(k): v for k, v in errs.items()}}
    print(f"f_scale: {f_scale}   (alpha error of the oracle vs real tph.opt_min_curv on Berlin: "
          + ", ".join(f"F={k}: {v:.2e}" for k, v in errs.items()) + ")")
    # velocity profile slice: a raceline whose start line lies in a braking zone separates the two readings
    real_b = run_path(tph, rt_b, PARS["width_opt"], PARS["curvlim"], False, vel)
    kap, el = real_b["rl_kappa"], real_b["rl_el_lengths"]
    v0 = real_b["vx"]
    shift = int(np.argmin(np.diff(np.append(v0, v0[0]))))            # strongest deceleration: put the start line there
    kap_r, el_r = np.roll(kap, -shift), np.roll(el, -shift)
    vx_real = tph.calc_vel_profile.calc_vel_profile(ggv=vel["ggv"], ax_max_machines=vel["ax_max_machines"], v_max=vel["v_max"],
                                                    kappa=kap_r, el_lengths=el_r, closed=True, filt_window=None,
                                                    dyn_model_exp=vel["dyn_model_exp"], drag_coeff=vel["dragcoeff"], m_veh=vel["mass"])
    serr = {}
    for upper in (True, False):
        VP.DECEL_LAP_SLICE_UPPER = upper
        serr[int(upper)] = rel(VP.calc_vel_profile(ggv=vel["ggv"], ax_max_machines=vel["ax_max_machines"], v_max=vel["v_max"],
                                                   kappa=kap_r, el_lengths=el_r, closed=True, filt_window=None,
                                                   dyn_model_exp=vel["dyn_model_exp"], drag_coeff=vel["dragcoeff"],
                                                   m_veh=vel["mass"]), vx_real)
    slice_upper = min(serr, key=serr.get)
    VP.DECEL_LAP_SLICE_UPPER = bool(slice_upper)
    report["decel_slice_upper"] = {"chosen": slice_upper, "vx_rel_err": {str(k): v for k, v in serr.items()}}
    print(f"decel_slice_upper: {slice_upper}   (vx error with the start line in a braking zone: "
          + ", ".join(f"{k}: {v:.2e}" for k, v in serr.items()) + ")")

    # ---- 2. every function, every track ----
    cases = {}
    for name, csv, w_veh in (("berlin", "berlin_2018", PARS["width_opt"]), ("handling", "handling_track", 2.0),
                             ("modena", "modena_2019", 2.0), ("rounded_rectangle", "rounded_rectangle", 2.0)):
        if csv in raws and not (args.quick and name in ("modena", "rounded_rectangle")):
            cases[name] = (prep_track(tph, raws[csv])[0], w_veh, PARS["curvlim"], name != "modena")
    for f in sorted(os.listdir(GOLD)):                                   # the synthetic fixtures: same reftrack, real packages
        if f.startswith(("synth", "berlin500")) and f.endswith(".npz") and not args.quick:
            g = np.load(os.path.join(GOLD, f))
            cases[f[:-4]] = (g["reftrack"], float(g["w_veh"]), float(g["kappa_bound"]), "iqp_alpha" in g.files)
    worst, table, goldens = 0.0, {}, {}
    for name, (rt, w_veh, kb, with_iqp) in cases.items():
        try:
            real = run_path(tph, rt, w_veh, kb, with_iqp, vel)
            mine = run_path(orc, rt, w_veh, kb, with_iqp, vel)
        except Exception as e:
            table[name] = {"error": f"{type(e).__name__}: {e}"}
            worst = float("inf")
            print(f"  {name:20s} FAILED: {table[name]['error']}")
            continue
        table[name] = compare(real, mine)
        goldens[name] = real
        w = max(table[name].values())
        worst = max(worst, w)
        bad = {k: f"{v:.1e}" for k, v in table[name].items() if v > TOL}
        print(f"  {name:20s} N={rt.shape[0]:5d}  max rel err {w:.2e}" + (f"   ABOVE {TOL:g}: {bad}" if bad else ""))
    report["oracle_vs_real"] = table
    report["pinned"] = bool(worst <= TOL)
    print(("PINNED: the oracle reproduces the real packages to %.1e" % worst) if report["pinned"] else
          ("NOT pinned: worst deviation %.2e > %g -- the restatement differs beyond the two constants" % (worst, TOL)))

    # ---- 3. goldens from the real packages ----
    if args.write:
        os.makedirs(args.out, exist_ok=True)
        for name, real in goldens.items():
            np.savez_compressed(os.path.join(args.out, name + ".npz"), **{k: np.asarray(v) for k, v in real.items()})
        json.dump(report, open(os.path.join(args.out, "pin.json"), "w"), indent=1)
        print(f"wrote {len(goldens)} fixtures and pin.json to {args.out}")
    return 0 if report["pinned"] else 1


if __name__ == "__main__":
    sys.exit(main())
