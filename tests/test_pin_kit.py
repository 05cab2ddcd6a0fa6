"""The parity-pinning kit (tools/pin_against_tph.py) exercised end to end WITHOUT the real packages: a stand-in
``trajectory_planning_helpers`` / ``quadprog`` pair is generated in a temporary directory from oracle/ (with scipy for
spline_approximation) and the tool is run against it in a subprocess.  This checks the plumbing -- raw tracks out of the
committed fixture, the stock prep_track statements, detection of the two constants, the comparison table, the fixtures
and pin.json it writes -- so that the day the real packages are importable the one documented command just works.
With the stand-in's constants flipped the tool must report the flipped values."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIM = '''
import sys, types
sys.path.insert(0, {root!r})
sys.path.insert(0, {tools!r})
from oracle import tph_dense as _T, tph_velprofile as _VP
import make_golden as _mg
__version__ = "0.76-standin"
_T.F_SCALE = {f_scale}
def _mod(name, fn):
    m = types.ModuleType(__name__ + "." + name); setattr(m, name, fn); sys.modules[m.__name__] = m; globals()[name] = m
def _opt_min_curv(reftrack, normvectors, A, kappa_bound, w_veh, print_debug=False, plot_debug=False, **kw):
    keep = _T.F_SCALE; _T.F_SCALE = {f_scale}
    try: return _T.opt_min_curv(reftrack, normvectors, A, kappa_bound, w_veh)
    finally: _T.F_SCALE = keep
def _iqp(reftrack, normvectors, A, kappa_bound, w_veh, print_debug, plot_debug, stepsize_interp, iters_min=3, curv_error_allowed=0.01):
    keep = _T.F_SCALE; _T.F_SCALE = {f_scale}
    try: return _T.iqp_handler(reftrack, normvectors, A, kappa_bound, w_veh, False, False, stepsize_interp, iters_min, curv_error_allowed)
    finally: _T.F_SCALE = keep
def _vel(**kw):
    keep = _VP.DECEL_LAP_SLICE_UPPER; _VP.DECEL_LAP_SLICE_UPPER = {slice_upper}
    try: return _VP.calc_vel_profile(**kw)
    finally: _VP.DECEL_LAP_SLICE_UPPER = keep
def _spline_approximation(track, k_reg=3, s_reg=10, stepsize_prep=1.0, stepsize_reg=3.0, debug=False):
    return _mg.spline_approximation(track, k_reg, s_reg, stepsize_prep, stepsize_reg)
_mod("spline_approximation", _spline_approximation)
_mod("calc_splines", _T.calc_splines)
_mod("opt_min_curv", _opt_min_curv)
_mod("opt_shortest_path", lambda reftrack, normvectors, w_veh, print_debug=False: _T.opt_shortest_path(reftrack, normvectors, w_veh))
_mod("create_raceline", _T.create_raceline)
_mod("calc_head_curv_an", _T.calc_head_curv_an)
_mod("iqp_handler", _iqp)
_mod("calc_vel_profile", _vel)
_mod("calc_ax_profile", _VP.calc_ax_profile)
_mod("calc_t_profile", _VP.calc_t_profile)
'''


def _run_kit(tmp_path, f_scale, slice_upper):
    pkg = tmp_path / "site" / "trajectory_planning_helpers"
    pkg.mkdir(parents=True)
    (pkg / "__init__.py").write_text(textwrap.dedent(SHIM.format(root=ROOT, tools=os.path.join(ROOT, "tools"), f_scale=f_scale,
                                                                 slice_upper=slice_upper)))
    (tmp_path / "site" / "quadprog.py").write_text("__version__ = 'standin'\n")
    out = tmp_path / "golden_real"
    env = dict(os.environ, PYTHONPATH=str(tmp_path / "site") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pin_against_tph.py"), "--write", "--quick", "--out", str(out)],
                       capture_output=True, text=True, env=env, timeout=900)
    return r, out


@pytest.mark.parametrize("f_scale,slice_upper", [(2.0, True), (1.0, False)])
def test_pin_kit_detects_the_constants_and_writes_fixtures(tmp_path, f_scale, slice_upper):
    r, out = _run_kit(tmp_path, f_scale, slice_upper)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    pin = json.load(open(out / "pin.json"))
    assert pin["pinned"] is True
    assert pin["f_scale"]["chosen"] == f_scale and pin["decel_slice_upper"]["chosen"] == int(slice_upper)
    # the other reading is clearly rejected (the constants are observable on these tracks)
    assert pin["f_scale"]["alpha_rel_err"][str(3.0 - f_scale)] > 1e-2
    assert pin["decel_slice_upper"]["vx_rel_err"][str(1 - int(slice_upper))] > 1e-6
    import numpy as np
    g = np.load(out / "berlin.npz")
    assert {"reftrack", "alpha_mincurv", "alpha_shpath", "rl_kappa", "iqp_alpha", "vx", "t"} <= set(g.files)


def test_pin_kit_reports_missing_packages(tmp_path):
    env = dict(os.environ, PYTHONPATH="")
    r = subprocess.run([sys.executable, "-c", "import sys; sys.modules['trajectory_planning_helpers'] = None; "
                        "sys.argv = ['pin']; import runpy; runpy.run_path(%r, run_name='__main__')"
                        % os.path.join(ROOT, "tools", "pin_against_tph.py")], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 2 and "pip install trajectory-planning-helpers==0.76 quadprog" in r.stderr
