"""Parity against the REAL third-party packages (auto-skipped where they are not importable -- e.g. the offline build
container and the GPU box; see tools/pin_against_tph.py for the one command that makes them available to the suite).

With ``trajectory_planning_helpers`` and ``quadprog`` installed these tests run the reference's own call sequence
(/root/reference/helper_funcs_glob/src/prep_track.py:39-51, /root/reference/main_globaltraj.py:264-290,371-387,400-421)
with the real packages on Berlin / Modena / the handling track / the rounded rectangle and compare
  * the CPU oracle (oracle/tph_dense.py, oracle/tph_velprofile.py)           -- tolerance 1e-6, no GPU needed
  * the CUDA path through the C-ABI (global_racetrajectory_optimization_b200)  -- north_star tolerances, -m gpu
with them.  The two constants that cannot be confirmed offline are taken from tests/golden_real/pin.json when the
pinning kit has written it, else determined on the fly (f_scale) the same way the kit does."""
import os
import sys

import numpy as np
import pytest

tph = pytest.importorskip("trajectory_planning_helpers", reason="real tph not installed (pip install trajectory-planning-helpers==0.76)")
pytest.importorskip("quadprog", reason="real quadprog not installed")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import pin_against_tph as kit  # noqa: E402
from conftest import rel_max  # noqa: E402

TRACKS = [("berlin_2018", 3.4, True), ("handling_track", 2.0, True), ("modena_2019", 2.0, False), ("rounded_rectangle", 2.0, True)]


@pytest.fixture(scope="module")
def prepared():
    raws = kit.raw_tracks()
    return {name: kit.prep_track(tph, raws[name])[0] for name, _, _ in TRACKS}


@pytest.fixture(scope="module")
def vel():
    g = np.load(os.path.join(kit.GOLD, "velprofile.npz"))
    return dict(ggv=g["ggv"], ax_max_machines=g["ax_max_machines"], v_max=float(g["v_max"]), mass=float(g["mass"]),
                dragcoeff=float(g["dragcoeff"]), dyn_model_exp=float(g["dyn_model_exp"]))


@pytest.fixture(scope="module")
def oracle_pkg(prepared):
    """oracle/ with the constants the real packages imply (pin.json if present, else decided here on Berlin)."""
    orc = kit.OraclePackage()
    import conftest
    if conftest.PIN is None:
        rt = prepared["berlin_2018"]
        _, _, A, nv = tph.calc_splines.calc_splines(path=np.vstack((rt[:, :2], rt[0, :2])))
        a_real = tph.opt_min_curv.opt_min_curv(rt, nv, A, 0.12, 3.4, print_debug=False, plot_debug=False)[0]
        errs = {}
        for fs in (1.0, 2.0):
            orc.T.F_SCALE = fs
            errs[fs] = rel_max(orc.T.opt_min_curv(rt, nv, np.asarray(A), 0.12, 3.4)[0], a_real)
        orc.T.F_SCALE = min(errs, key=errs.get)
    return orc


@pytest.mark.parametrize("name,w_veh,with_iqp", TRACKS)
def test_oracle_reproduces_the_real_packages(prepared, vel, oracle_pkg, name, w_veh, with_iqp):
    rt = prepared[name]
    real = kit.run_path(tph, rt, w_veh, 0.12, with_iqp, vel)
    mine = kit.run_path(oracle_pkg, rt, w_veh, 0.12, with_iqp, vel)
    errs = kit.compare(real, mine)
    assert max(errs.values()) <= kit.TOL, {k: v for k, v in errs.items() if v > kit.TOL}


@pytest.mark.gpu
@pytest.mark.parametrize("name,w_veh,with_iqp", TRACKS)
def test_cuda_path_reproduces_the_real_packages(prepared, vel, oracle_pkg, name, w_veh, with_iqp):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import global_racetrajectory_optimization_b200 as mine
    from global_racetrajectory_optimization_b200 import batch
    batch.F_SCALE = float(oracle_pkg.T.F_SCALE)
    batch.VP_DECEL_SLICE_UPPER = int(bool(oracle_pkg.VP.DECEL_LAP_SLICE_UPPER))
    rt = prepared[name]
    real = kit.run_path(tph, rt, w_veh, 0.12, with_iqp, vel)
    got = kit.run_path(mine, rt, w_veh, 0.12, with_iqp, vel)
    assert rel_max(got["alpha_mincurv"], real["alpha_mincurv"]) <= 1e-4          # BASELINE.json north_star
    assert rel_max(got["alpha_shpath"], real["alpha_shpath"]) <= 1e-4
    assert got["rl_kappa"].shape == real["rl_kappa"].shape and rel_max(got["rl_kappa"], real["rl_kappa"]) <= 1e-3
    assert np.abs(got["rl_raceline_interp"] - real["rl_raceline_interp"]).max() <= 1e-3     # [m]
    assert rel_max(got["vx"], real["vx"]) <= 1e-3 and abs(got["t"][-1] - real["t"][-1]) <= 1e-3 * real["t"][-1]
    if with_iqp:
        assert got["iqp_alpha"].shape == real["iqp_alpha"].shape and rel_max(got["iqp_alpha"], real["iqp_alpha"]) <= 1e-3
