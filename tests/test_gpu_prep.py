"""GPU tests (-m gpu) of the prep_track front end on the device (csrc/prep_track.cu): the reference's
four raw tracks through tph.spline_approximation's statements with the Reinsch smoothing spline.

Checker: oracle/tph_prep.spline_approximation_reinsch (dense numpy statement of the same algorithm; fixtures in
tests/golden/prep_track.npz from tools/make_golden_prep.py).  The distance to the scipy/FITPACK route -- what the
reference's prep_track produces -- is REPORTED and bounded loosely: FITPACK's adaptive-knot spline is a different
smoother with the same residual budget (and is itself defined only up to its 1e-3 tolerance on s)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import pin_against_tph as kit  # noqa: E402
import global_racetrajectory_optimization_b200 as tph  # noqa: E402
from global_racetrajectory_optimization_b200 import batch as B_  # noqa: E402

TRACKS = ["rounded_rectangle", "handling_track", "berlin_2018", "modena_2019"]


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(scope="module")
def fx():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "prep_track.npz"))), kit.raw_tracks()


def _dist_to_closed_polyline(pts, poly):
    """max over pts of the distance to the closed polyline through poly."""
    a, b = poly, np.roll(poly, -1, axis=0)
    ab = b - a
    t = np.clip(np.einsum("pij,ij->pi", pts[:, None, :] - a[None], ab) / np.maximum(np.einsum("ij,ij->i", ab, ab), 1e-300), 0.0, 1.0)
    d = np.linalg.norm(pts[:, None, :] - (a[None] + t[..., None] * ab[None]), axis=2)
    return float(d.min(axis=1).max())


@pytest.mark.parametrize("name", TRACKS)
def test_spline_approximation_matches_the_restatement(fx, name):
    gold, raws = fx
    got = tph.spline_approximation.spline_approximation(track=raws[name], k_reg=3, s_reg=10, stepsize_prep=1.0, stepsize_reg=3.0)
    ref = gold[name + "_reinsch"]
    assert got.shape == ref.shape
    assert np.abs(got[:, :2] - ref[:, :2]).max() <= 1e-7            # [m]
    assert np.abs(got[:, 2:] - ref[:, 2:]).max() <= 1e-5            # widths [m] (closest-point refinement: bounded scalar search vs safeguarded Newton)
    fit = gold[name + "_fitpack"]
    dist = _dist_to_closed_polyline(got[:, :2], fit[:, :2])         # (point counts may differ by one: compare curves, not indices)
    print(f"{name}: N = {got.shape[0]} (FITPACK route: {fit.shape[0]}), max distance to the FITPACK curve {dist:.3f} m")
    assert abs(got.shape[0] - fit.shape[0]) <= 2 and dist < 0.7     # same curve up to the choice of smoother


def test_batch_of_ragged_raw_tracks_and_min_width(fx):
    gold, raws = fx
    names = ["rounded_rectangle", "handling_track", "berlin_2018"]
    n_raw_max = max(raws[k].shape[0] for k in names)
    arr = np.zeros((3, n_raw_max, 4))
    for i, k in enumerate(names):
        arr[i, :raws[k].shape[0]] = raws[k]
    n_raw = torch.tensor([raws[k].shape[0] for k in names], dtype=torch.int32)
    out, n_out, lam = B_.spline_approximation_batch(torch.tensor(arr, device="cuda"), n_raw=n_raw)
    for i, k in enumerate(names):
        ref = gold[k + "_reinsch"]
        assert int(n_out[i]) == ref.shape[0]
        assert np.abs(out[i, :ref.shape[0]].cpu().numpy() - ref).max() <= 1e-5
    assert bool((lam > 0).all())
    mw, _, _ = B_.spline_approximation_batch(torch.tensor(raws["rounded_rectangle"][None], device="cuda"), min_width=6.0)
    ref = gold["rounded_rectangle_minwidth6"]
    assert np.abs(mw[0, :ref.shape[0]].cpu().numpy() - ref).max() <= 1e-5 and (ref[:, 2] + ref[:, 3]).min() >= 6.0 - 1e-12


def test_prep_track_mirror_feeds_the_path(fx):
    """helper_funcs_glob.src.prep_track.prep_track -> opt_min_curv, the reference's call sequence
    (/root/reference/main_globaltraj.py:252-271) on a raw track, entirely on the device."""
    gold, raws = fx
    pt = tph.helper_funcs_glob.src.prep_track.prep_track
    rt, nv, A, cx, cy = pt(reftrack_imp=raws["berlin_2018"], reg_smooth_opts=dict(k_reg=3, s_reg=10),
                           stepsize_opts=dict(stepsize_prep=1.0, stepsize_reg=3.0), debug=False, min_width=None)
    assert rt.shape == gold["berlin_2018_reinsch"].shape and nv.shape == (rt.shape[0], 2) and cx.shape == (rt.shape[0], 4)
    alpha, cerr = tph.opt_min_curv.opt_min_curv(reftrack=rt, normvectors=nv, A=A, kappa_bound=0.12, w_veh=3.4)
    assert np.all(alpha <= rt[:, 2] - 1.7 + 1e-6) and np.all(-alpha <= rt[:, 3] - 1.7 + 1e-6) and np.abs(alpha).max() > 1.0
    with pytest.raises(IOError, match="spline normals are crossed"):
        bad = raws["rounded_rectangle"].copy()
        bad[:, 2:] = 40.0                                # widths far beyond the corner radii: the normals must cross
        pt(bad, dict(k_reg=3, s_reg=10), dict(stepsize_prep=1.0, stepsize_reg=3.0), debug=False)
