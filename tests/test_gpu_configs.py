"""GPU tests (-m gpu) shaped like the configurations of BASELINE.json that are not the bench line,
at reduced batch but full N, through the C-ABI:

  C2  Berlin centre line, N = 500, width-jitter variants x vehicle-width grid        (golden + live oracle)
  C3  iterative minimum curvature, 5 outer iterations per track, N = 1000            (oracle at small N, properties at N = 1000)
  C4  N = 2000 tracks, width-jitter x vehicle-width sweep                            (golden synth2000 + properties)
  C5  shortest-path QPs interleaved with minimum-curvature QPs on two streams, N = 500
"""
import numpy as np
import pytest
import torch

from conftest import rel_max
from global_racetrajectory_optimization_b200 import batch as B_
from global_racetrajectory_optimization_b200 import synth
from oracle import tph_dense as T

pytestmark = pytest.mark.gpu

ALPHA_TOL = 1e-4          # north_star: alpha within 1e-4 (max |d alpha| / max |alpha|)
BOX_TOL = 1e-6            # [m] active bounds are met to the interior-point tolerance, not to the last bit


def _box_violation(al, rts, wv):
    """Largest violation [m] of -(w_l - w_veh/2) <= alpha <= w_r - w_veh/2 over a batch."""
    rtn, wvn = np.stack(rts), np.asarray(wv, dtype=float)[:, None]
    return max(float((al - (rtn[:, :, 2] - wvn / 2)).max()), float((-(rtn[:, :, 3] - wvn / 2) - al).max()))


def _closed(rt):
    return np.vstack((rt[:, :2], rt[0, :2]))


def _oracle_alpha(rt, kappa_bound, w_veh):
    _, _, A, nv = T.calc_splines(_closed(rt))
    return T.opt_min_curv(rt, nv, A, kappa_bound, w_veh)[0]


def test_c2_berlin_n500_width_jitter_and_vehicle_width_grid(golden):
    ga, gb = golden("berlin500_jitter_a"), golden("berlin500_jitter_b")
    dev = torch.device("cuda")
    grid = np.linspace(1.6, 3.4, 7)
    rts, wv = [ga["reftrack"], gb["reftrack"]], [float(ga["w_veh"]), float(gb["w_veh"])]
    for i in range(30):
        rts.append(synth.jitter_widths(ga["reftrack"] if i % 2 == 0 else gb["reftrack"], 100 + i))
        wv.append(float(grid[i % grid.size]))
    rt = torch.tensor(np.stack(rts), device=dev)
    cx, cy, nv, h = B_.calc_splines_batch(rt)
    res = B_.opt_min_curv_batch(rt, nv, h, 0.12, torch.tensor(wv, device=dev))
    st = res["status"].cpu().numpy()
    assert np.all(st == 0), st
    al = res["alpha"].cpu().numpy()
    assert rel_max(al[0], ga["alpha_mincurv"]) <= ALPHA_TOL and rel_max(al[1], gb["alpha_mincurv"]) <= ALPHA_TOL
    for i in (5, 18):                                            # two variants against the live dense oracle
        assert rel_max(al[i], _oracle_alpha(rts[i], 0.12, wv[i])) <= ALPHA_TOL
    # every variant stays inside its own track with its own vehicle width
    assert _box_violation(al, rts, wv) <= BOX_TOL


@pytest.mark.parametrize("name", ["synth200", "synth333"])
def test_c3_five_outer_iterations_match_the_oracle(golden, name):
    g = golden(name)
    dev = torch.device("cuda")
    rt = g["reftrack"]
    _, _, A, onv = T.calc_splines(_closed(rt))
    hist = []
    oa, ort, onv2 = T.iqp_handler(rt.copy(), onv, A, float(g["kappa_bound"]), float(g["w_veh"]), False, False, 3.0, 5, 1e9,
                                  history=hist)
    assert len(hist) == 5
    rtd = torch.tensor(np.stack([rt, rt]), device=dev)
    cx, cy, nv, h = B_.calc_splines_batch(rtd)
    res = B_.iqp_batch(rtd, nv, h, float(g["kappa_bound"]), float(g["w_veh"]), 3.0, iters_min=5, fixed_iters=5)
    assert res["qp_solves"] == 10 and np.all(res["outer_iters"].cpu().numpy() == 5) and np.all(res["status"].cpu().numpy() == 0)
    n = int(res["n_pts"][0])
    assert n == oa.size == int(res["n_pts"][1])
    a = res["alpha"][0, :n].cpu().numpy()
    assert rel_max(a, oa) <= 20 * ALPHA_TOL                      # five chained QPs + resamplings
    assert np.abs(res["reftrack"][0, :n].cpu().numpy() - ort).max() <= 2e-3
    assert np.abs(res["normvec"][0, :n].cpu().numpy() - onv2).max() <= 2e-4
    assert torch.equal(res["alpha"][0], res["alpha"][1])         # identical instances -> identical results


def test_c3_five_outer_iterations_at_n1000():
    dev = torch.device("cuda")
    n, B = 1000, 6
    base = synth.make_batch(800, 3, n)
    rt = torch.tensor(np.concatenate([base, base]), device=dev)
    cx, cy, nv, h = B_.calc_splines_batch(rt)
    res = B_.iqp_batch(rt, nv, h, 0.12, 2.0, 3.0, fixed_iters=5)
    assert res["qp_solves"] == 5 * B
    assert np.all(res["status"].cpu().numpy() == 0) and np.all(res["outer_iters"].cpu().numpy() == 5)
    npn = res["n_pts"].cpu().numpy()
    # the synthetic tracks have ~3.35 m point spacing, the IQP re-samples the raceline every 3.0 m: N grows by ~10 %
    # (beyond n + 64: iqp_batch sizes its buffers from the track length)
    assert np.all(npn > n) and np.all(npn < 1.2 * n)
    for i in range(3):
        assert npn[i] == npn[i + 3] and torch.equal(res["alpha"][i], res["alpha"][i + 3])
    # the re-linearisation converges in tph's sense: the linearisation error of the last QP is below tph's default
    # curv_error_allowed (the shifts themselves do not vanish: with F_SCALE = 2 every QP steps past the
    # Gauss-Newton point)
    assert np.all(res["curv_error_max"].cpu().numpy() <= 0.01), res["curv_error_max"]
    # final reference lines are closed, roughly equidistant (3 m re-sampling) and inside the original track
    for i in range(3):
        m = int(npn[i])
        p = res["reftrack"][i, :m, :2].cpu().numpy()
        d = np.linalg.norm(np.diff(np.vstack((p, p[0])), axis=0), axis=1)
        assert d.max() < 3.3 and d.min() > 2.0
        w = res["reftrack"][i, :m, 2:].cpu().numpy()
        assert w.min() > 2.0 / 2 - 1e-6                          # never closer to a boundary than half the vehicle


def test_c4_n2000_golden_and_sweep(golden):
    g = golden("synth2000")
    dev = torch.device("cuda")
    rt0 = g["reftrack"]
    assert rt0.shape[0] == 2000
    grid = np.linspace(1.6, 3.4, 5)
    rts, wv = [rt0], [float(g["w_veh"])]
    for i in range(9):
        rts.append(synth.jitter_widths(rt0, 300 + i))
        wv.append(float(grid[i % grid.size]))
    rt = torch.tensor(np.stack(rts), device=dev)
    cx, cy, nv, h = B_.calc_splines_batch(rt)
    assert np.abs(cx[0].cpu().numpy() - g["coeffs_x"]).max() <= 1e-9 and np.abs(nv[0].cpu().numpy() - g["normvec"]).max() <= 1e-10
    res = B_.opt_min_curv_batch(rt, nv, h, float(g["kappa_bound"]), torch.tensor(wv, device=dev))
    assert np.all(res["status"].cpu().numpy() == 0)
    al = res["alpha"].cpu().numpy()
    assert rel_max(al[0], g["alpha_mincurv"]) <= ALPHA_TOL
    assert abs(float(res["curv_error_max"][0]) - float(g["curv_error_max"])) <= 1e-3 * float(g["curv_error_max"]) + 1e-6
    assert res["iters"].max().item() <= 30
    assert _box_violation(al, rts, wv) <= BOX_TOL
    # raceline / kappa of the N = 2000 fixture
    rl = B_.create_raceline_batch(rt[:1], nv[:1], res["alpha"][:1], 2.0)
    m = int(rl["n_out"][0])
    assert m == g["rl_kappa"].size
    assert rel_max(rl["kappa"][0, :m].cpu().numpy(), g["rl_kappa"]) <= 1e-3
    assert rel_max(rl["raceline_interp"][0, :m].cpu().numpy(), g["rl_raceline_interp"]) <= 1e-6
    # shortest path on the same N = 2000 track
    sp = B_.opt_shortest_path_batch(rt[:1], nv[:1], float(g["w_veh"]))
    assert rel_max(sp["alpha"][0].cpu().numpy(), g["alpha_shpath"]) <= ALPHA_TOL


def test_c5_shortest_path_interleaved_with_mincurv_on_two_streams(golden):
    g = golden("synth500")
    dev = torch.device("cuda")
    n = 500
    base = synth.make_batch(40, 8, n)
    rt_sp = np.stack([g["reftrack"]] + [synth.jitter_widths(base[i % 8], 500 + i) for i in range(511)])
    rt_mc = np.stack([g["reftrack"]] + [synth.jitter_widths(base[i % 8], 700 + i) for i in range(63)])
    d_sp, d_mc = torch.tensor(rt_sp, device=dev), torch.tensor(rt_mc, device=dev)
    _, _, nv_sp, _ = B_.calc_splines_batch(d_sp, want_coeffs=False)
    _, _, nv_mc, h_mc = B_.calc_splines_batch(d_mc, want_coeffs=False)
    # sequential reference run on the default stream
    sp0 = B_.opt_shortest_path_batch(d_sp, nv_sp, 2.0)
    mc0 = B_.opt_min_curv_batch(d_mc, nv_mc, h_mc, 0.12, 2.0)
    torch.cuda.synchronize()
    assert rel_max(sp0["alpha"][0].cpu().numpy(), g["alpha_shpath"]) <= ALPHA_TOL
    assert rel_max(mc0["alpha"][0].cpu().numpy(), g["alpha_mincurv"]) <= ALPHA_TOL
    assert np.all(sp0["status"].cpu().numpy() == 0) and np.all(mc0["status"].cpu().numpy() == 0)
    # interleaved: three rounds of both QP forms in flight at once on two streams
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    outs = []
    for _ in range(3):
        with torch.cuda.stream(s1):
            sp = B_.opt_shortest_path_batch(d_sp, nv_sp, 2.0)
        with torch.cuda.stream(s2):
            mc = B_.opt_min_curv_batch(d_mc, nv_mc, h_mc, 0.12, 2.0)
        outs.append((sp, mc))
    s1.synchronize()
    s2.synchronize()
    for sp, mc in outs:
        assert torch.equal(sp["alpha"], sp0["alpha"]) and torch.equal(sp["status"], sp0["status"])
        assert torch.equal(mc["alpha"], mc0["alpha"]) and torch.equal(mc["status"], mc0["status"])


def test_device_generated_variants_match_the_host_mirror():
    """mc_jitter_widths_batch (inputs of the C2 / C4 sweeps generated on the device from one 64-bit seed per variant)
    against synth.jitter_widths_hash, the numpy statement of the same stateless hash."""
    dev = torch.device("cuda")
    base = synth.make_batch(300, 3, 257)
    seeds = np.array([0, 1, 2, 12345678901, 2 ** 40 + 7, 99], dtype=np.int64)
    cid = np.array([0, 1, 2, 0, 1, 2], dtype=np.int32)
    out, n_out = B_.jitter_widths_batch(torch.tensor(base, device=dev), torch.tensor(seeds), rel=0.1, centre_id=torch.tensor(cid))
    out = out.cpu().numpy()
    assert n_out.cpu().tolist() == [257] * 6
    for v in range(6):
        ref = synth.jitter_widths_hash(base[cid[v]], int(seeds[v]), 0.1)
        assert np.array_equal(out[v, :, :2], ref[:, :2]) and np.abs(out[v, :, 2:] - ref[:, 2:]).max() <= 1e-12
    assert np.abs(out[0, :, 2] / base[0, :, 2] - 1.0).max() <= 0.1 + 1e-12 and np.abs(out[0, :, 2] - out[3, :, 2]).max() > 1e-3
    # default centre assignment: variant v uses track v % n_base
    out2, _ = B_.jitter_widths_batch(torch.tensor(base, device=dev), torch.tensor(seeds[:3]))
    assert np.array_equal(out2.cpu().numpy(), out[:3])


@pytest.mark.gpu
def test_shared_centre_lines_give_the_same_results_as_the_unshared_call():
    """Width variants of one track share H, f and k_ref: with centre_id the assembly runs once per centre line and is
    copied (mc_mincurv_solve_batch_shared).  The results must be bitwise those of the unshared call; a follower that
    points at a non-owner is refused with status -1."""
    dev = torch.device("cuda")
    n, n_base, V = 400, 3, 14
    base = synth.make_batch(4242, n_base, n)
    rts = np.stack([base[v % n_base] if v < n_base else synth.jitter_widths(base[v % n_base], 99 + v) for v in range(V)])
    rt = torch.tensor(rts, device=dev)
    cx, cy, nv, h = B_.calc_splines_batch(rt)
    w_veh = torch.linspace(1.8, 2.6, V, dtype=torch.float64, device=dev)
    ref = B_.opt_min_curv_batch(rt, nv, h, 0.12, w_veh)
    cid = B_.shared_centre_ids(torch.arange(V, device=dev) % n_base)
    assert cid.tolist() == [v % n_base for v in range(V)]
    got = B_.opt_min_curv_batch(rt, nv, h, 0.12, w_veh, centre_id=cid)
    assert torch.equal(got["status"], ref["status"]) and int((ref["status"] == 0).sum()) == V
    assert torch.equal(got["alpha"], ref["alpha"]) and torch.equal(got["curv_error_max"], ref["curv_error_max"])
    # chunked: the owner of a chunk is its first instance with that centre line
    got2 = B_.opt_min_curv_batch(rt, nv, h, 0.12, w_veh, centre_id=cid, max_chunk=5)
    assert torch.equal(got2["alpha"], ref["alpha"])
    bad = cid.clone()
    bad[7] = 4                     # instance 4 is a follower itself
    got3 = B_.opt_min_curv_batch(rt, nv, h, 0.12, w_veh, centre_id=bad)
    st = got3["status"].tolist()
    assert st[7] == -1 and all(s == 0 for i, s in enumerate(st) if i != 7)
