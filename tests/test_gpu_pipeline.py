"""GPU test (-m gpu) of the batched main flow (globaltraj.globaltraj_batch = /root/reference/main_globaltraj.py:252-532
for the three QP-based opt_types) end to end: prepared reference track in, race trajectory / lap time / check quantities /
export file out, against the same chain through the CPU oracle (tests/golden/<name>.npz + refback_<name>.npz)."""
import numpy as np
import pytest
import torch

import global_racetrajectory_optimization_b200 as tph
from global_racetrajectory_optimization_b200 import globaltraj
from global_racetrajectory_optimization_b200 import helper_funcs_glob as hf

pytestmark = pytest.mark.gpu
NAMES = ["berlin", "handling", "synth333"]


def _batch(golden):
    gs = [golden(n) for n in NAMES]
    n_max = max(g["reftrack"].shape[0] for g in gs)
    rt = np.zeros((len(gs), n_max, 4))
    for i, g in enumerate(gs):
        rt[i, :g["reftrack"].shape[0]] = g["reftrack"]
    dev = torch.device("cuda")
    npts = torch.tensor([g["reftrack"].shape[0] for g in gs], dtype=torch.int32, device=dev)
    # berlin 3.4 (racecar.ini), the others 2.0 -- float64: a float32 3.4 is 9.5e-8 m wider and moves alpha by half of that
    w_opt = torch.tensor([float(g["w_veh"]) for g in gs], dtype=torch.float64, device=dev)
    return gs, torch.tensor(rt, device=dev), npts, w_opt


def test_mincurv_flow_reproduces_the_oracle_chain_and_the_export_file(golden, tmp_path):
    v = golden("velprofile")
    gs, rt, npts, w_opt = _batch(golden)
    pars = globaltraj.default_pars()
    pars["optim_opts"]["width_opt"] = w_opt
    out = globaltraj.globaltraj_batch(rt, "mincurv", pars, v["ggv"], v["ax_max_machines"], n_pts=npts)
    assert out["status"].cpu().tolist() == [0, 0, 0] and out["vel_status"].cpu().tolist() == [0, 0, 0]
    for i, name in enumerate(NAMES):
        r = golden("refback_" + name)
        want = r["traj_race_cl"]                       # main_globaltraj.py:501-512 on the oracle's results
        m = want.shape[0]
        assert int(out["n_out"][i]) == m - 1
        got = out["trajectory"][i, :m].cpu().numpy()
        assert np.abs(got[:, 0] - want[:, 0]).max() <= 1e-6                      # s
        assert np.abs(got[:, 1:3] - want[:, 1:3]).max() <= 1e-5                  # x, y
        assert np.abs(np.angle(np.exp(1j * (got[:, 3] - want[:, 3])))).max() <= 1e-6      # psi (mod 2 pi)
        assert np.abs(got[:, 4] - want[:, 4]).max() <= 1e-3 * np.abs(want[:, 4]).max()    # kappa: north_star bar
        assert np.abs(got[:, 5] - want[:, 5]).max() <= 1e-5 * want[:, 5].max()   # vx
        assert np.abs(got[:, 6] - want[:, 6]).max() <= 1e-3                      # ax (difference of squares)
        assert abs(float(out["laptime"][i]) - float(v[name + "_t"][-1])) <= 1e-6 * float(v[name + "_t"][-1])
        assert abs(float(out["min_dist"][i]) - r["min_dists_full"].min()) <= 1e-5
    # the export of the first track, through the mirror of the reference's writer, against the reference-written file
    r = golden("refback_berlin")
    m = r["traj_race_cl"].shape[0]
    path = tmp_path / "traj_race_cl.csv"
    hf.src.export_traj_race.export_traj_race(file_paths=dict(traj_race_export=str(path)),
                                             traj_race=out["trajectory"][0, :m].cpu().numpy())
    lines = path.read_text().split("\n")
    assert lines[2] == "# s_m; x_m; y_m; psi_rad; kappa_radpm; vx_mps; ax_mps2"
    got = np.loadtxt(path, comments="#", delimiter=";")
    ref_rows = [ln for ln in str(r["traj_race_export"]).split("\n")[2:] if ln and ";" in ln][:-1]   # fixture is truncated
    ref = np.array([[float(x) for x in ln.split(";")] for ln in ref_rows])
    assert got.shape == (m, 7) and np.abs(got[:ref.shape[0], :6] - ref[:, :6]).max() <= 1e-4
    # single-track drop-in surface gives the same numbers as row 0 of the batch
    g = gs[0]
    path_cl = np.vstack((g["reftrack"][:, :2], g["reftrack"][0, :2]))
    _, _, A, nv1 = tph.calc_splines.calc_splines(path=path_cl)
    a1, _ = tph.opt_min_curv.opt_min_curv(reftrack=g["reftrack"], normvectors=nv1, A=A, kappa_bound=0.12, w_veh=3.4)
    assert np.abs(a1 - out["alpha"][0, :a1.size].cpu().numpy()).max() <= 1e-6 * np.abs(a1).max()


def test_iqp_and_shortest_path_flows(golden):
    v = golden("velprofile")
    gs, rt, npts, w_opt = _batch(golden)
    pars = globaltraj.default_pars()
    pars["optim_opts"]["width_opt"] = w_opt
    sp = globaltraj.globaltraj_batch(rt, "shortest_path", pars, v["ggv"], v["ax_max_machines"], n_pts=npts)
    mc = globaltraj.globaltraj_batch(rt, "mincurv", pars, v["ggv"], v["ax_max_machines"], n_pts=npts, check=False)
    iq = globaltraj.globaltraj_batch(rt, "mincurv_iqp", pars, v["ggv"], v["ax_max_machines"], n_pts=npts)
    for i, g in enumerate(gs):
        n = g["reftrack"].shape[0]
        assert np.abs(sp["alpha"][i, :n].cpu().numpy() - g["alpha_shpath"]).max() <= 1e-4 * np.abs(g["alpha_shpath"]).max()
        ni = int(iq["n_pts"][i])
        assert ni == g["iqp_alpha"].size
        assert np.abs(iq["alpha"][i, :ni].cpu().numpy() - g["iqp_alpha"]).max() <= 1e-3 * np.abs(g["iqp_alpha"]).max()
    for res in (sp, iq):
        assert res["status"].cpu().tolist() == [0, 0, 0] and res["vel_status"].cpu().tolist() == [0, 0, 0]
        lap = res["laptime"].cpu().numpy()
        assert np.all(np.isfinite(lap)) and np.all(lap > 0.0)
        t_last = torch.gather(res["t"], 1, res["n_out"].long().unsqueeze(1)).squeeze(1)
        assert torch.equal(t_last, res["laptime"])                               # laptime = t_profile_cl[-1]
        assert np.all(res["min_dist"].cpu().numpy() > -1e-9)
    # the shortest path is shorter than the minimum-curvature line
    len_sp = torch.gather(sp["trajectory"][:, :, 0], 1, sp["n_out"].long().unsqueeze(1)).squeeze(1)
    len_mc = torch.gather(mc["trajectory"][:, :, 0], 1, mc["n_out"].long().unsqueeze(1)).squeeze(1)
    assert bool((len_sp < len_mc).all())
    with pytest.raises(IOError):
        globaltraj.globaltraj_batch(rt, "mintime", pars, v["ggv"], v["ax_max_machines"], n_pts=npts)
