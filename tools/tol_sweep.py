"""Interior-point tolerance sweep: alpha error against the golden oracle solutions and iteration counts for several
mu_rel (debug knob MC_DEBUG_PDIP_MU_REL, read by mc_mincurv_pdip_batch at every call).  Run on the GPU box."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import global_racetrajectory_optimization_b200 as tph
from global_racetrajectory_optimization_b200 import batch as B_, synth

dev = torch.device("cuda")
cases = {}
for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
    g = np.load(f)
    if bool(g["kappa_rows_active"]):
        continue
    cases[os.path.basename(f)[:-4]] = g
base = synth.make_batch(100, 16, 1000)
rts = np.stack([synth.jitter_widths(base[i % 16], 1000 + i) for i in range(256)])
rtd = torch.tensor(rts, device=dev)
cx, cy, nvd, hd = B_.calc_splines_batch(rtd)
os.environ["MC_DEBUG_PDIP_MU_REL"] = "1e-13"
os.environ["MC_DEBUG_PDIP_DX_REL"] = "1e-9"
ref = B_.opt_min_curv_batch(rtd, nvd, hd, 0.12, 2.0)["alpha"].clone()
for mu, dxr in [("1e-10", "0"), ("1e-10", "1e-5"), ("1e-10", "1e-6"), ("1e-10", "1e-7"), ("1e-8", "1e-6"), ("1e-6", "1e-6"), ("1e-12", "0")]:
    os.environ["MC_DEBUG_PDIP_MU_REL"] = mu
    os.environ["MC_DEBUG_PDIP_DX_REL"] = dxr
    errs, its = [], []
    for name, g in cases.items():
        rt = g["reftrack"]
        path = np.vstack((rt[:, :2], rt[0, :2]))
        _, _, A, nv = tph.calc_splines.calc_splines(path=path)
        res = B_.opt_min_curv_batch(torch.tensor(rt, device=dev).unsqueeze(0), torch.tensor(nv, device=dev).unsqueeze(0),
                                    torch.tensor(A.h, device=dev).unsqueeze(0), float(g["kappa_bound"]), float(g["w_veh"]))
        a = res["alpha"][0, :rt.shape[0]].cpu().numpy()
        errs.append(np.abs(a - g["alpha_mincurv"]).max() / np.abs(g["alpha_mincurv"]).max())
        its.append(int(res["iters"][0]))
    res = B_.opt_min_curv_batch(rtd, nvd, hd, 0.12, 2.0)
    e = ((res["alpha"] - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)).cpu().numpy()
    it = res["iters"].cpu().numpy()
    print(f"mu_rel {mu} dx_rel {dxr}: golden max err {max(errs):.2e} iters {its} | batch256 N=1000: max err vs mu_rel=1e-13 {e.max():.2e} "
          f"mean {e.mean():.2e}, iters mean {it.mean():.2f} max {it.max()}, status {np.bincount(res['status'].cpu().numpy(), minlength=5).tolist()}", flush=True)
