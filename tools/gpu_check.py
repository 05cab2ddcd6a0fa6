"""GPU bring-up check: every kernel against the CPU oracle on small synthetic tracks, with dumps of
intermediate bands for offline debugging.  Run on the GPU box:  python tools/gpu_check.py"""
import os, sys, time, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import global_racetrajectory_optimization_b200 as tph
from global_racetrajectory_optimization_b200 import batch as B_, synth
from oracle import tph_dense as T
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from proto_banded_model import Model

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda")
rep = {}

layout = B_.mincurv_slab_layout
VEC = B_.SLAB_VECTORS

for N in [128, 200, 333]:
    rt = synth.make_track(1, N)
    path = np.vstack((rt[:, :2], rt[0, :2]))
    cx, cy, A, nv = T.calc_splines(path)
    gx, gy, gM, gnv = tph.calc_splines.calc_splines(path)
    r = dict(cx=float(np.abs(gx - cx).max()), cy=float(np.abs(gy - cy).max()), nv=float(np.abs(gnv - nv).max()))
    a_ref, e_ref = T.opt_min_curv(rt, nv, A, 0.12, 2.0)
    t0 = time.time()
    res = B_.opt_min_curv_batch(torch.tensor(rt, device=dev).unsqueeze(0), torch.tensor(nv, device=dev).unsqueeze(0),
                                torch.tensor(gM.h, device=dev).unsqueeze(0), 0.12, 2.0)
    torch.cuda.synchronize()
    a = res["alpha"][0].cpu().numpy()
    r.update(alpha_rel=float(np.abs(a - a_ref).max() / np.abs(a_ref).max()), status=int(res["status"][0]), iters=int(res["iters"][0]),
             curv_err=float(res["curv_error_max"][0]), curv_err_ref=e_ref, kmax=float(res["kappa_lin_max"][0]), t=time.time() - t0)
    # intermediate checks against the numpy model of the banded formulation
    ws = [v for k, v in B_._WS.items() if k[0] == "mincurv"][0].view(torch.float64).cpu().numpy()
    L = layout(N)
    md = Model(rt, nv, gM.h / np.roll(gM.h, -1), 0.12, 2.0, BZ=36)
    HBm = md.Hband()
    vec = lambda name: ws[VEC.index(name) * L["np"]: VEC.index(name) * L["np"] + N]
    HBg = ws[L["o_hb"]: L["o_hb"] + L["np"] * 34].reshape(-1, 34)[:N, :33]
    r.update(hb=float(np.abs(HBg - HBm).max() / np.abs(HBm).max()), f=float(np.abs(vec("F") - md.f).max() / np.abs(md.f).max()),
             kref=float(np.abs(vec("KREF") - md.kref).max()), xp=float(np.abs(vec("XP") - md.xp).max()))
    np.savez(os.path.join(OUT, f"dbg_{N}.npz"), ws=ws[:L["stride"]], alpha=a, alpha_ref=a_ref)
    a_sp_ref = T.opt_shortest_path(rt, nv, 2.0)
    a_sp = tph.opt_shortest_path.opt_shortest_path(rt, nv, 2.0)
    r.update(sp_rel=float(np.abs(a_sp - a_sp_ref).max() / np.abs(a_sp_ref).max()))
    ro = T.create_raceline(rt[:, :2], nv, a_ref, 2.0)
    rg = tph.create_raceline.create_raceline(rt[:, :2], nv, a_ref, 2.0)
    r.update(rl_n=(len(ro[0]), len(rg[0])))
    if len(ro[0]) == len(rg[0]):
        r.update(rl_xy=float(np.abs(ro[0] - rg[0]).max()), rl_t=float(np.abs(ro[5] - rg[5]).max()), rl_ind=int(np.abs(ro[4] - rg[4]).max()),
                 rl_s=float(np.abs(ro[6] - rg[6]).max()), rl_len=float(np.abs(ro[7] - rg[7]).max()), rl_el=float(np.abs(ro[8] - rg[8]).max()))
        po, ko = T.calc_head_curv_an(ro[2], ro[3], ro[4], ro[5])
        pg, kg = tph.calc_head_curv_an.calc_head_curv_an(rg[2], rg[3], rg[4], rg[5])
        r.update(psi=float(np.abs(po - pg).max()), kappa=float(np.abs(ko - kg).max() / np.abs(ko).max()))
    rep[N] = r
    print(N, json.dumps(r), flush=True)

# timing: batch of N=1000 tracks
for Bn, N in [(64, 1000), (592, 1000)]:
    rts = synth.make_batch(100, min(Bn, 16), N)
    rts = np.concatenate([rts] * ((Bn + len(rts) - 1) // len(rts)))[:Bn]
    rtd = torch.tensor(rts, device=dev)
    cx, cy, nvd, hd = B_.calc_splines_batch(rtd)
    for rep_i in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        res = B_.opt_min_curv_batch(rtd, nvd, hd, 0.12, 2.0)
        torch.cuda.synchronize(); dt = time.time() - t0
    st = res["status"].cpu().numpy(); it = res["iters"].cpu().numpy()
    fail = np.nonzero(st == 3)[0]; print("fail idx", fail[:8], "iters at fail", it[fail][:8]); print("timing", Bn, N, f"{dt*1e3:.1f} ms", f"{Bn/dt:.0f} QP/s", "status", np.bincount(st[st >= 0], minlength=5).tolist(), "iters", it.min(), it.mean(), it.max(), flush=True)
    rep[f"timing_{Bn}_{N}"] = dict(ms=dt * 1e3, qps=Bn / dt, iters_mean=float(it.mean()))
json.dump(rep, open(os.path.join(OUT, "gpu_check.json"), "w"), indent=1, default=str)
