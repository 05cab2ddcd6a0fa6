"""Interior-point start-point / step-fraction sweep (debug knobs MC_DEBUG_PDIP_LAM0, MC_DEBUG_PDIP_ETA): mean iteration
count, time and alpha difference against the default on the headline workload.  Run on the GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from global_racetrajectory_optimization_b200 import batch as B_, synth
Bn, N = 2368, 1000
dev = torch.device("cuda")
base = synth.make_batch(100, 16, N)
rts = np.stack([synth.jitter_widths(base[i % 16], 1000 + i) for i in range(Bn)])
rtd = torch.tensor(rts, device=dev)
cx, cy, nvd, hd = B_.calc_splines_batch(rtd)
cid = (torch.arange(Bn, device=dev) % 16).to(torch.int32)
ref = None
for lam0 in ("1e-2", "1e-3", "1e-1", "1", "1e-4"):
    for eta in ("0.995", "0.999", "0.9999", "0.99"):
        os.environ["MC_DEBUG_PDIP_LAM0"], os.environ["MC_DEBUG_PDIP_ETA"] = lam0, eta
        ts = []
        for r in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = B_.opt_min_curv_batch(rtd, nvd, hd, 0.12, 2.0, centre_id=cid); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        a = res["alpha"]; it = res["iters"].float(); st = res["status"]
        if ref is None:
            ref = a.clone()
        print("lam0 %-5s eta %-6s: %.2f ms  iters mean %.2f max %d  status!=0 %d  max|alpha - default| %.2e" % (
            lam0, eta, min(ts), float(it.mean()), int(it.max()), int((st != 0).sum()), float((a - ref).abs().max())), flush=True)
