"""Kernel-only timing of opt_min_curv_batch (setup + interior-point kernel) on the headline workload, for A/B runs of
library variants: MC_B200_LIB=<variant.so> python tools/ipm_time.py [B] [N] [reps] [alpha_out.npy | alpha_ref.npy]."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from global_racetrajectory_optimization_b200 import batch as B_, synth
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ref = sys.argv[4] if len(sys.argv) > 4 else None
dev = torch.device("cuda")
base = synth.make_batch(100, 16, N)
rts = np.stack([synth.jitter_widths(base[i % 16], 1000 + i) for i in range(Bn)])
rtd = torch.tensor(rts, device=dev)
cx, cy, nvd, hd = B_.calc_splines_batch(rtd)
ts = []
for r in range(reps + 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = B_.opt_min_curv_batch(rtd, nvd, hd, 0.12, 2.0)
    e1.record(); torch.cuda.synchronize()
    if r >= 2:
        ts.append(e0.elapsed_time(e1))
st = res["status"].cpu().numpy(); it = res["iters"].cpu().numpy(); a = res["alpha"].cpu().numpy()
msg = ""
if ref:
    if os.path.exists(ref):
        msg = " max|alpha - ref| %.2e" % np.abs(a - np.load(ref)).max()
    else:
        np.save(ref, a)
print("ipm_time %s B=%d N=%d: %.2f ms (min %.2f)  %.0f QP/s  status %s iters %.2f%s" % (
    os.path.basename(os.environ.get("MC_B200_LIB", "in-tree")), Bn, N, float(np.median(ts)), min(ts), Bn / (np.median(ts) * 1e-3),
    np.bincount(st[st >= 0], minlength=5).tolist(), float(it.mean()), msg))
