"""Short single-GPU run of the mincurv path for ncu captures: B instances of N points, 1 warm-up + 1 measured call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from global_racetrajectory_optimization_b200 import batch as B_, synth
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 592
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda")
base = synth.make_batch(100, 16, N)
rts = np.stack([synth.jitter_widths(base[i % 16], 1000 + i) for i in range(Bn)])
rtd = torch.tensor(rts, device=dev)
cx, cy, nvd, hd = B_.calc_splines_batch(rtd)
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    res = B_.opt_min_curv_batch(rtd, nvd, hd, 0.12, 2.0)
    torch.cuda.synchronize(); dt = time.time() - t0
st = res["status"].cpu().numpy(); it = res["iters"].cpu().numpy()
print("prof_run", Bn, N, f"{dt*1e3:.2f} ms {Bn/dt:.0f} QP/s status", np.bincount(st[st >= 0], minlength=5).tolist(), "iters", it.min(), float(it.mean()), it.max())
