"""Single-GPU run of the velocity-profile stage alone (for timing and ncu captures of vel_profile_kernel).

    python tools/prof_velprofile.py [racelines=512] [reps=5] [--lib path/to/libmincurv_b200.so]

Inputs: the raceline kappa / el_lengths of the committed fixtures (tests/golden), cycled, rotated and perturbed to the
requested number of racelines; variants = the reference's lap-time matrix grid (14 ggv scales x 11 top speeds,
/root/reference/main_globaltraj.py:77-82).  One launch per repetition, timed with CUDA events."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from global_racetrajectory_optimization_b200 import build as _build  # noqa: E402

argv = [a for a in sys.argv[1:]]
if "--lib" in argv:
    k = argv.index("--lib")
    _build.LIB_PATH = os.path.abspath(argv[k + 1])
    _build.needs_build = lambda: False
    del argv[k:k + 2]
from global_racetrajectory_optimization_b200 import batch as B_  # noqa: E402

nb = int(argv[0]) if len(argv) > 0 else 512
reps = int(argv[1]) if len(argv) > 1 else 5
G = os.path.join(ROOT, "tests", "golden")
v = np.load(os.path.join(G, "velprofile.npz"))
rows = []
for name in ("synth1000", "berlin", "modena", "synth333", "handling"):
    g = np.load(os.path.join(G, name + ".npz"))
    rows.append((g["rl_kappa"], g["rl_el_lengths"]))
n_max = max(r[0].size for r in rows) + 3
kap = np.zeros((nb, n_max))
el = np.ones((nb, n_max))
npts = np.zeros(nb, dtype=np.int32)
for i in range(nb):
    k, e = rows[i % len(rows)]
    n = k.size
    u = np.arange(n) / n
    kap[i, :n] = np.roll(k, 7 * i) * (1.0 + 0.05 * np.sin(2 * np.pi * (3 * u + 0.1 * i)))
    el[i, :n] = np.roll(e, 7 * i)
    npts[i] = n
dev = torch.device("cuda")
kap_d, el_d, n_d = torch.tensor(kap, device=dev), torch.tensor(el, device=dev), torch.tensor(npts, device=dev)
scales = np.linspace(0.3, 1.0, int((1.0 - 0.3) / 0.05) + 1)
speeds = np.linspace(100.0 / 3.6, 150.0 / 3.6, int((150.0 - 100.0) / 5.0) + 1)
vm = torch.tensor(np.repeat(speeds, scales.size))
sc = torch.tensor(np.tile(scales, speeds.size))
args = dict(ggv=v["ggv"], ax_max_machines=v["ax_max_machines"], v_max=vm, drag_coeff=float(v["dragcoeff"]),
            m_veh=float(v["mass"]), n_pts=n_d, ggv_scales=sc, want_profiles=False)
res = B_.vel_profile_batch(kap_d, el_d, **args)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    res = B_.vel_profile_batch(kap_d, el_d, **args)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
P = nb * vm.numel()
pts = float(npts.mean())
lap = res["laptime"]
print(f"prof_velprofile lib={os.path.basename(_build.LIB_PATH)} racelines={nb} variants={vm.numel()} profiles={P} "
      f"points/profile={pts:.0f} ms/launch={ms:.3f} profiles/s={P / (ms * 1e-3):.0f} point-steps/s={P * pts / (ms * 1e-3):.3e} "
      f"status_bad={int((res['status'] != 0).sum())} laptime_checksum={float(lap.double().sum()):.9f}")
