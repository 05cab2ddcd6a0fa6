#!/usr/bin/env python
"""Sweep in the style of BASELINE.json configs[3]: width-jitter variants x vehicle widths of one centre line, solved as
ONE batch on the device, with the race trajectory, the lap time and the check_traj flags of every variant -- plus the
reference's lap-time matrix (ggv scale x top speed) for the fastest variant.

    python examples/sweep_lap_times.py [n_variants=64] [n_points=1000]

Needs a B200 (there is no CPU fallback).  Inputs are synthetic (global_racetrajectory_optimization_b200.synth); a real
track goes through the reference's prep_track first and is passed as the same [N, 4] array."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from global_racetrajectory_optimization_b200 import batch, globaltraj, synth  # noqa: E402


def main():
    n_var = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    dev = torch.device("cuda")
    base = synth.make_track(7, n)
    widths = np.linspace(1.6, 3.4, 7)                                           # veh_width grid of the sweep
    tracks = np.stack([base if i == 0 else synth.jitter_widths(base, 100 + i) for i in range(n_var)])
    w_opt = torch.tensor([widths[i % widths.size] for i in range(n_var)], dtype=torch.float64, device=dev)
    # stock vehicle of /root/reference/inputs/veh_dyn_info: flat 12 m/s^2 ggv, 5.3 m/s^2 machine limit fading above 40 m/s
    ggv = np.column_stack((np.linspace(0.0, 72.0, 19), np.full(19, 12.0), np.full(19, 12.0)))
    mach = np.column_stack((np.linspace(0.0, 72.0, 19), np.interp(np.linspace(0.0, 72.0, 19), [0.0, 36.0, 72.0], [5.3, 5.3, 2.0])))
    pars = globaltraj.default_pars()
    pars["optim_opts"]["width_opt"] = w_opt
    out = globaltraj.globaltraj_batch(torch.tensor(tracks, device=dev), "mincurv", pars, ggv, mach)
    flags = batch.check_traj_flags(out, ggv, mach, pars["veh_params"]["v_max"], pars["veh_params"]["curvlim"])
    lap = out["laptime"].cpu().numpy()
    ok = (out["status"] == 0).cpu().numpy()
    print(f"{n_var} variants of a {n}-point track: {int(ok.sum())} solved, lap time {lap[ok].min():.2f} .. {lap[ok].max():.2f} s")
    for name, f in flags.items():
        print(f"  check_traj '{name}': {int(f.sum())} variant(s) would warn")
    best = int(np.argmin(np.where(ok, lap, np.inf)))
    scales = np.linspace(0.3, 1.0, 15)
    speeds = np.linspace(100.0, 150.0, 11) / 3.6
    ltm = batch.lap_time_matrix_batch(out["kappa"][best:best + 1], out["el_lengths_interp"][best:best + 1], ggv, mach, scales,
                                      speeds, pars["veh_params"]["dragcoeff"], pars["veh_params"]["mass"],
                                      n_pts=out["n_out"][best:best + 1])[0].cpu().numpy()
    print(f"lap-time matrix of variant {best} (rows: top speed 100..150 km/h, columns: ggv scale 0.3..1.0):")
    print(np.array2string(ltm, precision=2, max_line_width=200))


if __name__ == "__main__":
    main()
