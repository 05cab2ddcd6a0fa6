#!/usr/bin/env python
"""Benchmark of the batched minimum-curvature QP path (BASELINE.json metric: min-curv QPs/s, N=1000 closed tracks).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference-style CPU path (oracle) on the host cores

One "step" = one pass of the hot path over one batch of synthetic closed tracks resident in HBM:
calc_splines -> assemble banded QP -> interior-point solve -> curvature check -> create_raceline + heading/curvature
(everything main_globaltraj.py does on its mincurv branch between prep_track and the velocity profile).
Prints ONE JSON line (rank 0).  See the task contract for the keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "min-curv QPs/sec (N=1000 closed track, batched)"
UNIT = "QP/s"
N_POINTS = 1000
BATCH_PER_GPU = 2368          # 4 instances per resident CTA (148 SMs x 4 CTAs)
N_BASE_LINES = 32          # distinct centre lines per rank; the rest of the batch are width-jitter variants
KAPPA_BOUND = 0.12
W_VEH = 2.0
STEP_INTERP = 2.0
ALG_BYTES_PER_POINT_K2 = 40.0      # 32 B reftrack row in + 8 B alpha out


def make_inputs(batch: int, n: int, seed0: int) -> np.ndarray:
    from global_racetrajectory_optimization_b200 import synth
    base = synth.make_batch(seed0, min(N_BASE_LINES, batch), n)
    out = np.empty((batch, n, 4))
    for i in range(batch):
        out[i] = base[i] if i < len(base) else synth.jitter_widths(base[i % len(base)], seed0 + 7919 * i)
    return out


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            def const(new, old):
                return getattr(pynvml, new, None) or getattr(pynvml, old)
            names = {const("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"): "hw_slowdown",
                     const("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"): "hw_thermal_slowdown",
                     const("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"): "sw_thermal_slowdown",
                     const("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"): "sw_power_cap"}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._halt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = get_reasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:  # clocks are evidence, not a dependency of the measurement
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
def run_b200(args) -> dict:
    import ctypes
    import torch
    import torch.distributed as dist
    from global_racetrajectory_optimization_b200 import _lib, batch as B_, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs CUDA devices (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # rank 0 must print exactly one line: some images export NCCL_DEBUG=VERSION, which makes NCCL write its banner to stdout
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    n, bl = args.npoints, args.batch
    total = bl * world                                   # weak scaling: fixed work per GPU
    s0, s1 = sharding.shard_range(total, rank, world)
    host = make_inputs(s1 - s0, n, seed0=10_000 + 1_000_003 * rank)
    host_pinned = torch.from_numpy(host).pin_memory()
    rt = host_pinned.to(dev, non_blocking=False)
    n_out_max = int(np.ceil(1.25 * n * 3.0 / STEP_INTERP)) + 64
    alpha_host = torch.empty((s1 - s0, n), dtype=torch.float64).pin_memory()
    status_host = torch.empty((s1 - s0,), dtype=torch.int32).pin_memory()

    ws = B_._workspace("mincurv", lib.mc_mincurv_workspace_bytes(s1 - s0, n), dev)
    ev = {"pdip": [], "setup": [], "splines": [], "raceline": []}   # one CUDA-event pair per timed step, on the launching stream
    state = {}

    def step(rt_dev, timed_kernels: bool):
        """One pass of the hot path; returns dict of device results (all launches on the current stream)."""
        if timed_kernels:
            for k in ev:
                ev[k].append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            ev["splines"][-1][0].record()
        cx, cy, nv, h = B_.calc_splines_batch(rt_dev, want_coeffs=False)
        if timed_kernels:
            ev["splines"][-1][1].record()
        Bq = rt_dev.shape[0]
        alpha = torch.empty((Bq, n), dtype=torch.float64, device=dev)
        cerr = torch.empty((Bq,), dtype=torch.float64, device=dev)
        kmax = torch.empty((Bq,), dtype=torch.float64, device=dev)
        st = torch.empty((Bq,), dtype=torch.int32, device=dev)
        iters = torch.empty((Bq,), dtype=torch.int32, device=dev)
        p = B_._ptr
        s = B_._stream()
        if timed_kernels:
            ev["setup"][-1][0].record()
        _lib.check(lib.mc_mincurv_setup_batch(Bq, n, None, p(rt_dev), p(nv), p(h), W_VEH, None, p(st), p(ws), ws.numel(), s), "setup")
        if timed_kernels:
            ev["setup"][-1][1].record()
            ev["pdip"][-1][0].record()
        _lib.check(lib.mc_mincurv_pdip_batch(Bq, n, None, p(alpha), p(st), p(iters), p(ws), ws.numel(), s), "pdip")
        if timed_kernels:
            ev["pdip"][-1][1].record()
        _lib.check(lib.mc_mincurv_finalize_batch(Bq, n, None, p(alpha), KAPPA_BOUND, p(cerr), p(kmax), p(st), p(ws), ws.numel(), s), "finalize")
        # curvature-row phase for the instances the box-only phase flagged (none on this workload: the kernel scans the
        # status words and returns) + re-evaluation -- together the five launches of mc_mincurv_solve_batch
        _lib.check(lib.mc_mincurv_kappa_batch(Bq, n, None, KAPPA_BOUND, p(alpha), p(st), p(iters), p(ws), ws.numel(), s), "kappa")
        _lib.check(lib.mc_mincurv_finalize_batch(Bq, n, None, p(alpha), KAPPA_BOUND, p(cerr), p(kmax), p(st), p(ws), ws.numel(), s), "finalize")
        if timed_kernels:
            ev["raceline"][-1][0].record()
        rl = B_.create_raceline_batch(rt_dev, nv, alpha, STEP_INTERP, n_out_max=n_out_max, with_head_curv=True)
        if timed_kernels:
            ev["raceline"][-1][1].record()
        return dict(alpha=alpha, status=st, iters=iters, kappa=rl["kappa"], raceline=rl["raceline_interp"], n_out=rl["n_out"],
                    el=rl["el_lengths_interp"])

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement ("value") ----
    for _ in range(args.warmup):
        res = step(rt, False)
        if world > 1:
            sharding.gather_batch(res["alpha"], total)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        res = step(rt, True)
        if world > 1:
            state["gathered"] = sharding.gather_batch(res["alpha"], total)     # the single collective of the path
        state["last"] = res
    t1.record()
    sync_all()
    clocks = sampler.stop()
    ms_local = t0.elapsed_time(t1)
    # average launch duration of the dominant kernels over the timed steps (CUDA events on the launching stream)
    pdip_last = float(np.mean([a.elapsed_time(b) for a, b in ev["pdip"]]))
    setup_last = float(np.mean([a.elapsed_time(b) for a, b in ev["setup"]]))
    splines_last = float(np.mean([a.elapsed_time(b) for a, b in ev["splines"]]))
    raceline_last = float(np.mean([a.elapsed_time(b) for a, b in ev["raceline"]]))
    tmax = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    res = state["last"]
    st = res["status"].cpu().numpy()
    ok_frac = float(np.mean((st == 0) | (st == 4)))

    # ---- end-to-end measurement ("e2e"): host buffers in, alpha + status out, every step ----
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rt2 = torch.empty_like(rt)
    for _ in range(min(args.warmup, 2)):
        rt2.copy_(host_pinned, non_blocking=True)
        r2 = step(rt2, False)
        alpha_host.copy_(r2["alpha"], non_blocking=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        rt2.copy_(host_pinned, non_blocking=True)
        r2 = step(rt2, False)
        alpha_host.copy_(r2["alpha"], non_blocking=True)
        status_host.copy_(r2["status"], non_blocking=True)
    e1.record()
    sync_all()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())

    qps = total * args.steps / (ms_total * 1e-3)
    e2e_qps = total * args.steps / (e2e_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = ALG_BYTES_PER_POINT_K2 * n * (s1 - s0)
    achieved = alg_bytes / (pdip_last * 1e-3) / 1e9
    traffic = None
    try:  # per-launch DRAM bytes of the dominant kernel from the committed ncu capture (B=592 launch, scaled per QP)
        prof = json.load(open(os.path.join(ROOT, "profiles", "pdip_traffic.json")))
        traffic = float(prof["dram_bytes_per_qp"]) * (s1 - s0)
    except Exception:
        pass
    # per-kernel HBM roofline fractions (SURVEY 8d algorithmic bytes: K1 96 B/point; K2 = assembly + solve 40 B/point;
    # K3 40 B/point in + 40 B per raceline station out), each over its own event-timed duration
    n_stations = float(res["n_out"].double().mean().item())
    def _k(name, bytes_per_qp, ms):
        gbs = bytes_per_qp * (s1 - s0) / (ms * 1e-3) / 1e9
        return {"kernel": name, "ms": ms, "algorithmic_bytes_per_qp": bytes_per_qp, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
    per_kernel = [_k("calc_splines_kernel", 96.0 * n, splines_last),
                  _k("mincurv_setup_kernel + mincurv_pdip_kernel", ALG_BYTES_PER_POINT_K2 * n, setup_last + pdip_last),
                  _k("create_raceline_kernel (+ psi/kappa)", 40.0 * n + 40.0 * n_stations, raceline_last)]
    line = {
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"batch {bl} per GPU of synthetic closed tracks, N={n} points, mincurv (non-iterative) "
                               "QP + raceline/kappa evaluation",
                   "global_batch": total, "n_points": n, "kappa_bound": KAPPA_BOUND, "w_veh": W_VEH,
                   "parallelism": f"batch-sharded x{world}, one all-gather of alpha",
                   "l2": "per-step working set (factor tiles + bands, ~2 MB per QP) exceeds the 126 MB L2; no explicit flush",
                   "solved_ok_fraction": ok_frac,
                   "ipm_iters_mean": float(res["iters"].double().mean().item())},
        "clocks": clocks,
        "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": int(host_pinned.numel() * 8),
                "d2h_bytes_per_step": int(alpha_host.numel() * 8 + status_host.numel() * 4)},
        "gpu_launches": 7 * args.steps,
        "roofline": {"bound": "hbm", "kernel": "mincurv_pdip_kernel", "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                     "kernel_ms": pdip_last, "setup_kernel_ms": setup_last, "per_kernel": per_kernel,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     # implementation traffic (ncu dram bytes per QP x QPs of this launch) over the live kernel time
                     "traffic_gbs": (traffic / (pdip_last * 1e-3) / 1e9) if traffic else None,
                     "traffic_frac_of_peak": (traffic / (pdip_last * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                     "note": "algorithmic bytes = 40 B/point (SURVEY 8d). The kernel re-streams its block-Cholesky factor "
                             "(written once, read four times per interior-point iteration): traffic >> algorithmic bytes by "
                             "construction; it is bound by the serial pivot/sweep chain of one warp per instance and by that "
                             "implementation traffic, see DESIGN.md section 5"},
    }
    if world == 1:
        line["next_stage"] = velprofile_stage(B_, res, dev)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(n, sample=1)
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


# ------------------------------------------------------------------------------------------------
def velprofile_stage(B_, res: dict, dev) -> dict:
    """Outside the timed region and not part of `value`: the stage after the path on the racelines the
    last step produced -- the reference's lap-time matrix (15 ggv scales x 11 top speeds,
    /root/reference/main_globaltraj.py:77-82, :442-496) for the first 512 racelines in one launch of vel_profile_kernel.
    Reported for information; a failure here is recorded, never raised."""
    import torch
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "velprofile.npz"))
        scales = np.linspace(0.3, 1.0, int((1.0 - 0.3) / 0.05) + 1)
        speeds = np.linspace(100.0 / 3.6, 150.0 / 3.6, int((150.0 - 100.0) / 5.0) + 1)
        nb = min(512, res["kappa"].shape[0])
        kap, el, npts = res["kappa"][:nb].contiguous(), res["el"][:nb].contiguous(), res["n_out"][:nb].contiguous()
        args = (kap, el, g["ggv"], g["ax_max_machines"], scales, speeds, float(g["dragcoeff"]), float(g["mass"]))
        B_.lap_time_matrix_batch(*args, n_pts=npts)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        for _ in range(reps):
            ltm = B_.lap_time_matrix_batch(*args, n_pts=npts)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        profiles = nb * scales.size * speeds.size
        pts = float(npts.double().mean().item())
        return {"stage": "calc_vel_profile + calc_ax_profile + calc_t_profile (lap-time matrix)", "kernel": "vel_profile_kernel",
                "racelines": nb, "variants_per_raceline": int(scales.size * speeds.size), "profiles": profiles,
                "points_per_profile_mean": pts, "ms_per_launch": ms, "profiles_per_s": profiles / (ms * 1e-3),
                "lap_time_s_stock_car_mean": float(ltm[:, -1, -1].mean().item()),
                "note": "includes the launch-side host work of lap_time_matrix_batch (table upload, status check)"}
    except Exception as e:      # informational stage: record, do not fail the bench line
        return {"stage": "calc_vel_profile", "error": f"{type(e).__name__}: {e}"}


# ------------------------------------------------------------------------------------------------
def _oracle_one(rt: np.ndarray) -> float:
    """One reference-style CPU solve of the same path (dense tph restatement + Goldfarb-Idnani)."""
    from oracle import tph_dense as T
    t0 = time.perf_counter()
    path = np.vstack((rt[:, :2], rt[0, :2]))
    cx, cy, A, nv = T.calc_splines(path)
    alpha, _ = T.opt_min_curv(rt, nv, A, KAPPA_BOUND, W_VEH)
    rl = T.create_raceline(rt[:, :2], nv, alpha, STEP_INTERP)
    T.calc_head_curv_an(rl[2], rl[3], rl[4], rl[5])
    return time.perf_counter() - t0


def cpu_baseline(n: int, sample: int = 1) -> dict:
    from oracle import quadprog_gi
    quadprog_gi.build()
    rts = make_inputs(sample, n, seed0=10_000)
    secs = sum(_oracle_one(rts[i]) for i in range(sample))
    cores = os.cpu_count() or 1
    return {"value": sample / secs, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} QP(s) of the same workload (N={n}) through oracle/tph_dense.py (dense 4N x 4N numpy/LAPACK, "
                      f"BLAS threads <= {cores}) + oracle/quadprog_gi.c (1 thread); {secs:.1f} s"}


def _oracle_worker(job):
    """One reference-style solve in a worker process with a bounded BLAS pool (several workers share the host cores)."""
    rt, blas_threads = job
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=blas_threads):
        return _oracle_one(rt)


def run_reference(args) -> dict:
    """Reference arm: the reference-style CPU path (dense numpy/LAPACK tph restatement + Goldfarb-Idnani in C) on ALL host
    cores: `workers` processes solve different QPs of the workload at the same time, each with cores/workers BLAS threads
    (a single solve does not scale past ~8 threads: the 4N x 4N inverse is the only threaded part).  One step = `workers`
    QPs in flight; the run is bounded to a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import multiprocessing as mp
    from oracle import quadprog_gi
    quadprog_gi.build()
    n = args.npoints
    cores = os.cpu_count() or 1
    workers = max(1, min(16, cores // 8))
    if os.environ.get("MC_REF_WORKERS"):                 # override for experiments
        workers = max(1, int(os.environ["MC_REF_WORKERS"]))
    blas_threads = max(1, cores // workers)
    budget_s = 240.0
    rts = make_inputs(workers, n, seed0=10_000)
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        jobs = [(rts[i], blas_threads) for i in range(workers)]
        t0 = time.perf_counter()
        pool.map(_oracle_worker, jobs)                   # warm-up step (also sizes the run)
        t_first = time.perf_counter() - t0
        warm = 1
        steps = max(1, min(args.steps, int(budget_s / max(t_first, 1e-3)) - 1))
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_oracle_worker, jobs)
        secs = time.perf_counter() - t0
    qps = steps * workers / secs
    sample = (f"each step = {workers} QPs of the workload (N={n}) solved concurrently by {workers} processes x {blas_threads} BLAS "
              f"threads through the dense numpy/LAPACK tph restatement + Goldfarb-Idnani C solver; {steps} steps timed "
              f"(run bounded to ~{budget_s:.0f} s)")
    return {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic closed tracks, N={n} points, mincurv (non-iterative) QP + raceline/kappa evaluation",
                       "n_points": n, "kappa_bound": KAPPA_BOUND, "w_veh": W_VEH, "qps_per_step": workers},
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="QP instances per GPU")
    ap.add_argument("--npoints", type=int, default=N_POINTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        line = run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch ourselves under torchrun (the driver launches torchrun itself)
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__), "--gpus", str(args.gpus),
                   "--steps", str(args.steps), "--warmup", str(args.warmup), "--batch", str(args.batch),
                   "--npoints", str(args.npoints)]
            sys.exit(subprocess.call(cmd))
        line = run_b200(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
