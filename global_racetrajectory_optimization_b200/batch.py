"""Batched device API of the B200 minimum-curvature path.

Thin host logic above the C-ABI (include/mincurv_b200.h): torch tensors are used only as device
buffer holders (allocation, streams); every computation happens in the CUDA kernels of
``libmincurv_b200.so``.  There is no CPU fallback -- without a CUDA device these functions raise.

The batched entry points have no counterpart in the reference (which solves one track per call);
the single-track modules next to this file (``opt_min_curv.py`` ...) mirror the tph call surface
used by /root/reference/main_globaltraj.py:264-290,371-387 on top of them.
"""
from __future__ import annotations

import ctypes
import functools
import math
from typing import Optional, Union

import torch

from . import _lib

_GRID_Y_MAX = 65535  # tracks per launch of the kernels that put the track index on gridDim.y
N_MIN = 80          # smallest closed track the banded solver supports (csrc/common.cuh N_MIN)
STATUS_TEXT = {
    0: "ok",
    1: "Problem not solvable, track might be too small to run with current safety distance!",
    2: "interior-point iteration cap reached",
    3: "numerical breakdown (non-positive pivot)",
    4: "curvature rows still violated after the curvature-row phase (constraints inconsistent)",
    -1: "unsupported track size",
}

# The two constants of trajectory_planning_helpers that cannot be confirmed offline (include/mincurv_b200.h,
# DESIGN.md section 2).  Run-time parameters of the C-ABI (*_ex entry points); tools/pin_against_tph.py determines them
# from the real package when it is importable and tests/test_real_tph.py then runs with what it found.
F_SCALE = 2.0               # tph.opt_min_curv: f = F_SCALE * E^T k_ref
VP_DECEL_SLICE_UPPER = 1    # tph.calc_vel_profile (closed): half of the doubled lap kept after the backward pass

_WS = {}

# mirror of csrc/mincurv_ws.cuh (debugging / tests read intermediate results out of the workspace)
SLAB_VECTORS = ("H DIAG DFW DBW LFW INVD TII RHOP RHOM PX PY NX NY MX MY XP YP SX SY KREF LB UB F "
                "T0 T1 T2 T3 T4 T5 ALPHA LU LL RD RHS DX DD DLU DLL SU SL ISU ISL YPAD "
                "S3 S4 L3 L4 KL WK EDX T3K T4K VV IH").split()
HB_PITCH = 34
ZB_PITCH = 108


def mincurv_slab_layout(n_max: int) -> dict:
    np_ = ((n_max + 31) // 32) * 32 + 64
    nb_max = max(1, (n_max - 32 + 31) // 32)
    o = len(SLAB_VECTORS) * np_
    o_zb = o
    o += n_max * ZB_PITCH
    o_hb = o
    o += np_ * HB_PITCH
    o_tiles = o
    o += np_ * 76
    return dict(np=np_, nb_max=nb_max, o_zb=o_zb, o_hb=o_hb, o_tiles=o_tiles, stride=(o + 15) & ~15)


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise _lib.MinCurvLibError("global_racetrajectory_optimization_b200 needs a CUDA device (B200, sm_100a); "
                                   "there is no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device_guard(fn):
    """Run fn with the device of its first CUDA tensor argument current: the library launches on the calling thread's
    current device and on that device's current stream."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = next((a.device for a in list(args) + list(kwargs.values()) if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None:
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _workspace(kind: str, nbytes: int, device) -> torch.Tensor:
    """Scratch buffer of the library for one (kind, device, stream): calls on different streams never share slabs or the
    work counter of the persistent solver kernels.  A buffer is allocated while its stream is current, so when a larger
    one replaces it the caching allocator re-uses the old block in that stream's order only."""
    dev = torch.device(device)
    key = (kind, str(dev), int(torch.cuda.current_stream(dev).cuda_stream) if dev.type == "cuda" else 0)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        _WS[key] = None
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def release_workspaces() -> None:
    _WS.clear()


def _f64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float64:
        raise TypeError(f"{name} must be float64")
    return t.contiguous()


def _npts(n_pts, B, device):
    if n_pts is None:
        return None
    n_pts = n_pts.to(device=device, dtype=torch.int32).contiguous()
    if n_pts.numel() != B:
        raise ValueError("n_pts must have one entry per track")
    return n_pts


# ------------------------------------------------------------------------------------------------
@_device_guard
def calc_splines_batch(xy: torch.Tensor, n_pts: Optional[torch.Tensor] = None,
                       el_lengths: Optional[torch.Tensor] = None, use_dist_scaling: bool = True,
                       want_coeffs: bool = True):
    """Closed cubic splines through the points of every track.

    xy: [B, n_max, 2] (or a reftrack [B, n_max, 4]); returns (coeffs_x, coeffs_y, normvec, h)."""
    _require_cuda()
    lib = _lib.load()
    xy = _f64(xy, "xy")
    B, n_max, stride = xy.shape
    dev = xy.device
    n_pts = _npts(n_pts, B, dev)
    cx = torch.zeros((B, n_max, 4), dtype=torch.float64, device=dev) if want_coeffs else None
    cy = torch.zeros((B, n_max, 4), dtype=torch.float64, device=dev) if want_coeffs else None
    nv = torch.zeros((B, n_max, 2), dtype=torch.float64, device=dev)
    h = torch.ones((B, n_max), dtype=torch.float64, device=dev)
    if el_lengths is not None:
        el_lengths = _f64(el_lengths, "el_lengths")
    nbytes = lib.mc_calc_splines_workspace_bytes(B, n_max)
    ws = _workspace("splines", nbytes, dev)
    rc = lib.mc_calc_splines_batch(B, n_max, _ptr(n_pts), _ptr(xy), stride, _ptr(el_lengths), int(bool(use_dist_scaling)),
                                   _ptr(cx), _ptr(cy), _ptr(nv), _ptr(h), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "mc_calc_splines_batch")
    return cx, cy, nv, h


def _wveh(w_veh, B, dev):
    if isinstance(w_veh, torch.Tensor):
        wb = w_veh.to(device=dev, dtype=torch.float64).contiguous()
        if wb.numel() != B:
            raise ValueError("w_veh tensor must have one entry per track")
        return 0.0, wb
    return float(w_veh), None


def _chunk(B: int, per_item_bytes: int, device) -> int:
    free, _ = torch.cuda.mem_get_info(device)
    budget = max(int(free * 0.6), per_item_bytes)
    return max(1, min(B, budget // max(per_item_bytes, 1)))


@_device_guard
def opt_min_curv_batch(reftrack: torch.Tensor, normvec: torch.Tensor, h: torch.Tensor, kappa_bound: float,
                       w_veh: Union[float, torch.Tensor], n_pts: Optional[torch.Tensor] = None,
                       max_chunk: Optional[int] = None, f_scale: Optional[float] = None,
                       centre_id: Optional[torch.Tensor] = None) -> dict:
    """Batched tph.opt_min_curv (closed tracks).  Returns a dict of device tensors:
    alpha [B, n_max], curv_error_max [B], kappa_lin_max [B], status [B] (int32), iters [B] (int32).
    f_scale: None = the module default F_SCALE (see there).
    centre_id [B] (optional): for batches in which several instances share a centreline (x, y, normvec, h, n_pts identical;
    only the widths differ -- a width sweep of one track): centre_id[b] = index of the instance that owns b's centreline
    (owners: centre_id[b] == b).  H, f and k_ref are then assembled once per centreline (same results, less work);
    shared_centre_ids() builds the tensor from group labels."""
    _require_cuda()
    lib = _lib.load()
    reftrack = _f64(reftrack, "reftrack")
    normvec = _f64(normvec, "normvec")
    h = _f64(h, "h")
    B, n_max, four = reftrack.shape
    if four != 4 or normvec.shape != (B, n_max, 2) or h.shape != (B, n_max):
        raise RuntimeError("Array size of reftrack should be the same as normvectors!")
    if n_max < N_MIN:
        raise NotImplementedError(f"closed tracks with fewer than {N_MIN} points are not supported by the banded solver")
    dev = reftrack.device
    n_pts = _npts(n_pts, B, dev)
    w_scalar, w_batch = _wveh(w_veh, B, dev)
    alpha = torch.empty((B, n_max), dtype=torch.float64, device=dev)
    cerr = torch.empty((B,), dtype=torch.float64, device=dev)
    kmax = torch.empty((B,), dtype=torch.float64, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    iters = torch.empty((B,), dtype=torch.int32, device=dev)
    per_item = lib.mc_mincurv_workspace_bytes(1, n_max)
    chunk = _chunk(B, per_item, dev) if max_chunk is None else min(B, max_chunk)
    ws = _workspace("mincurv", lib.mc_mincurv_workspace_bytes(chunk, n_max), dev)
    if centre_id is not None:
        if centre_id.shape != (B,):
            raise ValueError("centre_id must be [B]")
        centre_id = centre_id.to(device=dev, dtype=torch.int32).contiguous()
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        cid = None
        if centre_id is not None:
            # owners are addressed inside the chunk: the first instance of the chunk with the same owner takes the role
            cid = centre_id if (s == 0 and e == B) else shared_centre_ids(centre_id[s:e])
        rc = lib.mc_mincurv_solve_batch_shared(e - s, n_max, _ptr(n_pts[s:e]) if n_pts is not None else None,
                                               _ptr(reftrack[s:e]), _ptr(normvec[s:e]), _ptr(h[s:e]), float(kappa_bound),
                                               w_scalar, _ptr(w_batch[s:e]) if w_batch is not None else None,
                                               float(F_SCALE if f_scale is None else f_scale), _ptr(cid),
                                               _ptr(alpha[s:e]), _ptr(cerr[s:e]), _ptr(kmax[s:e]), _ptr(status[s:e]),
                                               _ptr(iters[s:e]), _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "mc_mincurv_solve_batch_shared")
    return dict(alpha=alpha, curv_error_max=cerr, kappa_lin_max=kmax, status=status, iters=iters)


def shared_centre_ids(group: torch.Tensor) -> torch.Tensor:
    """centre_id for opt_min_curv_batch from group labels [B] (any integers: instances with equal labels share a
    centreline): the first instance of every group becomes its owner.  Returns int32 [B] on the labels' device."""
    vals, inv = torch.unique(group, return_inverse=True)
    idx = torch.arange(group.shape[0], device=group.device, dtype=torch.int64)
    first = torch.full((vals.shape[0],), group.shape[0], device=group.device, dtype=torch.int64)
    first = first.scatter_reduce(0, inv, idx, reduce="amin")
    return first[inv].to(torch.int32)


@_device_guard
def opt_shortest_path_batch(reftrack: torch.Tensor, normvec: torch.Tensor, w_veh: Union[float, torch.Tensor],
                            n_pts: Optional[torch.Tensor] = None) -> dict:
    """Batched tph.opt_shortest_path.  Returns dict(alpha, status, iters)."""
    _require_cuda()
    lib = _lib.load()
    reftrack = _f64(reftrack, "reftrack")
    normvec = _f64(normvec, "normvec")
    B, n_max, four = reftrack.shape
    if four != 4 or normvec.shape != (B, n_max, 2):
        raise RuntimeError("Array size of reftrack should be the same as normvectors!")
    dev = reftrack.device
    n_pts = _npts(n_pts, B, dev)
    w_scalar, w_batch = _wveh(w_veh, B, dev)
    alpha = torch.empty((B, n_max), dtype=torch.float64, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    iters = torch.empty((B,), dtype=torch.int32, device=dev)
    ws = _workspace("shortest", lib.mc_shortest_path_workspace_bytes(B, n_max), dev)
    rc = lib.mc_shortest_path_solve_batch(B, n_max, _ptr(n_pts), _ptr(reftrack), _ptr(normvec), w_scalar, _ptr(w_batch),
                                          _ptr(alpha), _ptr(status), _ptr(iters), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "mc_shortest_path_solve_batch")
    return dict(alpha=alpha, status=status, iters=iters)


def _closed_polygon_length(pts: torch.Tensor, n_pts: Optional[torch.Tensor], normvec: Optional[torch.Tensor] = None,
                           shift: Optional[torch.Tensor] = None, shift_stride: int = 1, sign: float = 1.0) -> torch.Tensor:
    """Length [B] of the closed polygon through the first n_pts[b] points p_i (+ sign * shift_i * n_i) of every row
    (mc_polygon_length_batch: a block reduction per track on the device)."""
    lib = _lib.load()
    B, n_max, stride = pts.shape
    out = torch.zeros((B,), dtype=torch.float64, device=pts.device)
    rc = lib.mc_polygon_length_batch(B, n_max, _ptr(n_pts), _ptr(pts), stride, _ptr(normvec), _ptr(shift), int(shift_stride),
                                     float(sign), _ptr(out), _stream())
    _lib.check(rc, "mc_polygon_length_batch")
    return out


@_device_guard
def create_raceline_batch(refline: torch.Tensor, normvec: torch.Tensor, alpha: torch.Tensor, stepsize_interp: float,
                          n_pts: Optional[torch.Tensor] = None, n_out_max: Optional[int] = None,
                          with_head_curv: bool = True) -> dict:
    """Batched tph.create_raceline (+ tph.calc_head_curv_an at the resampled points).

    refline: [B, n_max, 2] or a reftrack [B, n_max, 4].  If n_out_max is None an upper bound is derived from
    the polygon length of the shifted line (one small device->host read) and the call is repeated with a larger one if
    a track still does not fit; with an explicit n_out_max an overflow is reported as n_out[b] = -(points needed)."""
    _require_cuda()
    lib = _lib.load()
    refline = _f64(refline, "refline")
    normvec = _f64(normvec, "normvec")
    alpha = _f64(alpha, "alpha")
    B, n_max, stride = refline.shape
    dev = refline.device
    n_pts = _npts(n_pts, B, dev)
    derived = n_out_max is None
    if derived:
        poly = _closed_polygon_length(refline, n_pts, normvec=normvec, shift=alpha)
        n_out_max = int(math.ceil(float(poly.max().item()) * 1.1 / float(stepsize_interp))) + 16
    n_out_max = int(n_out_max)
    while True:
        out = _create_raceline_once(lib, refline, normvec, alpha, stepsize_interp, n_pts, n_out_max, with_head_curv)
        if not derived:           # the caller fixed the capacity: an overflow is reported as n_out[b] = -(points needed)
            return out
        need = int((-out["n_out"]).max().item())
        if need <= 0:
            return out
        n_out_max = need + 16     # (the spline is longer than 1.1 x its polygon: re-run with what the kernel asked for)


def _create_raceline_once(lib, refline, normvec, alpha, stepsize_interp, n_pts, n_out_max, with_head_curv):
    B, n_max, stride = refline.shape
    dev = refline.device
    f64 = dict(dtype=torch.float64, device=dev)
    out = dict(
        coeffs_x=torch.zeros((B, n_max, 4), **f64), coeffs_y=torch.zeros((B, n_max, 4), **f64),
        spline_lengths=torch.zeros((B, n_max), **f64), n_out=torch.zeros((B,), dtype=torch.int32, device=dev),
        raceline_interp=torch.zeros((B, n_out_max, 2), **f64),
        spline_inds=torch.zeros((B, n_out_max), dtype=torch.int32, device=dev),
        t_values=torch.zeros((B, n_out_max), **f64), s_interp=torch.zeros((B, n_out_max), **f64),
        el_lengths_interp=torch.zeros((B, n_out_max), **f64),
        psi=torch.zeros((B, n_out_max), **f64) if with_head_curv else None,
        kappa=torch.zeros((B, n_out_max), **f64) if with_head_curv else None,
    )
    ws = _workspace("splines", lib.mc_create_raceline_workspace_bytes(B, n_max), dev)
    rc = lib.mc_create_raceline_batch(B, n_max, _ptr(n_pts), _ptr(refline), stride, _ptr(normvec), _ptr(alpha),
                                      float(stepsize_interp), n_out_max, _ptr(out["coeffs_x"]), _ptr(out["coeffs_y"]),
                                      _ptr(out["spline_lengths"]), _ptr(out["n_out"]), _ptr(out["raceline_interp"]),
                                      _ptr(out["spline_inds"]), _ptr(out["t_values"]), _ptr(out["s_interp"]),
                                      _ptr(out["el_lengths_interp"]), _ptr(out["psi"]), _ptr(out["kappa"]),
                                      _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "mc_create_raceline_batch")
    return out


@_device_guard
def calc_head_curv_batch(coeffs_x: torch.Tensor, coeffs_y: torch.Tensor, ind_spls: torch.Tensor, t_spls: torch.Tensor,
                         n_eval: Optional[torch.Tensor] = None, calc_curv: bool = True, calc_dcurv: bool = False):
    _require_cuda()
    lib = _lib.load()
    if not calc_curv and calc_dcurv:
        raise ValueError("dkappa cannot be calculated without kappa!")
    coeffs_x = _f64(coeffs_x, "coeffs_x")
    coeffs_y = _f64(coeffs_y, "coeffs_y")
    t_spls = _f64(t_spls, "t_spls")
    B, n_max, _ = coeffs_x.shape
    dev = coeffs_x.device
    ind = ind_spls.to(device=dev, dtype=torch.int32).contiguous()
    n_eval_max = t_spls.shape[1]
    psi = torch.empty((B, n_eval_max), dtype=torch.float64, device=dev)
    kappa = torch.empty_like(psi) if calc_curv else None
    dkappa = torch.empty_like(psi) if calc_dcurv else None
    rc = lib.mc_calc_head_curv_batch(B, n_max, _ptr(coeffs_x), _ptr(coeffs_y), n_eval_max,
                                     _ptr(_npts(n_eval, B, dev)), _ptr(ind), _ptr(t_spls), _ptr(psi), _ptr(kappa),
                                     _ptr(dkappa), _stream())
    _lib.check(rc, "mc_calc_head_curv_batch")
    return psi, kappa, dkappa


@_device_guard
def iqp_relinearise_batch(reftrack, normvec, alpha, stepsize_interp, n_pts=None, active=None, n_max_new=None):
    """One re-linearisation step of tph.iqp_handler for every (active) track: returns
    (reftrack_new [B, n_max_new, 4], normvec_new [B, n_max_new, 2], n_pts_new [B])."""
    _require_cuda()
    lib = _lib.load()
    reftrack = _f64(reftrack, "reftrack")
    normvec = _f64(normvec, "normvec")
    alpha = _f64(alpha, "alpha")
    B, n_max, _ = reftrack.shape
    dev = reftrack.device
    n_pts = _npts(n_pts, B, dev)
    if n_max_new is None:
        n_max_new = n_max + 64
    if active is not None:
        active = active.to(device=dev, dtype=torch.int32).contiguous()
    rnew = torch.zeros((B, n_max_new, 4), dtype=torch.float64, device=dev)
    nnew = torch.zeros((B, n_max_new, 2), dtype=torch.float64, device=dev)
    npn = torch.zeros((B,), dtype=torch.int32, device=dev)
    ws = _workspace("iqp", lib.mc_iqp_relinearise_workspace_bytes(B, n_max, n_max_new), dev)
    rc = lib.mc_iqp_relinearise_batch(B, n_max, _ptr(n_pts), _ptr(active), _ptr(reftrack), _ptr(normvec), _ptr(alpha),
                                      float(stepsize_interp), int(n_max_new), _ptr(rnew), _ptr(nnew), _ptr(npn),
                                      _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "mc_iqp_relinearise_batch")
    return rnew, nnew, npn


@_device_guard
def scale_alpha_batch(alpha: torch.Tensor, scale: Union[float, torch.Tensor]) -> None:
    lib = _lib.load()
    B, n_max = alpha.shape
    sb = scale.to(device=alpha.device, dtype=torch.float64).contiguous() if isinstance(scale, torch.Tensor) else None
    rc = lib.mc_scale_alpha_batch(B, n_max, _ptr(alpha), _ptr(sb), 1.0 if sb is not None else float(scale), _stream())
    _lib.check(rc, "mc_scale_alpha_batch")


@_device_guard
def iqp_batch(reftrack: torch.Tensor, normvec: torch.Tensor, h: torch.Tensor, kappa_bound: float,
              w_veh: Union[float, torch.Tensor], stepsize_interp: float, iters_min: int = 3,
              curv_error_allowed: float = 0.01, n_pts: Optional[torch.Tensor] = None, max_iters: int = 50,
              fixed_iters: Optional[int] = None, _n_cap_min: int = 0) -> dict:
    """Batched tph.iqp_handler: per-instance outer iterations with damping and
    re-linearisation; an instance leaves the loop once iter >= iters_min and its curv_error_max <=
    curv_error_allowed.  Returns dict(alpha [B, n_cap], reftrack [B, n_cap, 4], normvec [B, n_cap, 2],
    n_pts [B], outer_iters [B], status [B], qp_solves) -- every array refers to the instance's LAST iteration.
    status: the QP status of that iteration, or 2 if max_iters outer iterations did not reach curv_error_allowed (tph
    would keep iterating).  One small device->host read per outer iteration.

    fixed_iters: run exactly that many outer iterations for every instance (bench config C3).  Nothing has to be decided
    on the host between the iterations then, so they are queued back to back and the one thing the host would have
    looked at -- a re-sampled track that does not fit the capacity -- is checked once at the end (and the call repeated
    with the larger capacity, which the 5 % + 50 m of slack makes a rare event)."""
    _require_cuda()
    reftrack0, normvec0, h0 = reftrack, normvec, h           # (the caller's arrays: never written)
    reftrack = _f64(reftrack, "reftrack").clone()
    normvec = _f64(normvec, "normvec").clone()
    h = _f64(h, "h").clone()
    B, n_max, _ = reftrack.shape
    dev = reftrack.device
    cur_n = (n_pts.to(device=dev, dtype=torch.int32).clone() if n_pts is not None
             else torch.full((B,), n_max, dtype=torch.int32, device=dev))
    # capacity of the re-sampled tracks: the raceline is re-sampled every stepsize_interp metres, so its point count
    # follows from its length, not from n_max (a track given at a coarser spacing grows); 5 % + 50 m of slack on the
    # reference polygon, and the relinearisation step below grows the buffers if a track still does not fit
    poly = float(_closed_polygon_length(reftrack, cur_n).max().item())
    n_cap = max(n_max + 64, int(math.ceil((1.05 * poly + 50.0) / float(stepsize_interp))) + 16, int(_n_cap_min))

    def _alloc(cap):
        return dict(alpha=torch.zeros((B, cap), dtype=torch.float64, device=dev),
                    reftrack=torch.zeros((B, cap, 4), dtype=torch.float64, device=dev),
                    normvec=torch.zeros((B, cap, 2), dtype=torch.float64, device=dev))

    lib = _lib.load()
    fin = _alloc(n_cap)
    fin.update(n_pts=torch.zeros((B,), dtype=torch.int32, device=dev),
               outer_iters=torch.zeros((B,), dtype=torch.int32, device=dev),
               status=torch.zeros((B,), dtype=torch.int32, device=dev),
               curv_error_max=torch.zeros((B,), dtype=torch.float64, device=dev))
    active = torch.ones((B,), dtype=torch.int32, device=dev)
    counters = torch.zeros((2,), dtype=torch.int32, device=dev)
    n_active = B
    qp_solves = 0
    it = 0
    limit = fixed_iters if fixed_iters is not None else max_iters
    worst = None              # fixed_iters: smallest new point count seen so far (device scalar; negative = overflow)
    while True:
        it += 1
        n_cur_max = reftrack.shape[1]
        res = opt_min_curv_batch(reftrack, normvec, h, kappa_bound, w_veh, n_pts=cur_n)     # (finished tracks: n_pts = 0, skipped)
        qp_solves += n_active
        alpha = res["alpha"]
        if it < iters_min:
            scale_alpha_batch(alpha, it * 1.0 / iters_min)
        # per-track termination and the copy of finished tracks into the result buffers: on the device
        rc = lib.mc_iqp_finish_batch(B, n_cur_max, n_cap, it, int(iters_min), float(curv_error_allowed),
                                     int(fixed_iters) if fixed_iters is not None else 0, int(limit), _ptr(active),
                                     _ptr(res["status"]), _ptr(res["curv_error_max"]), _ptr(cur_n), _ptr(alpha), _ptr(reftrack),
                                     _ptr(normvec), _ptr(fin["alpha"]), _ptr(fin["reftrack"]), _ptr(fin["normvec"]),
                                     _ptr(fin["n_pts"]), _ptr(fin["outer_iters"]), _ptr(fin["status"]),
                                     _ptr(fin["curv_error_max"]), _ptr(counters), _stream())
        _lib.check(rc, "mc_iqp_finish_batch")
        if it >= limit:                       # every track was finished by this call (fixed count or cap): nothing to read back
            break
        if fixed_iters is not None:
            rt_new, nv_new, n_new = iqp_relinearise_batch(reftrack, normvec, alpha, stepsize_interp, n_pts=cur_n,
                                                          active=active, n_max_new=n_cap)
            worst = n_new.min() if worst is None else torch.minimum(worst, n_new.min())
            reftrack, normvec, cur_n = rt_new, nv_new, n_new
            h = torch.ones((B, n_cap), dtype=torch.float64, device=dev)
            continue
        while True:
            rt_new, nv_new, n_new = iqp_relinearise_batch(reftrack, normvec, alpha, stepsize_interp, n_pts=cur_n,
                                                          active=active, n_max_new=n_cap)
            # the ONE host read of the iteration: tracks still active, and the smallest new point count (negative = -(points
            # needed) for a track that does not fit the capacity)
            n_active, n_min = (int(v) for v in torch.stack((counters[0], n_new.min())).tolist())
            if n_active == 0 or n_min >= 0:
                break
            n_cap = -n_min + 64
            grown = _alloc(n_cap)
            for k in ("alpha", "reftrack", "normvec"):
                grown[k][:, :fin[k].shape[1]] = fin[k]
                fin[k] = grown[k]
        if n_active == 0:
            break
        reftrack, normvec, cur_n = rt_new, nv_new, n_new
        h = torch.ones((B, n_cap), dtype=torch.float64, device=dev)   # use_dist_scaling=False from iteration 2 on
    if worst is not None:
        n_min = int(worst.item())
        if n_min < 0:         # a re-sampled track did not fit: its later iterations ran on nothing -- repeat with room for it
            return iqp_batch(reftrack0, normvec0, h0, kappa_bound, w_veh, stepsize_interp, iters_min=iters_min,
                             curv_error_allowed=curv_error_allowed, n_pts=n_pts, max_iters=max_iters,
                             fixed_iters=fixed_iters, _n_cap_min=-n_min + 64)
    fin["qp_solves"] = qp_solves
    return fin


# ------------------------------------------------------------------------------------------------
# velocity-profile stage: tph.calc_vel_profile + calc_ax_profile + calc_t_profile
# ------------------------------------------------------------------------------------------------
def _table(t, cols: int, name: str, dev) -> torch.Tensor:
    t = torch.as_tensor(t, dtype=torch.float64).to(dev).contiguous()
    if t.ndim != 2 or t.shape[1] != cols:
        raise RuntimeError({3: "ggv diagram must consist of the three columns [vx, ax_max, ay_max]!",
                            2: "ax_max_machines must consist of the two columns [vx, ax_max_machines]!"}[cols])
    if t.shape[0] > 256:
        raise ValueError(f"{name}: at most 256 rows are supported")
    return t


@_device_guard
def vel_profile_batch(kappa: torch.Tensor, el_lengths: torch.Tensor, ggv, ax_max_machines, v_max,
                      drag_coeff: float, m_veh: float, dyn_model_exp: float = 1.0, filt_window: Optional[int] = None,
                      mu: Optional[torch.Tensor] = None, n_pts: Optional[torch.Tensor] = None,
                      ggv_scales=None, want_profiles: bool = True, max_chunk: Optional[int] = None,
                      decel_slice_upper: Optional[int] = None) -> dict:
    """Batched tph.calc_vel_profile (closed, ggv branch) + calc_ax_profile + calc_t_profile.

    kappa, el_lengths: [B, n_max] device tensors (n_pts[b] valid entries), e.g. the ``kappa`` /
    ``el_lengths_interp`` / ``n_out`` results of create_raceline_batch.  ``v_max`` is a float, or -- together with
    ``ggv_scales`` -- a sequence of V per-variant values: variant v of every track runs with
    ggv[:, 1:] * ggv_scales[v] and top speed v_max[v] (one cell of the reference's lap-time matrix,
    /root/reference/main_globaltraj.py:442-496).  Returns dict(laptime [B, V], status [B, V] and, if want_profiles,
    vx [B, V, n_max], ax [B, V, n_max], t [B, V, n_max + 1])."""
    _require_cuda()
    lib = _lib.load()
    kappa = _f64(kappa, "kappa")
    el_lengths = _f64(el_lengths, "el_lengths")
    B, n_max = kappa.shape
    if el_lengths.shape != (B, n_max):
        raise RuntimeError("kappa and el_lengths must have the same length if closed!")
    dev = kappa.device
    if mu is not None:
        mu = _f64(mu, "mu")
        if mu.shape != (B, n_max):
            raise RuntimeError("kappa and mu must have the same length!")
    n_pts = _npts(n_pts, B, dev)
    ggv_t = _table(ggv, 3, "ggv", dev)
    mach_t = _table(ax_max_machines, 2, "ax_max_machines", dev)
    if filt_window is not None and int(filt_window) % 2 != 1:
        raise RuntimeError("Window width of moving average filter must be odd!")
    fw = 0 if filt_window is None else int(filt_window)
    vmax_t = scale_t = None
    per_variant = ggv_scales is not None or isinstance(v_max, torch.Tensor) or hasattr(v_max, "__len__")
    if per_variant:
        vm = torch.as_tensor(v_max, dtype=torch.float64).reshape(-1)
        sc = torch.ones_like(vm) if ggv_scales is None else torch.as_tensor(ggv_scales, dtype=torch.float64).reshape(-1)
        if vm.numel() == 1 and sc.numel() > 1:
            vm = vm.expand(sc.numel()).clone()
        if sc.numel() != vm.numel():
            raise ValueError("v_max and ggv_scales must have one entry per variant")
        V = int(vm.numel())
        vmax_t, scale_t = vm.to(dev).contiguous(), sc.to(dev).contiguous()
        v_hi, v_scalar = float(vm.max().item()), 0.0
    else:
        V, v_hi, v_scalar = 1, float(v_max), float(v_max)
    # tph's range checks (the tables must cover the whole velocity range of the car)
    if float(mach_t[-1, 0].item()) < v_hi:
        raise RuntimeError("ax_max_machines has to cover the entire velocity range of the car (i.e. >= v_max)!")
    if float(ggv_t[-1, 0].item()) < v_hi:
        raise RuntimeError("ggv has to cover the entire velocity range of the car (i.e. >= v_max)!")
    f64 = dict(dtype=torch.float64, device=dev)
    laptime = torch.zeros((B, V), **f64)
    status = torch.zeros((B, V), dtype=torch.int32, device=dev)
    vx = torch.zeros((B, V, n_max), **f64) if want_profiles else None
    ax = torch.zeros((B, V, n_max), **f64) if want_profiles else None
    t = torch.zeros((B, V, n_max + 1), **f64) if want_profiles else None
    per_track = lib.mc_vel_profile_workspace_bytes(1, V, n_max)
    chunk = _chunk(B, per_track, dev) if max_chunk is None else min(B, int(max_chunk))
    chunk = max(1, min(chunk, (2 ** 31 - 1024) // V))
    ws = _workspace("velprofile", lib.mc_vel_profile_workspace_bytes(chunk, V, n_max), dev)
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        rc = lib.mc_vel_profile_batch_ex(e - s, n_max, _ptr(n_pts[s:e]) if n_pts is not None else None, _ptr(kappa[s:e]),
                                         _ptr(el_lengths[s:e]), _ptr(mu[s:e]) if mu is not None else None, V, _ptr(scale_t),
                                         _ptr(vmax_t), v_scalar, int(ggv_t.shape[0]), _ptr(ggv_t), int(mach_t.shape[0]),
                                         _ptr(mach_t), float(dyn_model_exp), float(drag_coeff), float(m_veh), fw,
                                         int(VP_DECEL_SLICE_UPPER if decel_slice_upper is None else decel_slice_upper),
                                         _ptr(vx[s:e]) if vx is not None else None, _ptr(ax[s:e]) if ax is not None else None,
                                         _ptr(t[s:e]) if t is not None else None, _ptr(laptime[s:e]), _ptr(status[s:e]),
                                         _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "mc_vel_profile_batch_ex")
    out = dict(laptime=laptime, status=status)
    if want_profiles:
        out.update(vx=vx, ax=ax, t=t)
    return out


def lap_time_matrix_batch(kappa: torch.Tensor, el_lengths: torch.Tensor, ggv, ax_max_machines, ggv_scales, top_speeds,
                          drag_coeff: float, m_veh: float, dyn_model_exp: float = 1.0,
                          filt_window: Optional[int] = None, n_pts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The lap-time matrix of /root/reference/main_globaltraj.py:442-496 for every track of the batch in one launch:
    returns [B, len(top_speeds), len(ggv_scales)] lap times (top speeds in m/s)."""
    ts = torch.as_tensor(top_speeds, dtype=torch.float64).reshape(-1)
    gs = torch.as_tensor(ggv_scales, dtype=torch.float64).reshape(-1)
    vm = ts.repeat_interleave(gs.numel())
    sc = gs.repeat(ts.numel())
    res = vel_profile_batch(kappa, el_lengths, ggv, ax_max_machines, vm, drag_coeff, m_veh, dyn_model_exp, filt_window,
                            n_pts=n_pts, ggv_scales=sc, want_profiles=False)
    bad_ = res["status"] != 0
    if bool(bad_.any().item()):
        raise RuntimeError("lap_time_matrix_batch: non-finite lap time for %i profile(s)" % int(bad_.sum().item()))
    return res["laptime"].reshape(kappa.shape[0], ts.numel(), gs.numel())


@_device_guard
def calc_ax_t_profile_batch(vx: torch.Tensor, el_lengths: torch.Tensor, ax_in: Optional[torch.Tensor] = None,
                            t_start: float = 0.0, n_pts: Optional[torch.Tensor] = None, want_t: bool = True):
    """Stand-alone tph.calc_ax_profile (ax_in None: vx holds n + 1 values per row) / tph.calc_t_profile.
    Returns (ax [P, n_max], t [P, n_max + 1] or None)."""
    _require_cuda()
    lib = _lib.load()
    vx = _f64(vx, "vx")
    el_lengths = _f64(el_lengths, "el_lengths")
    P, n_max = el_lengths.shape
    dev = vx.device
    if ax_in is not None:
        ax_in = _f64(ax_in, "ax_in")
    ax_out = torch.zeros((P, n_max), dtype=torch.float64, device=dev)
    t_out = torch.zeros((P, n_max + 1), dtype=torch.float64, device=dev) if want_t else None
    rc = lib.mc_calc_ax_t_profile_batch(P, n_max, _ptr(_npts(n_pts, P, dev)), _ptr(vx), int(vx.shape[1]), _ptr(el_lengths),
                                        _ptr(ax_in), float(t_start), _ptr(ax_out), _ptr(t_out), _stream())
    _lib.check(rc, "mc_calc_ax_t_profile_batch")
    return ax_out, t_out


# ------------------------------------------------------------------------------------------------
# trajectory back end: the reference's in-tree helpers interp_track, calc_min_bound_dists,
# check_traj and the trajectory assembly of main_globaltraj.py:501-512, batched
# ------------------------------------------------------------------------------------------------
@_device_guard
def interp_track_batch(pts: torch.Tensor, stepsize_approx: float = 1.0, n_pts: Optional[torch.Tensor] = None,
                       normvec: Optional[torch.Tensor] = None, normal_sign: float = 1.0, width_col: int = 2,
                       n_out_max: Optional[int] = None):
    """Batched helper_funcs_glob.src.interp_track.interp_track (/root/reference/helper_funcs_glob/src/interp_track.py).

    pts: [B, n_max, 2 or 4].  With ``normvec`` the re-sampled polyline is pts.xy + normal_sign * normvec * pts[..., width_col]
    (the track boundaries of check_traj.py:50-61).  Returns (out [B, n_out_max, 4], n_out [B])."""
    _require_cuda()
    lib = _lib.load()
    pts = _f64(pts, "pts")
    B, n_max, stride = pts.shape
    dev = pts.device
    n_pts = _npts(n_pts, B, dev)
    if normvec is not None:
        normvec = _f64(normvec, "normvec")
        if normvec.shape != (B, n_max, 2) or stride != 4:
            raise ValueError("normvec needs a [B, n_max, 4] track and must be [B, n_max, 2]")
    if n_out_max is None:
        if normvec is None:
            poly = _closed_polygon_length(pts, n_pts)
        else:       # the boundary polyline p + sign * w n: the width column of the track is the per-point shift (stride 4)
            poly = _closed_polygon_length(pts, n_pts, normvec=normvec, shift=pts.reshape(-1)[int(width_col):].view(-1),
                                          shift_stride=4, sign=float(normal_sign))
        n_out_max = int(math.ceil(float(poly.max().item()) / float(stepsize_approx))) + 8
    ws = _workspace("interp_track", lib.mc_interp_track_workspace_bytes(B, n_max), dev)
    while True:
        out = torch.zeros((B, int(n_out_max), 4), dtype=torch.float64, device=dev)
        n_out = torch.zeros((B,), dtype=torch.int32, device=dev)
        rc = lib.mc_interp_track_batch(B, n_max, _ptr(n_pts), _ptr(pts), stride, _ptr(normvec), float(normal_sign),
                                       int(width_col), float(stepsize_approx), int(n_out_max), _ptr(out), _ptr(n_out),
                                       _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "mc_interp_track_batch")
        need = int((-n_out).max().item())
        if need <= 0:
            return out, n_out
        n_out_max = need + 8


@_device_guard
def min_bound_dists_batch(xy: torch.Tensor, psi: torch.Tensor, bound1: torch.Tensor, bound2: torch.Tensor,
                          length_veh: float, width_veh: float, n_traj: Optional[torch.Tensor] = None,
                          nb1: Optional[torch.Tensor] = None, nb2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched helper_funcs_glob.src.calc_min_bound_dists.calc_min_bound_dists: [B, n_traj_max] minimum distances of
    the vehicle corners to the boundary points (bound1/bound2: [B, nb_max, >= 2], x and y first)."""
    _require_cuda()
    lib = _lib.load()
    xy, psi, bound1, bound2 = _f64(xy, "xy"), _f64(psi, "psi"), _f64(bound1, "bound1"), _f64(bound2, "bound2")
    B, n_traj_max, _ = xy.shape
    if bound1.shape[2] != bound2.shape[2]:
        raise ValueError("bound1 and bound2 must have the same row layout")
    dev = xy.device
    out = torch.zeros((B, n_traj_max), dtype=torch.float64, device=dev)
    n_traj, nb1, nb2 = _npts(n_traj, B, dev), _npts(nb1, B, dev), _npts(nb2, B, dev)
    sl = lambda t, s, e: _ptr(t[s:e]) if t is not None else None
    for s in range(0, B, _GRID_Y_MAX):           # the track index is the y dimension of the launch grid
        e = min(B, s + _GRID_Y_MAX)
        rc = lib.mc_min_bound_dists_batch(e - s, n_traj_max, sl(n_traj, s, e), _ptr(xy[s:e]), _ptr(psi[s:e]),
                                          int(bound1.shape[1]), sl(nb1, s, e), _ptr(bound1[s:e]), int(bound2.shape[1]),
                                          sl(nb2, s, e), _ptr(bound2[s:e]), int(bound1.shape[2]), float(length_veh),
                                          float(width_veh), _ptr(out[s:e]), _stream())
        _lib.check(rc, "mc_min_bound_dists_batch")
    return out


EXTREMA = ("min_dist", "kappa_abs_max", "ay_max", "ax_wo_drag_max", "ax_wo_drag_min", "a_tot_max", "vx_max", "n_points")


@_device_guard
def traj_extrema_batch(kappa: torch.Tensor, vx: torch.Tensor, ax: torch.Tensor, dragcoeff: float, mass_veh: float,
                       min_dists: Optional[torch.Tensor] = None, n_traj: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, 8] extrema per trajectory, columns as in EXTREMA (min_dist = inf without min_dists)."""
    _require_cuda()
    lib = _lib.load()
    kappa, vx, ax = _f64(kappa, "kappa"), _f64(vx, "vx"), _f64(ax, "ax")
    B, n_max = kappa.shape
    dev = kappa.device
    if min_dists is not None:
        min_dists = _f64(min_dists, "min_dists")
    ext = torch.zeros((B, 8), dtype=torch.float64, device=dev)
    rc = lib.mc_traj_extrema_batch(B, n_max, _ptr(_npts(n_traj, B, dev)), _ptr(kappa), _ptr(vx), _ptr(ax), _ptr(min_dists),
                                   float(dragcoeff), float(mass_veh), _ptr(ext), _stream())
    _lib.check(rc, "mc_traj_extrema_batch")
    return ext


@_device_guard
def check_traj_batch(reftrack: torch.Tensor, normvec: torch.Tensor, xy: torch.Tensor, psi: torch.Tensor,
                     kappa: torch.Tensor, vx: torch.Tensor, ax: torch.Tensor, length_veh: float, width_veh: float,
                     dragcoeff: float, mass_veh: float, n_pts: Optional[torch.Tensor] = None,
                     n_traj: Optional[torch.Tensor] = None, bound_stepsize: float = 1.0) -> dict:
    """The quantities helper_funcs_glob.src.check_traj.check_traj tests, for a batch of trajectories: boundaries
    re-sampled every ``bound_stepsize`` metres, minimum distance of the vehicle corners to them for every trajectory
    point (against ALL boundary points -- the reference passes only the first point of each boundary, check_traj.py:58-61,
    which the single-track mirror reproduces), and the extrema of curvature, accelerations and speed.
    Returns dict(min_dists [B, n_traj_max], bound_r/bound_l [B, nb, 4] with nb_r/nb_l [B], and one [B] tensor per name
    in EXTREMA)."""
    _require_cuda()
    kappa = _f64(kappa, "kappa")
    B = kappa.shape[0]
    n_traj = _npts(n_traj, B, kappa.device)
    br, nbr = interp_track_batch(reftrack, bound_stepsize, n_pts=n_pts, normvec=normvec, normal_sign=1.0, width_col=2)
    bl, nbl = interp_track_batch(reftrack, bound_stepsize, n_pts=n_pts, normvec=normvec, normal_sign=-1.0, width_col=3)
    md = min_bound_dists_batch(xy, psi, br, bl, length_veh, width_veh, n_traj=n_traj, nb1=nbr, nb2=nbl)
    ext = traj_extrema_batch(kappa, vx, ax, dragcoeff, mass_veh, min_dists=md, n_traj=n_traj)
    out = dict(min_dists=md, bound_r=br, bound_l=bl, nb_r=nbr, nb_l=nbl)
    out.update({name: ext[:, i] for i, name in enumerate(EXTREMA)})
    return out


def check_traj_flags(chk: dict, ggv, ax_max_machines, v_max: float, curvlim: float, min_dist_warn: float = 1.0) -> dict:
    """The comparisons of check_traj.py:74-139 as boolean [B] tensors (True = the reference would print the warning)."""
    import numpy as _np
    f = dict(min_dist=chk["min_dist"] < min_dist_warn, curvature=chk["kappa_abs_max"] > curvlim,
             v_max=chk["vx_max"] > v_max + 0.1)
    if ggv is not None:
        g = _np.asarray(ggv, dtype=float)
        f.update(ay=chk["ay_max"] > float(g[:, 2].max()) + 0.1, ax_pos=chk["ax_wo_drag_max"] > float(g[:, 1].max()) + 0.1,
                 ax_neg=chk["ax_wo_drag_min"] < float((-g[:, 1]).min()) - 0.1, a_tot=chk["a_tot_max"] > float(g[:, 1:].max()) + 0.1)
    if ax_max_machines is not None:
        m = _np.asarray(ax_max_machines, dtype=float)
        f["ax_machines"] = chk["ax_wo_drag_max"] > float(m[:, 1].max()) + 0.1
    return f


@_device_guard
def assemble_trajectory_batch(s: torch.Tensor, xy: torch.Tensor, psi: torch.Tensor, kappa: torch.Tensor, vx: torch.Tensor,
                              ax: torch.Tensor, spline_lengths: torch.Tensor, n_traj: Optional[torch.Tensor] = None,
                              n_spl: Optional[torch.Tensor] = None) -> torch.Tensor:
    """trajectory_opt / traj_race_cl of /root/reference/main_globaltraj.py:501-512 for a batch: [B, n_max + 1, 7] rows
    [s, x, y, psi, kappa, vx, ax]; row n_traj[b] closes the lap with s = sum(spline_lengths[b])."""
    _require_cuda()
    lib = _lib.load()
    s, xy, psi, kappa, vx, ax = (_f64(t, nm) for t, nm in ((s, "s"), (xy, "xy"), (psi, "psi"), (kappa, "kappa"), (vx, "vx"),
                                                          (ax, "ax")))
    spline_lengths = _f64(spline_lengths, "spline_lengths")
    B, n_max = s.shape
    dev = s.device
    traj = torch.zeros((B, n_max + 1, 7), dtype=torch.float64, device=dev)
    rc = lib.mc_assemble_trajectory_batch(B, n_max, _ptr(_npts(n_traj, B, dev)), _ptr(s), _ptr(xy), _ptr(psi), _ptr(kappa),
                                          _ptr(vx), _ptr(ax), int(spline_lengths.shape[1]), _ptr(_npts(n_spl, B, dev)),
                                          _ptr(spline_lengths), _ptr(traj), _stream())
    _lib.check(rc, "mc_assemble_trajectory_batch")
    return traj


@_device_guard
def check_normals_crossing_batch(track: torch.Tensor, normvec: torch.Tensor, horizon: int = 10,
                                 n_pts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched tph.check_normals_crossing (/root/reference/helper_funcs_glob/src/prep_track.py:57-59): bool [B], True where
    two normals at most ``horizon`` points apart cross inside the track."""
    _require_cuda()
    lib = _lib.load()
    track, normvec = _f64(track, "track"), _f64(normvec, "normvec")
    B, n_max, four = track.shape
    if four != 4 or normvec.shape != (B, n_max, 2):
        raise ValueError("track must be [B, n_max, 4] and normvec [B, n_max, 2]")
    dev = track.device
    n_pts = _npts(n_pts, B, dev)
    smallest = n_max if n_pts is None else int(n_pts.min().item())
    if horizon >= smallest:
        raise RuntimeError("Horizon of %i points is too large for a track with %i points, reduce horizon!" % (horizon, smallest))
    crossing = torch.zeros((B,), dtype=torch.int32, device=dev)
    for s in range(0, B, _GRID_Y_MAX):
        e = min(B, s + _GRID_Y_MAX)
        rc = lib.mc_check_normals_crossing_batch(e - s, n_max, _ptr(n_pts[s:e]) if n_pts is not None else None,
                                                 _ptr(track[s:e]), _ptr(normvec[s:e]), int(horizon), _ptr(crossing[s:e]),
                                                 _stream())
        _lib.check(rc, "mc_check_normals_crossing_batch")
    return crossing != 0


# ------------------------------------------------------------------------------------------------
# sweep inputs generated on the device: width-jitter variants from seeds
# ------------------------------------------------------------------------------------------------
@_device_guard
def jitter_widths_batch(base: torch.Tensor, seeds: torch.Tensor, rel: float = 0.1, centre_id: Optional[torch.Tensor] = None,
                        n_pts_base: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """V = len(seeds) width-jitter variants of the prepared tracks ``base`` [n_base, n_max, 4] (variant v uses track
    centre_id[v], default v % n_base): w <- w (1 + rel g(s)), g smooth with |g| <= 1 drawn from seeds[v] (int64) by the
    stateless hash that synth.jitter_widths_hash mirrors on the host.  Returns (tracks [V, n_max, 4], n_pts [V])."""
    _require_cuda()
    lib = _lib.load()
    base = _f64(base, "base")
    n_base, n_max, four = base.shape
    if four != 4:
        raise ValueError("base must be [n_base, n_max, 4]")
    dev = base.device
    seeds = seeds.to(device=dev, dtype=torch.int64).contiguous()
    V = int(seeds.numel())
    if centre_id is not None:
        centre_id = centre_id.to(device=dev, dtype=torch.int32).contiguous()
        if centre_id.numel() != V:
            raise ValueError("centre_id must have one entry per variant")
    if out is None:
        out = torch.empty((V, n_max, 4), dtype=torch.float64, device=dev)
    n_out = torch.empty((V,), dtype=torch.int32, device=dev)
    rc = lib.mc_jitter_widths_batch(V, n_max, _ptr(_npts(n_pts_base, n_base, dev)), n_base, _ptr(base), _ptr(centre_id),
                                    _ptr(seeds), float(rel), _ptr(out), _ptr(n_out), _stream())
    _lib.check(rc, "mc_jitter_widths_batch")
    return out, n_out


# ------------------------------------------------------------------------------------------------
# prep_track front end: tph.spline_approximation + min-width inflation + splines + normals check
# ------------------------------------------------------------------------------------------------
@_device_guard
def spline_approximation_batch(track: torch.Tensor, k_reg: int = 3, s_reg: float = 10.0, stepsize_prep: float = 1.0,
                               stepsize_reg: float = 3.0, n_raw: Optional[torch.Tensor] = None,
                               min_width: Optional[float] = None):
    """Batched tph.spline_approximation (+ prep_track's min-width inflation) for imported tracks [B, n_raw_max, 4]
    (unclosed; n_raw[b] points each).  Returns (reftrack_interp [B, n_out_max, 4], n_pts [B], smoothing_lambda [B]).
    The smoothing spline is the Reinsch formulation with residual budget s_reg (csrc/prep_track.cu)."""
    _require_cuda()
    lib = _lib.load()
    track = _f64(track, "track")
    B, n_raw_max, four = track.shape
    if four != 4:
        raise ValueError("track must be [B, n_raw_max, 4]")
    dev = track.device
    n_raw = _npts(n_raw, B, dev)
    length = float(_closed_polygon_length(track, n_raw).max().item())
    n_int_max = int(math.ceil(length / float(stepsize_prep))) + 8
    n_out_max = int(math.ceil(1.05 * length / float(stepsize_reg))) + 16
    while True:
        out = torch.zeros((B, n_out_max, 4), dtype=torch.float64, device=dev)
        n_out = torch.zeros((B,), dtype=torch.int32, device=dev)
        lam = torch.zeros((B,), dtype=torch.float64, device=dev)
        ws = _workspace("prep_track", lib.mc_prep_track_workspace_bytes(B, n_raw_max, n_int_max), dev)
        rc = lib.mc_prep_track_batch(B, n_raw_max, _ptr(n_raw), _ptr(track), int(k_reg), float(s_reg), float(stepsize_prep),
                                     float(stepsize_reg), float(min_width) if min_width is not None else 0.0, n_int_max,
                                     n_out_max, _ptr(out), _ptr(n_out), _ptr(lam), _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "mc_prep_track_batch")
        need = int((-n_out).max().item())
        if need <= 0:
            return out, n_out, lam
        if need + 16 <= min(n_out_max, n_int_max):     # the kernel refused for another reason than capacity
            raise ValueError("spline_approximation_batch: a track is too short to be smoothed (fewer than 5 points after "
                             "pre-interpolation)")
        n_out_max = max(n_out_max, need + 16)          # (a curve longer than 1.05 x its polygon, or a tiny capacity)
        n_int_max = max(n_int_max, need + 16)


def prep_track_batch(track: torch.Tensor, reg_smooth_opts: dict, stepsize_opts: dict, n_raw: Optional[torch.Tensor] = None,
                     min_width: Optional[float] = None, check_normals: bool = True) -> dict:
    """Batched helper_funcs_glob.src.prep_track.prep_track (/root/reference/helper_funcs_glob/src/prep_track.py): smoothing
    and re-sampling, closed splines of the result, check of the normals, min-width inflation.  Returns dict(reftrack_interp,
    n_pts, normvec_normalized_interp, h (the spline system in moment form), coeffs_x_interp, coeffs_y_interp,
    normals_crossing [B] bool)."""
    rt, n_pts, lam = spline_approximation_batch(track, k_reg=reg_smooth_opts["k_reg"], s_reg=reg_smooth_opts["s_reg"],
                                                stepsize_prep=stepsize_opts["stepsize_prep"],
                                                stepsize_reg=stepsize_opts["stepsize_reg"], n_raw=n_raw, min_width=None)
    cx, cy, nv, h = calc_splines_batch(rt, n_pts=n_pts)
    crossing = check_normals_crossing_batch(rt, nv, 10, n_pts=n_pts) if check_normals else None
    if min_width is not None:                     # (the reference inflates AFTER the normals check: prep_track.py:89-98)
        rt, n_pts, lam = spline_approximation_batch(track, k_reg=reg_smooth_opts["k_reg"], s_reg=reg_smooth_opts["s_reg"],
                                                    stepsize_prep=stepsize_opts["stepsize_prep"],
                                                    stepsize_reg=stepsize_opts["stepsize_reg"], n_raw=n_raw, min_width=min_width)
    return dict(reftrack_interp=rt, n_pts=n_pts, normvec_normalized_interp=nv, h=h, coeffs_x_interp=cx, coeffs_y_interp=cy,
                normals_crossing=crossing, smoothing_lambda=lam)
