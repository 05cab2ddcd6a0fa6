"""prep_track -- call site /root/reference/main_globaltraj.py:252-257: the stage between import_track and the optimisation."""
import sys

import numpy as np

from ... import batch as _b
from ...spline_system import SplineSystem
from ._dev import up


def prep_track(reftrack_imp: np.ndarray, reg_smooth_opts: dict, stepsize_opts: dict, debug: bool = True,
               min_width: float = None) -> tuple:
    """Smoothing / re-sampling of the imported track, closed splines, check of the spline normals, optional min-width
    inflation.  Returns (reftrack_interp, normvec_normalized_interp, a_interp, coeffs_x_interp, coeffs_y_interp) like the
    reference; ``a_interp`` is the spline system in moment form (SplineSystem: np.asarray() gives the dense matrix)."""
    track = np.asarray(reftrack_imp, dtype=np.float64)
    res = _b.prep_track_batch(up(track), reg_smooth_opts, stepsize_opts, min_width=min_width)
    n = int(res["n_pts"][0].item())
    if bool(res["normals_crossing"][0].item()):
        raise IOError("At least two spline normals are crossed, check input or increase smoothing factor!")
    rt = res["reftrack_interp"][0, :n].cpu().numpy()
    if min_width is not None:
        rt_plain = _b.spline_approximation_batch(up(track), k_reg=reg_smooth_opts["k_reg"], s_reg=reg_smooth_opts["s_reg"],
                                                 stepsize_prep=stepsize_opts["stepsize_prep"],
                                                 stepsize_reg=stepsize_opts["stepsize_reg"])[0][0, :n].cpu().numpy()
        if np.any(rt[:, 2:] != rt_plain[:, 2:]):
            print("WARNING: Track region was smaller than requested minimum track width -> Applied artificial inflation in"
                  " order to match the requirements!", file=sys.stderr)
    a_interp = SplineSystem(res["h"][0, :n].cpu().numpy())
    return (rt, res["normvec_normalized_interp"][0, :n].cpu().numpy(), a_interp,
            res["coeffs_x_interp"][0, :n].cpu().numpy(), res["coeffs_y_interp"][0, :n].cpu().numpy())
