"""interp_track -- call sites /root/reference/helper_funcs_glob/src/prep_track.py:32-34, check_traj.py:58-61."""
import numpy as np

from ... import batch as _b
from ._dev import up


def interp_track(reftrack: np.ndarray, stepsize_approx: float = 1.0) -> np.ndarray:
    """Equidistant linear re-sampling of the closed track [x, y, w_tr_right, w_tr_left] -> unclosed [m, 4] array."""
    track = np.asarray(reftrack, dtype=np.float64)
    if track.ndim != 2 or track.shape[1] != 4:
        raise ValueError("interp_track expects an [n, 4] array [x, y, w_tr_right, w_tr_left]")
    out, n_out = _b.interp_track_batch(up(track), float(stepsize_approx))
    return out[0, :int(n_out[0].item())].cpu().numpy()
