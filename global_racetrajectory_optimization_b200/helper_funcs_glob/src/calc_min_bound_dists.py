"""calc_min_bound_dists -- call site /root/reference/helper_funcs_glob/src/check_traj.py:64-68."""
import numpy as np

from ... import batch as _b
from ._dev import up


def calc_min_bound_dists(trajectory: np.ndarray, bound1: np.ndarray, bound2: np.ndarray, length_veh: float,
                         width_veh: float) -> np.ndarray:
    """Smallest distance between the four vehicle corners and any boundary point, per trajectory row
    (columns used: x = 1, y = 2, psi = 3).  bound1/bound2 may be [k, >= 2] arrays or single points."""
    traj = np.asarray(trajectory, dtype=np.float64)
    sides = []
    for bound in (bound1, bound2):
        pts = np.atleast_2d(np.asarray(bound, dtype=np.float64))[:, :2]
        sides.append(up(pts))
    dists = _b.min_bound_dists_batch(up(traj[:, 1:3]), up(traj[:, 3]), sides[0], sides[1], float(length_veh), float(width_veh))
    return dists[0].cpu().numpy()
