"""CPU ORACLE (test infrastructure, NOT product code) -- restatement of ``tph.check_normals_crossing`` (tph 0.76), the
validity check of the prepared track at /root/reference/helper_funcs_glob/src/prep_track.py:57-59.
PARITY UNPINNED like the other tph restatements (package not vendored, no reference tests); follows the published
algorithm: for every point the normals of the +-horizon neighbours are intersected with its own normal by a 2 x 2 linear
solve; a crossing counts if both intersection parameters lie inside the track widths."""
from __future__ import annotations

import numpy as np


def check_normals_crossing(track: np.ndarray, normvec_normalized: np.ndarray, horizon: int = 10) -> bool:
    no_points = track.shape[0]
    if horizon >= no_points:
        raise RuntimeError("Horizon of %i points is too large for a track with %i points, reduce horizon!"
                           % (horizon, no_points))
    elif horizon >= no_points / 2:
        print("WARNING: Horizon of %i points makes no sense for a track with %i points, reduce horizon!"
              % (horizon, no_points))
    les_mat = np.zeros((2, 2))
    idx_list = list(range(0, no_points))
    idx_list = idx_list[-horizon:] + idx_list + idx_list[:horizon]
    for idx in range(no_points):
        idx_neighbours = idx_list[idx:idx + 2 * horizon + 1]
        del idx_neighbours[horizon]
        idx_neighbours = np.array(idx_neighbours)
        # normals (almost) parallel to the current one cannot cross it
        # (tph: np.cross of 2-D vectors, written out because numpy >= 2.0 deprecates that form)
        n_cur, n_nb = normvec_normalized[idx], normvec_normalized[idx_neighbours]
        is_collinear_b = np.isclose(n_cur[0] * n_nb[:, 1] - n_cur[1] * n_nb[:, 0], 0.0)
        idx_neighbours_rel = idx_neighbours[np.nonzero(np.invert(is_collinear_b))[0]]
        for idx_comp in list(idx_neighbours_rel):
            # p_1 + lambda_1 n_1 = p_2 + lambda_2 n_2
            const = track[idx_comp, :2] - track[idx, :2]
            les_mat[:, 0] = normvec_normalized[idx]
            les_mat[:, 1] = -normvec_normalized[idx_comp]
            lambdas = np.linalg.solve(les_mat, const)
            if -track[idx, 3] <= lambdas[0] <= track[idx, 2] and -track[idx_comp, 3] <= lambdas[1] <= track[idx_comp, 2]:
                return True
    return False
