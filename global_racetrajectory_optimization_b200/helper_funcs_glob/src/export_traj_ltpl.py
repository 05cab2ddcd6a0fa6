"""export_traj_ltpl -- call site /root/reference/main_globaltraj.py:544-551; file format of
/root/reference/helper_funcs_glob/src/export_traj_ltpl.py:36-86 (12 columns, one row per reference point)."""
import numpy as np

from ._export_common import write_csv

COLUMNS = ("x_ref_m", "y_ref_m", "width_right_m", "width_left_m", "x_normvec_m", "y_normvec_m", "alpha_m", "s_racetraj_m",
           "psi_racetraj_rad", "kappa_racetraj_radpm", "vx_racetraj_mps", "ax_racetraj_mps2")


def export_traj_ltpl(file_paths: dict, spline_lengths_opt, trajectory_opt, reftrack, normvec_normalized, alpha_opt) -> None:
    traj = np.asarray(trajectory_opt, dtype=np.float64)
    # arc length of the raceline at every reference point = start of its spline
    s_ref = np.concatenate(([0.0], np.cumsum(spline_lengths_opt)))[:-1]
    # psi, kappa, vx, ax are taken from the trajectory row whose s is closest to that arc length (first one on ties)
    nearest = np.abs(traj[:, 0][None, :] - s_ref[:, None]).argmin(axis=1)
    table = np.column_stack((reftrack, normvec_normalized, alpha_opt, s_ref, traj[nearest, 3:7]))
    # (the reference also builds a closed copy of this table but writes the unclosed one: export_traj_ltpl.py:62-86)
    write_csv(file_paths["traj_ltpl_export"], file_paths.get("ggv_file"), table, COLUMNS)
