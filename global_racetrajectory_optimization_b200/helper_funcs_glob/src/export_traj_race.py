"""export_traj_race -- call site /root/reference/main_globaltraj.py:539-541; file format of
/root/reference/helper_funcs_glob/src/export_traj_race.py:21-41 (2 comment lines: UUID, SHA1(ggv file); 7 columns)."""
import numpy as np

from ._export_common import write_csv

COLUMNS = ("s_m", "x_m", "y_m", "psi_rad", "kappa_radpm", "vx_mps", "ax_mps2")


def export_traj_race(file_paths: dict, traj_race: np.ndarray) -> None:
    write_csv(file_paths["traj_race_export"], file_paths.get("ggv_file"), np.asarray(traj_race, dtype=np.float64), COLUMNS)
