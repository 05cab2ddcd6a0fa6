from . import (calc_min_bound_dists, check_traj, export_traj_ltpl, export_traj_race, import_track,  # noqa: F401
               interp_track, prep_track)
