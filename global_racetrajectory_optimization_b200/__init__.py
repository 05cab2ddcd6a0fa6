"""global_racetrajectory_optimization_b200 -- B200-native batched minimum-curvature / shortest-path QP path.

``import global_racetrajectory_optimization_b200 as tph`` gives the call surface that
/root/reference/main_globaltraj.py uses on its mincurv / mincurv_iqp / shortest_path branches
(``tph.opt_min_curv.opt_min_curv(...)`` etc.); ``batch`` holds the batched device API.
"""
from . import (batch, calc_ax_profile, calc_head_curv_an, check_normals_crossing, calc_splines, calc_t_profile,  # noqa: F401
               calc_vel_profile, create_raceline, import_veh_dyn_info, iqp_handler, opt_min_curv,
               opt_shortest_path, spline_approximation, synth)
from . import globaltraj, helper_funcs_glob  # noqa: F401
from .spline_system import SplineSystem  # noqa: F401

__version__ = "0.1.0"
