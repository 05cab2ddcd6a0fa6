"""Drop-in module for ``trajectory_planning_helpers.calc_t_profile`` (see tph_api.py for the reference call sites)."""
from .tph_api import calc_t_profile  # noqa: F401
