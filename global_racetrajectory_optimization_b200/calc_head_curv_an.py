"""Drop-in module for ``trajectory_planning_helpers.calc_head_curv_an`` (see tph_api.py for the reference call sites)."""
from .tph_api import calc_head_curv_an  # noqa: F401
