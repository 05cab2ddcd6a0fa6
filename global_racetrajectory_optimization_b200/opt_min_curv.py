"""Drop-in module for ``trajectory_planning_helpers.opt_min_curv`` (see tph_api.py for the reference call sites)."""
from .tph_api import opt_min_curv  # noqa: F401
