"""Drop-in module for ``trajectory_planning_helpers.calc_splines`` (see tph_api.py for the reference call sites)."""
from .tph_api import calc_splines  # noqa: F401
