"""Drop-in module for ``trajectory_planning_helpers.calc_ax_profile`` (see tph_api.py for the reference call sites)."""
from .tph_api import calc_ax_profile  # noqa: F401
