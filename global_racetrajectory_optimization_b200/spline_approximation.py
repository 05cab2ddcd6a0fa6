"""Drop-in module for ``trajectory_planning_helpers.spline_approximation`` (see tph_api.py for the reference call sites)."""
from .tph_api import spline_approximation  # noqa: F401
