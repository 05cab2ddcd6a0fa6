"""Drop-in module for ``trajectory_planning_helpers.opt_shortest_path`` (see tph_api.py for the reference call sites)."""
from .tph_api import opt_shortest_path  # noqa: F401
