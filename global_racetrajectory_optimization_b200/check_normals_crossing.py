"""Drop-in module for ``trajectory_planning_helpers.check_normals_crossing`` (see tph_api.py for the reference call sites)."""
from .tph_api import check_normals_crossing  # noqa: F401
