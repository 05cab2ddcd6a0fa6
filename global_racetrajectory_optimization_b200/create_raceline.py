"""Drop-in module for ``trajectory_planning_helpers.create_raceline`` (see tph_api.py for the reference call sites)."""
from .tph_api import create_raceline  # noqa: F401
