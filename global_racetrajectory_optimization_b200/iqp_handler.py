"""Drop-in module for ``trajectory_planning_helpers.iqp_handler`` (see tph_api.py for the reference call sites)."""
from .tph_api import iqp_handler  # noqa: F401
