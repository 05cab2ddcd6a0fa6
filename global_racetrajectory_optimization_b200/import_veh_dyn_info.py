"""Drop-in module for ``trajectory_planning_helpers.import_veh_dyn_info`` (see tph_api.py for the reference call sites)."""
from .tph_api import import_veh_dyn_info  # noqa: F401
