"""Device-backed stand-in for the reference's in-tree package ``helper_funcs_glob`` (back-end half: track import,
re-sampling, trajectory checks, exports).  ``import global_racetrajectory_optimization_b200.helper_funcs_glob as
helper_funcs_glob`` keeps the call sites of /root/reference/main_globaltraj.py:520-553 unchanged."""
from . import src  # noqa: F401
