import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
GOLD_REAL = os.path.join(ROOT, "tests", "golden_real")      # written by tools/pin_against_tph.py --write (real tph + quadprog)


def _apply_pin():
    """If the pinning kit has been run, every parity test runs against what the REAL packages produced: the two
    constants go into the oracle and into the run-time parameters of the C-ABI, and load_golden() prefers the fixtures
    regenerated from the real packages."""
    pin_file = os.path.join(GOLD_REAL, "pin.json")
    if not os.path.exists(pin_file):
        return None
    import json
    pin = json.load(open(pin_file))
    from oracle import tph_dense, tph_velprofile
    tph_dense.F_SCALE = float(pin["f_scale"]["chosen"])
    tph_velprofile.DECEL_LAP_SLICE_UPPER = bool(pin["decel_slice_upper"]["chosen"])
    try:
        from global_racetrajectory_optimization_b200 import batch
        batch.F_SCALE = float(pin["f_scale"]["chosen"])
        batch.VP_DECEL_SLICE_UPPER = int(pin["decel_slice_upper"]["chosen"])
    except Exception:
        pass
    return pin


PIN = _apply_pin()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a machine without a CUDA device, whatever -m says."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False))
    real = os.path.join(GOLD_REAL, name + ".npz")
    if PIN is not None and os.path.exists(real):
        g.update(dict(np.load(real, allow_pickle=False)))        # same keys, produced by the real packages
    return g


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


def rel_max(a, b):
    """max |a - b| / max |b|  (the parity norm of SURVEY.md 8c: alpha crosses zero, so elementwise
    relative error is meaningless)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
