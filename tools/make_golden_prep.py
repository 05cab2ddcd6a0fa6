"""Generates tests/golden/prep_track.npz -- fixtures of the prep_track front end:

    <track>_reinsch : oracle/tph_prep.spline_approximation_reinsch (the algorithm of csrc/prep_track.cu, dense numpy)
    <track>_fitpack : the scipy/FITPACK route (tools/make_golden.spline_approximation: scipy restatement of
                      tph.spline_approximation) -- what the reference's own prep_track produces up to the scipy version

for the reference's four tracks (raw CSV bytes travel in tests/golden/refback_import_track.npz).

    python tools/make_golden_prep.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import tph_prep as P  # noqa: E402
import make_golden as mg  # noqa: E402
import pin_against_tph as kit  # noqa: E402

if __name__ == "__main__":
    raws = kit.raw_tracks()
    out = {}
    for name in ("rounded_rectangle", "handling_track", "berlin_2018", "modena_2019"):
        t0 = time.time()
        a = P.spline_approximation_reinsch(raws[name])
        b = mg.spline_approximation(raws[name])
        out[name + "_reinsch"], out[name + "_fitpack"] = a, b
        m = min(a.shape[0], b.shape[0])
        print(f"{name}: raw {raws[name].shape[0]} pts -> {a.shape[0]} (reinsch) / {b.shape[0]} (fitpack); "
              f"xy distance over the first {m} points {np.abs(a[:m, :2] - b[:m, :2]).max():.3f} m ({time.time() - t0:.0f} s)", flush=True)
    out["rounded_rectangle_minwidth6"] = P.spline_approximation_reinsch(raws["rounded_rectangle"], min_width=6.0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "prep_track.npz"), **out)
