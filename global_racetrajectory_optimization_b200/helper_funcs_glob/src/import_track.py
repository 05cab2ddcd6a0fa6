"""import_track -- call site /root/reference/main_globaltraj.py:193-195 (host-side CSV reader, no device work)."""
import numpy as np


def import_track(file_path: str, imp_opts: dict, width_veh: float) -> np.ndarray:
    """CSV with 3 ([x, y, w_total]), 4 ([x, y, w_right, w_left]) or 5 ([x, y, z, w_right, w_left]) columns ->
    [x, y, w_tr_right, w_tr_left], repeated ``num_laps`` times, optionally reversed and rolled to a new start point."""
    raw = np.loadtxt(file_path, comments="#", delimiter=",")
    ncol = raw.shape[1]
    if ncol == 3:
        widths = np.column_stack((raw[:, 2] / 2, raw[:, 2] / 2))
    elif ncol in (4, 5):
        widths = raw[:, ncol - 2:ncol]
    else:
        raise IOError("Track file cannot be read!")
    track = np.tile(np.column_stack((raw[:, :2], widths)), (imp_opts["num_laps"], 1))
    if imp_opts["flip_imp_track"]:
        track = track[::-1].copy()
    if imp_opts["set_new_start"]:
        d2 = (track[:, 0] - imp_opts["new_start"][0]) ** 2 + (track[:, 1] - imp_opts["new_start"][1]) ** 2
        track = np.roll(track, -int(np.argmin(d2)), axis=0)
    narrowest = float(np.amin(track[:, 2] + track[:, 3]))
    if narrowest < width_veh + 0.5:
        print("WARNING: Minimum track width %.2fm is close to or smaller than vehicle width!" % narrowest)
    return track
