"""CPU tests (-m "not gpu") of the velocity-profile stage:

* the oracle (oracle/tph_velprofile.py, PARITY UNPINNED like tph_dense.py) against the committed golden vectors and
  against physical properties of the forward/backward solver (ggv limits respected, lap-time monotonicity);
* the arithmetic the CUDA kernel runs (csrc/vel_profile_core.cuh), compiled for the host by tests/host_harness/
  and compared statement-for-statement with the oracle -- for both readings of the one detail of tph's closed-track
  solver that cannot be confirmed offline (which half of the doubled lap survives the backward pass).
The kernel launch itself is covered by tests/test_gpu_velprofile.py on the GPU box."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import tph_velprofile as VP

DP = ctypes.POINTER(ctypes.c_double)
VEH = dict(drag_coeff=0.75, m_veh=1200.0)


def _p(a):
    return a.ctypes.data_as(DP) if a is not None else None


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    """{True/False: ctypes lib} -- vel_profile_core.cuh built for the host with VP_DECEL_SLICE_UPPER = 1 / 0."""
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    out = {}
    d = tmp_path_factory.mktemp("vp_host")
    src = os.path.join(ROOT, "tests", "host_harness", "vp_host.cpp")
    for upper in (True, False):
        so = str(d / f"libvp_host_{int(upper)}.so")
        subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-std=c++17",
                               f"-DVP_DECEL_SLICE_UPPER={int(upper)}", "-o", so, src])
        lib = ctypes.CDLL(so)
        lib.vp_host_profile.restype = ctypes.c_int
        lib.vp_host_profile.argtypes = [ctypes.c_int, DP, DP, DP, ctypes.c_double, ctypes.c_double, ctypes.c_int, DP,
                                        ctypes.c_int, DP, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                        ctypes.c_int, DP, DP, DP, DP]
        lib.vp_host_ax_t.restype = None
        lib.vp_host_ax_t.argtypes = [ctypes.c_int, DP, DP, DP, ctypes.c_double, DP, DP]
        lib.vp_host_interp.restype = ctypes.c_double
        lib.vp_host_interp.argtypes = [ctypes.c_double, ctypes.c_int, DP, DP, ctypes.c_double, ctypes.c_int]
        out[upper] = lib
    return out


def host_profile(lib, kappa, el, mu, ggv, mach, scale, v_max, exp=1.0, filt=0, stride=1):
    n = kappa.size
    kappa, el = np.ascontiguousarray(kappa), np.ascontiguousarray(el)
    mu = np.ascontiguousarray(mu) if mu is not None else None
    ggv, mach = np.ascontiguousarray(ggv), np.ascontiguousarray(mach)
    vx, ax, t, lap = np.zeros(n), np.zeros(n), np.zeros(n + 1), np.zeros(1)
    st = lib.vp_host_profile(n, _p(kappa), _p(el), _p(mu), scale, v_max, ggv.shape[0], _p(ggv), mach.shape[0], _p(mach),
                             exp, VEH["drag_coeff"], VEH["m_veh"], filt, stride, _p(vx), _p(ax), _p(t), _p(lap))
    return st, vx, ax, t, float(lap[0])


def oracle_profile(kappa, el, mu, ggv, mach, scale, v_max, exp=1.0, filt=None):
    g = ggv.copy()
    g[:, 1:] *= scale
    vx = VP.calc_vel_profile(ggv=g, ax_max_machines=mach, v_max=v_max, kappa=kappa, el_lengths=el, closed=True,
                             dyn_model_exp=exp, filt_window=filt, mu=mu, **VEH)
    ax = VP.calc_ax_profile(vx_profile=np.append(vx, vx[0]), el_lengths=el, eq_length_output=False)
    t = VP.calc_t_profile(vx_profile=vx, ax_profile=ax, el_lengths=el)
    return vx, ax, t


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["berlin", "handling", "modena", "synth333", "synth1000"])
def test_oracle_reproduces_velprofile_golden(golden, name):
    v, g = golden("velprofile"), golden(name)
    vx, ax, t = oracle_profile(g["rl_kappa"], g["rl_el_lengths"], None, v["ggv"], v["ax_max_machines"], 1.0,
                               float(v["v_max"]))
    assert np.array_equal(vx, v[name + "_vx"]) and np.array_equal(ax, v[name + "_ax"]) and np.array_equal(t, v[name + "_t"])


def test_oracle_profile_respects_the_ggv_and_top_speed(golden):
    v, g = golden("velprofile"), golden("berlin")
    ggv, mach = v["ggv"], v["ax_max_machines"]
    k, el = g["rl_kappa"], g["rl_el_lengths"]
    vx, ax, t = oracle_profile(k, el, None, ggv, mach, 1.0, 45.0)
    assert vx.max() <= 45.0 and vx.min() > 0.0 and np.all(np.diff(t) > 0.0)
    ay = vx ** 2 * np.abs(k)
    assert ay.max() <= 12.0 * (1.0 + 1e-12)                                  # lateral limit of the stock ggv
    drag = vx ** 2 * VEH["drag_coeff"] / VEH["m_veh"]
    assert np.all(ax[:-1] <= np.interp(vx[:-1], mach[:, 0], mach[:, 1]) - drag[:-1] + 1e-9)   # machine limit - drag
    assert ax.min() >= -(12.0 + drag.max()) - 0.75                           # braking limit (+ one-step overshoot)
    # a weaker car is never faster: smaller ggv scale or lower top speed => longer lap
    laps = [oracle_profile(k, el, None, ggv, mach, s, 70.0)[2][-1] for s in (1.0, 0.8, 0.5)]
    assert laps[0] < laps[1] < laps[2]
    assert oracle_profile(k, el, None, ggv, mach, 1.0, 30.0)[2][-1] > laps[0]


def test_oracle_input_checks(golden):
    v, g = golden("velprofile"), golden("handling")
    k, el = g["rl_kappa"], g["rl_el_lengths"]
    kw = dict(ax_max_machines=v["ax_max_machines"], kappa=k, el_lengths=el, closed=True, **VEH)
    with pytest.raises(RuntimeError, match="Either ggv or loc_gg"):
        VP.calc_vel_profile(**kw)
    with pytest.raises(RuntimeError, match="same length if closed"):
        VP.calc_vel_profile(ggv=v["ggv"], **{**kw, "el_lengths": el[:-1]})
    with pytest.raises(RuntimeError, match="entire velocity range"):
        VP.calc_vel_profile(ggv=v["ggv"], v_max=80.0, **kw)
    with pytest.raises(RuntimeError, match="must be odd"):
        VP.calc_vel_profile(ggv=v["ggv"], v_max=70.0, filt_window=4, **kw)
    with pytest.raises(RuntimeError, match="1 element bigger"):
        VP.calc_ax_profile(np.ones(5), np.ones(5))


# ------------------------------------------------------------------------------------------------
CASES = [  # (ggv scale, v_max, dyn_model_exp, filt_window, seam shift, with mu)
    (1.0, 70.0, 1.0, 0, 0, False),
    (0.6, 100.0 / 3.6, 1.0, 0, 0, False),
    (1.0, 70.0, 2.0, 0, 0, False),
    (0.8, 40.0, 1.5, 5, 0, False),
    (1.0, 70.0, 1.0, 0, 137, False),        # seam inside another part of the lap
    (1.0, 70.0, 1.0, 0, 400, False),        # Modena: seam inside a braking zone (the two readings differ here)
    (1.0, 55.0, 1.0, 3, 50, True),
]


@pytest.mark.parametrize("upper", [True, False])
@pytest.mark.parametrize("name", ["berlin", "handling", "modena", "synth333"])
def test_kernel_arithmetic_on_the_host_matches_the_oracle(golden, harness, monkeypatch, name, upper):
    v, g = golden("velprofile"), golden(name)
    monkeypatch.setattr(VP, "DECEL_LAP_SLICE_UPPER", upper)
    lib = harness[upper]
    rng = np.random.default_rng(5)
    for scale, v_max, exp, filt, shift, with_mu in CASES:
        k = np.roll(g["rl_kappa"], shift % g["rl_kappa"].size)
        el = np.roll(g["rl_el_lengths"], shift % g["rl_kappa"].size)
        mu = (0.8 + 0.3 * rng.random(k.size)) if with_mu else None
        vx, ax, t = oracle_profile(k, el, mu, v["ggv"], v["ax_max_machines"], scale, v_max, exp, filt or None)
        st, hvx, hax, ht, lap = host_profile(lib, k, el, mu, v["ggv"], v["ax_max_machines"], scale, v_max, exp, filt,
                                             stride=3)
        assert st == 0
        # identical statements, IEEE arithmetic without contraction on both sides: differences can only come from
        # libm pow (dyn_model_exp != 1), the summation order of np.mean / np.convolve, and branch ties
        assert np.abs(hvx - vx).max() <= 1e-11 * vx.max()
        assert np.abs(hax - ax).max() <= 1e-9 * np.abs(ax).max()
        assert np.abs(ht - t).max() <= 1e-11 * t[-1] and lap == ht[-1]


def test_both_readings_agree_unless_the_seam_is_in_a_braking_zone(golden, harness):
    v = golden("velprofile")
    g = golden("modena")
    k, el = g["rl_kappa"], g["rl_el_lengths"]
    a = host_profile(harness[True], k, el, None, v["ggv"], v["ax_max_machines"], 1.0, 70.0)
    b = host_profile(harness[False], k, el, None, v["ggv"], v["ax_max_machines"], 1.0, 70.0)
    assert np.array_equal(a[1], b[1])                      # stock start/finish line: identical profiles
    k2, el2 = np.roll(k, 400), np.roll(el, 400)
    a = host_profile(harness[True], k2, el2, None, v["ggv"], v["ax_max_machines"], 1.0, 70.0)
    b = host_profile(harness[False], k2, el2, None, v["ggv"], v["ax_max_machines"], 1.0, 70.0)
    assert abs(b[4] - 79.3903800448767) < 1e-9            # second-lap reading: invariant under the seam position
    assert a[2].min() < -100.0 < b[2].min()                # first-visited-lap reading: impossible braking at the seam


def test_host_interp_is_numpy_interp(harness):
    lib = harness[True]
    rng = np.random.default_rng(0)
    for _ in range(20):
        n = int(rng.integers(1, 40))
        xp = np.sort(rng.random(n) * 80.0)
        if n > 1 and np.any(np.diff(xp) == 0.0):
            continue
        fp = rng.random(n) * 15.0
        s = float(rng.choice([1.0, 0.35, 0.7]))
        xs = np.concatenate((rng.random(200) * 100.0 - 10.0, xp, [np.inf, -np.inf]))
        want = np.interp(xs, xp, fp * s)
        for hint in (0, max(n - 2, 0), max(n // 2 - 1, 0)):        # the segment hint never changes the result
            got = np.array([lib.vp_host_interp(float(x), n, _p(xp), _p(fp), s, hint) for x in xs])
            assert np.array_equal(got, want)
    assert np.isnan(lib.vp_host_interp(float("nan"), 3, _p(np.array([0.0, 1.0, 2.0])), _p(np.ones(3)), 1.0, 0))


def test_host_ax_t_profile_matches_the_oracle(harness):
    lib = harness[True]
    rng = np.random.default_rng(1)
    n = 300
    vx = 20.0 + 10.0 * rng.random(n + 1)
    vx[10] = vx[11]                                          # ax == 0 branch of calc_t_profile
    el = 1.5 + rng.random(n)
    ax_o = VP.calc_ax_profile(vx, el)
    t_o = VP.calc_t_profile(vx[:-1], el, t_start=2.5, ax_profile=ax_o)
    ax, t = np.zeros(n), np.zeros(n + 1)
    lib.vp_host_ax_t(n, _p(vx), _p(el), None, 2.5, _p(ax), _p(t))
    assert np.array_equal(ax, ax_o) and np.array_equal(t, t_o)
    t2 = np.zeros(n + 1)
    lib.vp_host_ax_t(n, _p(vx), _p(el), _p(ax_o), 0.0, None, _p(t2))
    assert np.array_equal(t2, VP.calc_t_profile(vx[:-1], el, ax_profile=ax_o))
