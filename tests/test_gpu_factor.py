"""GPU unit test (-m gpu) of the linear algebra inside mincurv_pdip_kernel (csrc/mincurv_ipm.cu): the bordered LDL^T
factorisation of the cyclic band M = H + D and the two kinds of solve (forward sweep fused into the factorisation /
full sweeps), through the C-ABI debug entry mc_debug_factor_solve, against a dense numpy solve of the same matrix.
Sizes cover the smallest supported track, sizes that are not multiples of 8 or 32, and the BASELINE size N = 1000."""
import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from global_racetrajectory_optimization_b200 import _lib, batch as B_  # noqa: E402

HBW = 32


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _cyclic_band_spd(n, rng, diag_boost):
    """H = R^T R with R cyclic banded (half-bandwidth 16): SPD-semidefinite, cyclic half-bandwidth 32."""
    R = np.zeros((n, n))
    for d in range(-16, 17):
        R[np.arange(n), (np.arange(n) + d) % n] = rng.standard_normal(n) * (0.6 ** abs(d))
    H = R.T @ R
    D = diag_boost * rng.uniform(0.1, 10.0, n)
    return H, D


def _fill_slab(ws, b, lay, n, H, D, g1, g2):
    vecs = B_.SLAB_VECTORS
    base = b * lay["stride"]
    slab = ws[base:base + lay["stride"]]
    hb = np.zeros((lay["np"], B_.HB_PITCH))
    idx = np.arange(n)
    for d in range(HBW + 1):
        hb[:n, d] = H[idx, (idx + d) % n]
    slab[lay["o_hb"]:lay["o_hb"] + hb.size] = hb.ravel()
    for name, v in (("DD", D), ("RHS", g1), ("T0", g2)):
        o = vecs.index(name) * lay["np"]
        slab[o:o + n] = v


def _read(ws, b, lay, name, n):
    o = b * lay["stride"] + B_.SLAB_VECTORS.index(name) * lay["np"]
    return ws[o:o + n].copy()


@pytest.mark.parametrize("n_max,sizes", [(128, [80, 81, 97, 128, 120, 127]), (333, [333, 300, 201]), (1000, [1000, 999, 777])])
@pytest.mark.parametrize("boost", [1.0, 1e-6])
def test_factor_and_solves_match_dense(n_max, sizes, boost):
    lib = _lib.load()
    rng = np.random.default_rng(n_max + int(-np.log10(boost)))
    lay = B_.mincurv_slab_layout(n_max)
    nB = len(sizes)
    nbytes = lib.mc_mincurv_workspace_bytes(nB, n_max)
    host = np.zeros(nbytes // 8)
    ref = []
    for b, n in enumerate(sizes):
        H, D = _cyclic_band_spd(n, rng, boost)
        g1, g2 = rng.standard_normal(n), rng.standard_normal(n)
        _fill_slab(host, b, lay, n, H, D, g1, g2)
        M = H + np.diag(D)
        ref.append((np.linalg.solve(M, g1), np.linalg.solve(M, g2), np.linalg.cond(M)))
    ws = torch.from_numpy(host).cuda()
    n_pts = torch.tensor(sizes, dtype=torch.int32, device="cuda")
    status = torch.full((nB,), -7, dtype=torch.int32, device="cuda")
    rc = lib.mc_debug_factor_solve(nB, n_max, ctypes.c_void_p(n_pts.data_ptr()), ctypes.c_void_p(status.data_ptr()),
                                   ctypes.c_void_p(ws.data_ptr()), ws.numel() * 8,
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mc_debug_factor_solve")
    torch.cuda.synchronize()
    out = ws.cpu().numpy()
    assert status.cpu().tolist() == [0] * nB
    for b, n in enumerate(sizes):
        x1, x2, cond = ref[b]
        tol = 1e-13 * max(cond, 1e2)
        for name, x in (("DX", x1), ("T1", x2), ("T2", x2)):
            got = _read(out, b, lay, name, n)
            err = np.abs(got - x).max() / np.abs(x).max()
            assert err <= tol, f"n={n} {name}: rel err {err:.2e} (cond {cond:.1e})"
