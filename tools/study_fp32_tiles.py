"""Numerics study for DESIGN.md section 9, option (2): store the block-Cholesky factor tiles of the interior-point
solve in fp32 (half the HBM and shared-memory traffic of the triangular sweeps) and let the interior-point iteration,
whose residuals stay fp64, absorb the inexact Newton directions.

CPU emulation (numpy): the block-cyclic factorisation of round 1's csrc/mincurv_pdip.cu (replaced by csrc/mincurv_ipm.cu in round 2) (32 x 32 chain blocks, 32-row
separator, explicit Linv tiles, T and F tiles) computed in fp64, then every stored tile rounded to fp32 before the
sweeps use it; Mehrotra predictor-corrector on the banded QP of the fixtures.  Reports iterations and the error of alpha
against the all-fp64 run and against the dense oracle.

    python tools/study_fp32_tiles.py [fixture ...]
"""
import os
import sys

import numpy as np
from scipy.linalg import solve_triangular

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tph_dense as T  # noqa: E402

B = 32


class BlockCyc:
    def __init__(self, H, round_tiles):
        self.n = H.shape[0]
        self.H = H
        self.na = self.n - B
        self.nb = (self.na + B - 1) // B
        self.rt = (lambda a: a.astype(np.float32).astype(np.float64)) if round_tiles else (lambda a: a)
        # padded chain index -> node (identity rows for the padding), separator = last 32 nodes
        self.idx = [np.arange(B * i, min(B * i + B, self.na)) for i in range(self.nb)]
        self.sep = np.arange(self.na, self.n)

    def _blk(self, r, c, M):
        out = np.zeros((B, B))
        out[:len(r), :len(c)] = M[np.ix_(r, c)]
        return out

    def factor(self, D):
        M = self.H + np.diag(D)
        self.Linv, self.Tm, self.F = [], [None], []
        S = M[np.ix_(self.sep, self.sep)].copy()
        Tprev = Fprev = None
        for i in range(self.nb):
            r = self.idx[i]
            A = self._blk(r, r, M)
            for k in range(len(r), B):
                A[k, k] = 1.0
            if i > 0:
                A = A - Tprev @ Tprev.T
            L = np.linalg.cholesky(A)
            Li = solve_triangular(L, np.eye(B), lower=True)
            Y = self._blk(self.sep, r, M)
            FW = Y - (Fprev @ Tprev.T if i > 0 else 0.0)
            F = FW @ Li.T
            S = S - F @ F.T                        # the separator update uses the fp64 F (it is formed before the store)
            if i < self.nb - 1:
                Bn = self._blk(self.idx[i + 1], r, M)
                Tn = Bn @ Li.T
                self.Tm.append(self.rt(Tn))
                Tprev = Tn
            Fprev = F
            self.Linv.append(self.rt(Li))
            self.F.append(self.rt(F))
        Ls = np.linalg.cholesky(S)
        self.LinvS = solve_triangular(Ls, np.eye(B), lower=True)       # separator factor stays on chip: fp64

    def solve(self, g):
        nb = self.nb
        y = np.zeros((nb, B))
        gs = g[self.sep].copy()
        for i in range(nb):
            t = np.zeros(B)
            t[:len(self.idx[i])] = g[self.idx[i]]
            if i > 0:
                t -= self.Tm[i] @ y[i - 1]
            y[i] = self.Linv[i] @ t
            gs -= self.F[i] @ y[i]
        xs = self.LinvS.T @ (self.LinvS @ gs)
        x = np.zeros((nb, B))
        for i in range(nb - 1, -1, -1):
            t = y[i] - self.F[i].T @ xs
            if i < nb - 1:
                t -= self.Tm[i + 1].T @ x[i + 1]
            x[i] = self.Linv[i].T @ t
        out = np.zeros(self.n)
        for i in range(nb):
            out[self.idx[i]] = x[i][:len(self.idx[i])]
        out[self.sep] = xs
        return out


def ipm(H, f, lb, ub, round_tiles, maxit=60):
    """Mehrotra predictor-corrector as in the kernel (centre start, fraction to the boundary 0.995, termination on mu,
    dual residual and the size of the last step)."""
    n = f.size
    bc = BlockCyc(H, round_tiles)
    a = 0.5 * (lb + ub)
    su, sl = ub - a, a - lb
    g = H @ a + f
    lu = np.maximum(-g, 0) + 1e-2 * np.abs(g).max() + 1e-8
    ll = np.maximum(g, 0) + 1e-2 * np.abs(g).max() + 1e-8
    mu0 = (su @ lu + sl @ ll) / (2 * n)
    rd_tol = 1e-8 * (np.abs(f).max() + np.abs(g).max())

    def maxstep(v, dv):
        m = dv < 0
        return min(1.0, (-v[m] / dv[m]).min()) if m.any() else 1.0
    for it in range(1, maxit + 1):
        rd = H @ a + f + lu - ll
        mu = (su @ lu + sl @ ll) / (2 * n)
        bc.factor(lu / su + ll / sl)
        da = bc.solve(-rd + lu - ll)
        dlu, dll = (-su * lu + lu * da) / su, (-sl * ll - ll * da) / sl
        ap = min(maxstep(su, -da), maxstep(sl, da))
        ad = min(maxstep(lu, dlu), maxstep(ll, dll))
        mu_aff = ((su - ap * da) @ (lu + ad * dlu) + (sl + ap * da) @ (ll + ad * dll)) / (2 * n)
        sigma = (mu_aff / mu) ** 3
        tu, tl = sigma * mu - su * lu + da * dlu, sigma * mu - sl * ll - da * dll
        da = bc.solve(-rd - tu / su + tl / sl)
        dlu, dll = (tu + lu * da) / su, (tl - ll * da) / sl
        ap = min(1.0, 0.995 * min(maxstep(su, -da), maxstep(sl, da)))
        ad = min(1.0, 0.995 * min(maxstep(lu, dlu), maxstep(ll, dll)))
        a = a + ap * da
        su, sl = su - ap * da, sl + ap * da
        lu, ll = lu + ad * dlu, ll + ad * dll
        mu = (su @ lu + sl @ ll) / (2 * n)
        rdn = np.abs(H @ a + f + lu - ll).max()
        if mu <= 1e-10 * mu0 and rdn <= rd_tol and np.abs(ap * da).max() <= 1e-5 * np.abs(a).max():
            return a, it, True
    return a, maxit, False


def band(H, b=32):
    n = H.shape[0]
    i, j = np.indices(H.shape)
    d = np.minimum((i - j) % n, (j - i) % n)
    return np.where(d <= b, H, 0.0)


def main():
    names = sys.argv[1:] or ["synth200", "synth333", "handling", "synth500", "synth500_narrow", "berlin", "modena"]
    print("%-16s %5s | %-22s | %-22s | %s" % ("fixture", "N", "fp64 tiles: it, err", "fp32 tiles: it, err", "fp32 vs fp64"))
    for name in names:
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        rt, nv = g["reftrack"], g["normvec"]
        _, _, A, _ = T.calc_splines(np.vstack((rt[:, :2], rt[0, :2])))
        qp = T.assemble_min_curv(rt, nv, A, float(g["kappa_bound"]), float(g["w_veh"]))
        n = rt.shape[0]
        H, f = band(qp["H"]), qp["f"]
        ub, lb = qp["h"][:n], -qp["h"][n:2 * n]
        ref = g["alpha_mincurv_boxonly"]
        scale = np.abs(ref).max()
        a64, it64, ok64 = ipm(H, f, lb, ub, False)
        a32, it32, ok32 = ipm(H, f, lb, ub, True)
        print("%-16s %5d | %2d%s %.2e          | %2d%s %.2e          | %.2e" % (
            name, n, it64, " " if ok64 else "!", np.abs(a64 - ref).max() / scale, it32, " " if ok32 else "!",
            np.abs(a32 - ref).max() / scale, np.abs(a32 - a64).max() / scale))


if __name__ == "__main__":
    main()
