/* CPU BASELINE / CHECKER (test infrastructure, NOT product code; only tests/ and bench.py's cpu_baseline legs may use it).
 *
 * The SAME algorithm the CUDA path runs, written for one host core in plain C, so that the benchmark can separate
 * "algorithm" (banded O(N b^2) instead of the reference's dense O((4N)^3)) from "B200" (SURVEY.md section 8d(ii),
 * "fair CPU banded baseline"):
 *   - closed cubic spline moments by the periodic tridiagonal LDL^T (what tph.calc_splines' 4N x 4N solve reduces to),
 *   - k_ref, f = f_scale E^T k_ref and the cyclic band (half-bandwidth 32) of H = E^T E from the banded rows of E
 *     (E = S_y Z N_y - S_x Z N_x, Z = Tri^-1 6 D2; entries of Tri^-1 from its decay ratios),
 *   - min 1/2 a^T H a + f^T a, lb <= a <= ub by the Mehrotra predictor-corrector iteration of csrc/mincurv_ipm.cu, each
 *     iteration one bordered band Cholesky of H + D (chain + 32-node separator that closes the cycle) and two solves.
 * It follows tph.opt_min_curv's box-only QP (/root/reference/main_globaltraj.py:264-271) and is checked
 * against oracle/tph_dense.py in tests/test_oracle.py.
 *
 *   gcc -O3 -march=native -shared -fPIC -o libbanded_cpu.so banded_cpu.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define HBW 32
#define BZ 40          /* half-width kept of the rows of E (entries decay ~0.27 per off-diagonal) */
#define WARM 64        /* warm-up of the periodic recurrences */

static inline int wrap(int i, int n) { i %= n; return i < 0 ? i + n : i; }

/* periodic tridiagonal T (diag, off[i] couples i and i+1): forward / backward pivots with warm-up */
static void tri_pivots(int n, const double *diag, const double *off, double *d, double *dl) {
    double prev = diag[wrap(-WARM, n)];
    for (int s = -WARM + 1; s < n; ++s) {
        int i = wrap(s, n);
        prev = diag[i] - off[wrap(i - 1, n)] * off[wrap(i - 1, n)] / prev;
        if (s >= 0) d[i] = prev;
    }
    double nxt = diag[wrap(n - 1 + WARM, n)];
    for (int s = n - 2 + WARM; s >= 0; --s) {
        int i = wrap(s, n);
        nxt = diag[i] - off[i] * off[i] / nxt;
        if (s < n) dl[i] = nxt;
    }
}
static void tri_solve(int n, const double *d, const double *off, const double *r, double *m, double *y) {
    double prev = 0.0;
    for (int s = -WARM; s < n; ++s) {
        int i = wrap(s, n), im = wrap(i - 1, n);
        prev = r[i] - off[im] / d[im] * prev;
        if (s >= 0) y[i] = prev;
    }
    double nxt = 0.0;
    for (int s = n - 1 + WARM; s >= 0; --s) {
        int i = wrap(s, n);
        nxt = (y[i] - off[i] * nxt) / d[i];      /* (warm-up part: y taken cyclically) */
        if (s < n) m[i] = nxt;
    }
}

/* bordered band Cholesky of the cyclic band M (row i: M[i][i..i+32] cyclic, diagonal includes D): chain 0..NA-1, separator NA..n-1 */
typedef struct { int n, NA; double *Lb, *F, *S; } Fact;      /* Lb[k][0..32]: column k of L (diag first); F[k][r]; S 32x32 */

static int factor(const double *MB, int n, Fact *f) {
    const int NA = n - 32;
    f->n = n; f->NA = NA;
    double *Lb = f->Lb, *F = f->F, *S = f->S;
    /* chain: copy lower band columns (column k: rows k..k+32, rows >= NA excluded) */
    for (int k = 0; k < NA; ++k)
        for (int d = 0; d <= HBW; ++d) Lb[k * 33 + d] = (k + d < NA) ? MB[k * 33 + d] : 0.0;
    /* Y = M[sep, chain] */
    memset(F, 0, sizeof(double) * (size_t)NA * 32);
    for (int r = 0; r < 32; ++r) {
        for (int k = 0; k <= r && k < NA; ++k) F[k * 32 + r] = MB[(NA + r) * 33 + (k + 32 - r)];          /* across the wrap */
        for (int k = NA + r - 32 < 0 ? 0 : NA + r - 32; k < NA; ++k) if (NA + r - k <= 32) F[k * 32 + r] = MB[k * 33 + (NA + r - k)];
    }
    for (int r = 0; r < 32; ++r)
        for (int c = 0; c < 32; ++c) { int lo = r < c ? r : c, dist = abs(r - c); S[r * 32 + c] = MB[(NA + lo) * 33 + dist]; }
    for (int k = 0; k < NA; ++k) {
        double *col = Lb + k * 33;
        if (!(col[0] > 0.0)) return 0;
        const double piv = sqrt(col[0]), ip = 1.0 / piv;
        col[0] = piv;
        for (int d = 1; d <= HBW; ++d) col[d] *= ip;
        double *fk = F + k * 32;
        for (int r = 0; r < 32; ++r) fk[r] *= ip;
        const int dmax = (NA - 1 - k < HBW) ? NA - 1 - k : HBW;
        for (int j = 1; j <= dmax; ++j) {          /* column k + j */
            const double lj = col[j];
            if (lj == 0.0) continue;
            double *cj = Lb + (k + j) * 33;
            for (int d = 0; j + d <= dmax; ++d) cj[d] -= col[j + d] * lj;
            double *fj = F + (k + j) * 32;
            for (int r = 0; r < 32; ++r) fj[r] -= fk[r] * lj;
        }
        for (int r = 0; r < 32; ++r) {
            const double fr = fk[r];
            for (int c = 0; c <= r; ++c) S[r * 32 + c] -= fr * fk[c];
        }
    }
    for (int k = 0; k < 32; ++k) {                 /* dense Cholesky of the separator (lower) */
        double s = S[k * 32 + k];
        if (!(s > 0.0)) return 0;
        s = sqrt(s); S[k * 32 + k] = s;
        for (int r = k + 1; r < 32; ++r) S[r * 32 + k] /= s;
        for (int c = k + 1; c < 32; ++c)
            for (int r = c; r < 32; ++r) S[r * 32 + c] -= S[r * 32 + k] * S[c * 32 + k];
    }
    return 1;
}
static void solve(const Fact *f, const double *g, double *x, double *y) {
    const int NA = f->NA;
    const double *Lb = f->Lb, *F = f->F, *S = f->S;
    double gs[32];
    for (int k = 0; k < NA; ++k) y[k] = g[k];
    for (int r = 0; r < 32; ++r) gs[r] = g[NA + r];
    for (int k = 0; k < NA; ++k) {
        const double *col = Lb + k * 33;
        const double yk = y[k] / col[0];
        y[k] = yk;
        const int dmax = (NA - 1 - k < HBW) ? NA - 1 - k : HBW;
        for (int d = 1; d <= dmax; ++d) y[k + d] -= col[d] * yk;
        const double *fk = F + k * 32;
        for (int r = 0; r < 32; ++r) gs[r] -= fk[r] * yk;
    }
    for (int k = 0; k < 32; ++k) { gs[k] /= S[k * 32 + k]; for (int r = k + 1; r < 32; ++r) gs[r] -= S[r * 32 + k] * gs[k]; }
    for (int k = 31; k >= 0; --k) { for (int r = k + 1; r < 32; ++r) gs[k] -= S[r * 32 + k] * gs[r]; gs[k] /= S[k * 32 + k]; }
    for (int r = 0; r < 32; ++r) x[NA + r] = gs[r];
    for (int k = NA - 1; k >= 0; --k) {
        const double *col = Lb + k * 33, *fk = F + k * 32;
        double v = y[k];
        for (int r = 0; r < 32; ++r) v -= fk[r] * gs[r];
        const int dmax = (NA - 1 - k < HBW) ? NA - 1 - k : HBW;
        for (int d = 1; d <= dmax; ++d) v -= col[d] * x[k + d];
        x[k] = v / col[0];
    }
}

/* returns the number of interior-point iterations (> 0), 0 if the track is too narrow, -1 on numerical breakdown, -2 at the cap */
int banded_mincurv_solve(int n, const double *reftrack, const double *normvec, const double *h_in, double w_veh,
                         double f_scale, double *alpha) {
    if (n < 80) return -3;
    const size_t N = (size_t)n;
    double *buf = (double *)calloc(N * (40 + 2 * (2 * BZ + 3) + 34 + 33 + 33 + 32) + 2048, sizeof(double));
    if (!buf) return -4;
    double *p = buf;
#define TAKE(k) (p += (k), p - (k))
    double *h = TAKE(N), *diag = TAKE(N), *d = TAKE(N), *dl = TAKE(N), *rx = TAKE(N), *ry = TAKE(N), *mx = TAKE(N), *my = TAKE(N);
    double *ytmp = TAKE(N), *sx = TAKE(N), *sy = TAKE(N), *kref = TAKE(N), *f = TAKE(N), *lb = TAKE(N), *ub = TAKE(N);
    double *tii = TAKE(N), *rhop = TAKE(N), *rhom = TAKE(N), *t0 = TAKE(N), *t1 = TAKE(N);
    double *al = TAKE(N), *lu = TAKE(N), *ll = TAKE(N), *su = TAKE(N), *sl = TAKE(N), *rd = TAKE(N), *rhs = TAKE(N), *dx = TAKE(N);
    double *tu = TAKE(N), *tl = TAKE(N), *dd = TAKE(N), *g0 = TAKE(N), *yy = TAKE(N);
    double *TB = TAKE(N * (2 * BZ + 3)), *EB = TAKE(N * (2 * BZ + 3));
    double *HB = TAKE(N * 34), *MB = TAKE(N * 33);
    Fact fc; fc.Lb = TAKE(N * 33); fc.F = TAKE(N * 32); fc.S = TAKE(1024);
    const double *nx = normvec, *px = reftrack;
    int ret = -2;
    for (int i = 0; i < n; ++i) {
        h[i] = h_in[i];
        ub[i] = reftrack[4 * i + 2] - 0.5 * w_veh;
        lb[i] = -(reftrack[4 * i + 3] - 0.5 * w_veh);
        if (lb[i] > ub[i]) { ret = 0; goto done; }
        if (ub[i] - lb[i] < 2e-8) { double mid = 0.5 * (lb[i] + ub[i]); lb[i] = mid - 1e-8; ub[i] = mid + 1e-8; }
    }
    for (int i = 0; i < n; ++i) diag[i] = 2.0 * (h[wrap(i - 1, n)] + h[i]);
    tri_pivots(n, diag, h, d, dl);
    for (int i = 0; i < n; ++i) {
        int im = wrap(i - 1, n), ip = wrap(i + 1, n);
        rx[i] = 6.0 * ((px[4 * ip] - px[4 * i]) / h[i] - (px[4 * i] - px[4 * im]) / h[im]);
        ry[i] = 6.0 * ((px[4 * ip + 1] - px[4 * i + 1]) / h[i] - (px[4 * i + 1] - px[4 * im + 1]) / h[im]);
    }
    tri_solve(n, d, h, rx, mx, ytmp);
    tri_solve(n, d, h, ry, my, ytmp);
    for (int i = 0; i < n; ++i) {
        int ip = wrap(i + 1, n);
        const double h2 = h[i] * h[i];
        const double xp = (px[4 * ip] - px[4 * i]) - h2 * (2.0 * mx[i] + mx[ip]) / 6.0;
        const double yp = (px[4 * ip + 1] - px[4 * i + 1]) - h2 * (2.0 * my[i] + my[ip]) / 6.0;
        const double q = xp * xp + yp * yp, c = 1.0 / (q * sqrt(q));
        sy[i] = c * xp * h2; sx[i] = c * yp * h2;
        kref[i] = sy[i] * my[i] - sx[i] * mx[i];
    }
    /* f = f_scale E^T k_ref */
    for (int i = 0; i < n; ++i) { t0[i] = sx[i] * kref[i]; t1[i] = sy[i] * kref[i]; }
    tri_solve(n, d, h, t0, rx, ytmp);
    tri_solve(n, d, h, t1, ry, ytmp);
    for (int i = 0; i < n; ++i) {
        int im = wrap(i - 1, n), ip = wrap(i + 1, n);
        const double zx = 6.0 * ((rx[ip] - rx[i]) / h[i] - (rx[i] - rx[im]) / h[im]);
        const double zy = 6.0 * ((ry[ip] - ry[i]) / h[i] - (ry[i] - ry[im]) / h[im]);
        f[i] = f_scale * (nx[2 * i + 1] * zy - nx[2 * i] * zx);
    }
    /* banded rows of E, band of H */
    {
        const int WB = 2 * BZ + 3;
        for (int i = 0; i < n; ++i) {
            tii[i] = 1.0 / (d[i] + dl[i] - diag[i]);
            rhop[i] = -h[wrap(i - 1, n)] / dl[i];
            rhom[i] = -h[i] / d[i];
        }
        for (int m = 0; m < n; ++m) {
            double *tb = TB + (size_t)m * WB;
            tb[BZ + 1] = tii[m];
            double v = tii[m];
            for (int o = 1; o <= BZ + 1; ++o) { v *= rhop[wrap(m + o, n)]; tb[BZ + 1 + o] = v; }
            v = tii[m];
            for (int o = 1; o <= BZ + 1; ++o) { v *= rhom[wrap(m - o, n)]; tb[BZ + 1 - o] = v; }
        }
        for (int m = 0; m < n; ++m) {
            const double *tb = TB + (size_t)m * WB;
            double *eb = EB + (size_t)m * WB;
            for (int o = -BZ; o <= BZ; ++o) {
                int i = wrap(m + o, n), im = wrap(i - 1, n);
                const double z = 6.0 * (tb[BZ + 1 + o - 1] / h[im] - tb[BZ + 1 + o] * (1.0 / h[im] + 1.0 / h[i]) + tb[BZ + 1 + o + 1] / h[i]);
                eb[BZ + o] = z * (sy[m] * nx[2 * i + 1] - sx[m] * nx[2 * i]);
            }
        }
        for (int i = 0; i < n; ++i)
            for (int k = 0; k <= HBW; ++k) {
                double s = 0.0;
                for (int m = i + k - BZ; m <= i + BZ; ++m) {
                    const double *eb = EB + (size_t)wrap(m, n) * WB;
                    s += eb[BZ + (i - m)] * eb[BZ + (i + k - m)];
                }
                HB[i * 34 + k] = s;
            }
    }
#define HMATVEC(v, out) do { for (int i_ = 0; i_ < n; ++i_) { double s_ = HB[i_ * 34] * (v)[i_]; \
        for (int k_ = 1; k_ <= HBW; ++k_) { s_ += HB[i_ * 34 + k_] * (v)[wrap(i_ + k_, n)] + HB[wrap(i_ - k_, n) * 34 + k_] * (v)[wrap(i_ - k_, n)]; } (out)[i_] = s_; } } while (0)
    /* ---- Mehrotra predictor-corrector (same start, step rule and stopping tests as csrc/mincurv_ipm.cu) ---- */
    for (int i = 0; i < n; ++i) al[i] = 0.5 * (lb[i] + ub[i]);
    HMATVEC(al, g0);
    double gmax = 0.0, fmax_ = 0.0;
    for (int i = 0; i < n; ++i) { g0[i] += f[i]; if (fabs(g0[i]) > gmax) gmax = fabs(g0[i]); if (fabs(f[i]) > fmax_) fmax_ = fabs(f[i]); }
    const double lam0 = 1e-2 * gmax + 1e-300;
    double musum = 0.0;
    for (int i = 0; i < n; ++i) {
        lu[i] = fmax(-g0[i], 0.0) + lam0; ll[i] = fmax(g0[i], 0.0) + lam0;
        rd[i] = g0[i] + lu[i] - ll[i];
        su[i] = ub[i] - al[i]; sl[i] = al[i] - lb[i];
        musum += su[i] * lu[i] + sl[i] * ll[i];
    }
    const double mu0 = musum / (2.0 * n), rd_tol = 1e-8 * (fmax_ + gmax) + 1e-300, mu_rel = 1e-10, eta = 0.995, dx_rel = 1e-5;
    double mu = mu0;
    for (int it = 0; it < 40; ++it) {
        for (int i = 0; i < n; ++i) { dd[i] = lu[i] / su[i] + ll[i] / sl[i]; rhs[i] = -rd[i] + lu[i] - ll[i]; }
        for (int i = 0; i < n; ++i) { for (int k = 0; k <= HBW; ++k) MB[i * 33 + k] = HB[i * 34 + k]; MB[i * 33] += dd[i]; }
        if (!factor(MB, n, &fc)) { ret = -1; goto done; }
        solve(&fc, rhs, dx, yy);
        double rp = 0.0, rdl = 0.0, c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
        for (int i = 0; i < n; ++i) {
            const double pp = dx[i] / su[i], mm = dx[i] / sl[i];
            rp = fmax(rp, fmax(pp, -mm)); rdl = fmax(rdl, fmax(1.0 - pp, 1.0 + mm));
            const double dlu = lu[i] * (pp - 1.0), dll = -ll[i] * (1.0 + mm);
            c00 += su[i] * lu[i] + sl[i] * ll[i]; c01 += su[i] * dlu + sl[i] * dll;
            c10 += dx[i] * (ll[i] - lu[i]); c11 += dx[i] * (dll - dlu);
        }
        double ap = rp > 1.0 ? 1.0 / rp : 1.0, ad = rdl > 1.0 ? 1.0 / rdl : 1.0;
        const double mua = (c00 + ad * c01 + ap * c10 + ap * ad * c11) / (2.0 * n);
        double sigma = mua / mu; sigma = sigma * sigma * sigma;
        const double smu = sigma * mu;
        for (int i = 0; i < n; ++i) {
            const double dlu = lu[i] * (dx[i] / su[i] - 1.0), dll = -ll[i] * (1.0 + dx[i] / sl[i]);
            tu[i] = smu - su[i] * lu[i] + dx[i] * dlu; tl[i] = smu - sl[i] * ll[i] - dx[i] * dll;
            rhs[i] = -rd[i] - tu[i] / su[i] + tl[i] / sl[i];
        }
        solve(&fc, rhs, dx, yy);
        rp = 0.0; rdl = 0.0;
        for (int i = 0; i < n; ++i) {
            const double dlu = (tu[i] + lu[i] * dx[i]) / su[i], dll = (tl[i] - ll[i] * dx[i]) / sl[i];
            rp = fmax(rp, fmax(dx[i] / su[i], -dx[i] / sl[i])); rdl = fmax(rdl, fmax(-dlu / lu[i], -dll / ll[i]));
        }
        ap = eta < rp ? eta / rp : 1.0; ad = eta < rdl ? eta / rdl : 1.0;
        double musum2 = 0.0, rdmax = 0.0, dxmax = 0.0, amax = 0.0;
        for (int i = 0; i < n; ++i) {
            const double dlu = (tu[i] + lu[i] * dx[i]) / su[i], dll = (tl[i] - ll[i] * dx[i]) / sl[i];
            rd[i] += ap * (rhs[i] - dd[i] * dx[i]) + ad * (dlu - dll);
            al[i] += ap * dx[i]; lu[i] += ad * dlu; ll[i] += ad * dll; su[i] -= ap * dx[i]; sl[i] += ap * dx[i];
            musum2 += su[i] * lu[i] + sl[i] * ll[i];
            rdmax = fmax(rdmax, fabs(rd[i])); dxmax = fmax(dxmax, fabs(dx[i])); amax = fmax(amax, fabs(al[i]));
        }
        mu = musum2 / (2.0 * n);
        const int settled = !(mu <= mu_rel * mu0) || ap * dxmax <= dx_rel * fmax(amax, 0.01);
        if (mu <= mu_rel * mu0 && rdmax <= rd_tol && settled) { ret = it + 1; break; }
        if (mu <= 1e-4 * mu_rel * mu0) { ret = (rdmax <= 1e3 * rd_tol) ? it + 1 : -2; break; }
    }
    for (int i = 0; i < n; ++i) alpha[i] = al[i];
done:
    free(buf);
    return ret;
}
