/*
 * CPU ORACLE (test infrastructure, NOT product code).
 *
 * Dense Goldfarb-Idnani dual active-set solver for strictly convex QPs, restating the
 * published algorithm (D. Goldfarb, A. Idnani, "A numerically stable dual method for solving
 * strictly convex quadratic programs", Math. Prog. 27 (1983) 1-33) behind the call semantics
 * of the third-party `quadprog` package (versions 0.1.6/0.1.7 named at
 * /root/reference/Readme.md:40; transitive dependency of trajectory_planning_helpers==0.76,
 * /root/reference/requirements.txt:3), which is what tph.opt_min_curv / tph.opt_shortest_path
 * call as  quadprog.solve_qp(H, -f, -G.T, -h, 0)[0]   (SURVEY.md A.3/A.4/A.8):
 *
 *     minimise 1/2 x^T G x - a^T x   subject to   C^T x >= b   (first meq rows equalities)
 *
 * PARITY UNPINNED: quadprog's source is not vendored under /root/reference and not installable
 * offline; this file follows the published algorithm.  Because the QP is strictly convex its
 * minimiser is unique, so any exact active-set solver returns the same x up to rounding.
 *
 * Storage is column-major throughout (C is n x m, one constraint normal per column).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define QP_OK 0
#define QP_INFEASIBLE 1      /* "constraints are inconsistent, no solution" */
#define QP_NOT_PD 2          /* "matrix G is not positive definite"          */
#define QP_MAXITER 3

static double hypot2(double a, double b) {
    double aa = fabs(a), bb = fabs(b);
    if (aa > bb) { double t = bb / aa; return aa * sqrt(1.0 + t * t); }
    if (bb > 0.0) { double t = aa / bb; return bb * sqrt(1.0 + t * t); }
    return 0.0;
}

/* in-place lower Cholesky of column-major n x n G; returns 0 on success */
static int cholesky_lower(double *G, int n) {
    for (int j = 0; j < n; ++j) {
        double d = G[j + (size_t)j * n];
        for (int k = 0; k < j; ++k) { double l = G[j + (size_t)k * n]; d -= l * l; }
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        G[j + (size_t)j * n] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = G[i + (size_t)j * n];
            for (int k = 0; k < j; ++k) s -= G[i + (size_t)k * n] * G[j + (size_t)k * n];
            G[i + (size_t)j * n] = s / d;
        }
    }
    return 0;
}

/*
 * qp_solve_gi: returns status code. Outputs: x[n], lagr[m] (multipliers), iact[m] (active set,
 * 1-based like quadprog), *nact, iters[2] = {additions, deletions}, *fval (objective value).
 */
int qp_solve_gi(int n, int m, int meq,
                const double *Gin, const double *a, const double *C, const double *b,
                double *x, double *lagr, int *iact, int *nact, int *iters, double *fval)
{
    const size_t nn = (size_t)n * n;
    double *L = (double *)malloc(nn * sizeof(double));    /* Cholesky factor of G (lower)       */
    double *J = (double *)malloc(nn * sizeof(double));    /* J = L^{-T} Q, column-major          */
    double *R = (double *)calloc(nn, sizeof(double));     /* upper triangular, q x q used        */
    double *d = (double *)malloc(n * sizeof(double));
    double *z = (double *)malloc(n * sizeof(double));
    double *r = (double *)malloc(n * sizeof(double));
    double *u = (double *)calloc(n + 1, sizeof(double));  /* multipliers of active constraints   */
    double *s = (double *)malloc((m > 0 ? m : 1) * sizeof(double));
    double *cnorm = (double *)malloc((m > 0 ? m : 1) * sizeof(double));
    int *A = (int *)malloc((n + 1) * sizeof(int));        /* active constraint ids (0-based)     */
    char *isact = (char *)calloc(m > 0 ? m : 1, 1);
    int status = QP_OK, q = 0, n_add = 0, n_del = 0;
    const double eps = 2.220446049250313e-16;

    memcpy(L, Gin, nn * sizeof(double));
    if (cholesky_lower(L, n)) { status = QP_NOT_PD; goto done; }

    /* J = L^{-T}: solve L^T J = I column by column (J upper triangular initially) */
    memset(J, 0, nn * sizeof(double));
    for (int c = 0; c < n; ++c) {
        /* back substitution for column c of L^{-T}: L^T y = e_c */
        for (int i = c; i >= 0; --i) {
            double sum = (i == c) ? 1.0 : 0.0;
            for (int k = i + 1; k <= c; ++k) sum -= L[k + (size_t)i * n] * J[k + (size_t)c * n];
            J[i + (size_t)c * n] = sum / L[i + (size_t)i * n];
        }
    }
    /* unconstrained minimum x = G^{-1} a : L y = a, L^T x = y */
    for (int i = 0; i < n; ++i) {
        double sum = a[i];
        for (int k = 0; k < i; ++k) sum -= L[i + (size_t)k * n] * x[k];
        x[i] = sum / L[i + (size_t)i * n];
    }
    for (int i = n - 1; i >= 0; --i) {
        double sum = x[i];
        for (int k = i + 1; k < n; ++k) sum -= L[k + (size_t)i * n] * x[k];
        x[i] = sum / L[i + (size_t)i * n];
    }
    for (int j = 0; j < m; ++j) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += C[i + (size_t)j * n] * C[i + (size_t)j * n];
        cnorm[j] = sqrt(t);
    }

    const int max_iter = 40 * (n + m) + 100;
    for (int it = 0; it < max_iter; ++it) {
        /* ---- step 1: pick the most violated constraint (equalities first, in order) ---- */
        int p = -1;
        double worst = 0.0;
        for (int j = 0; j < m; ++j) {
            double t = -b[j];
            const double *cj = C + (size_t)j * n;
            for (int i = 0; i < n; ++i) t += cj[i] * x[i];
            s[j] = t;
        }
        for (int j = 0; j < meq && p < 0; ++j)
            if (!isact[j]) { p = j; }
        if (p < 0) {
            for (int j = meq; j < m; ++j) {
                if (isact[j]) continue;
                /* scaled violation, tolerance relative to the constraint-normal size */
                double tol = 100.0 * eps * (cnorm[j] > 1.0 ? cnorm[j] : 1.0);
                double v = s[j] / (cnorm[j] > 0.0 ? cnorm[j] : 1.0);
                if (s[j] < -tol * (1.0 + fabs(b[j])) && v < worst) { worst = v; p = j; }
            }
        }
        if (p < 0) break;                                  /* all constraints satisfied: optimal */

        const double *np_ = C + (size_t)p * n;
        double up = 0.0;                                   /* multiplier of the entering constraint */
        const int is_eq = (p < meq);
        if (is_eq && s[p] > 0.0) { /* equality violated from above: use -n, handled by sign flip */ }
        const double sgn = (is_eq && s[p] > 0.0) ? -1.0 : 1.0;

        for (;;) {
            /* ---- step 2a: d = J^T n+, z = J2 d2, r = R^{-1} d1 ---- */
            for (int c = 0; c < n; ++c) {
                double t = 0.0;
                const double *jc = J + (size_t)c * n;
                for (int i = 0; i < n; ++i) t += jc[i] * np_[i];
                d[c] = sgn * t;
            }
            for (int i = 0; i < n; ++i) z[i] = 0.0;
            for (int c = q; c < n; ++c) {
                const double *jc = J + (size_t)c * n;
                const double dc = d[c];
                for (int i = 0; i < n; ++i) z[i] += jc[i] * dc;
            }
            for (int i = q - 1; i >= 0; --i) {
                double sum = d[i];
                for (int k = i + 1; k < q; ++k) sum -= R[i + (size_t)k * n] * r[k];
                r[i] = sum / R[i + (size_t)i * n];
            }
            /* ---- step 2b: step lengths ---- */
            int l = -1;
            double t1 = INFINITY;
            for (int k = 0; k < q; ++k) {
                if (A[k] < meq) continue;                  /* equalities are never dropped */
                if (r[k] > 0.0) {
                    double t = u[k] / r[k];
                    if (t < t1) { t1 = t; l = k; }
                }
            }
            double znorm2 = 0.0, zn = 0.0;
            for (int i = 0; i < n; ++i) { znorm2 += z[i] * z[i]; zn += z[i] * np_[i]; }
            zn *= sgn;
            double sp = sgn * s[p];
            double t2 = INFINITY;
            if (fabs(znorm2) > eps * eps && zn > 0.0) t2 = -sp / zn;
            double t = (t1 < t2) ? t1 : t2;

            if (!isfinite(t)) { status = QP_INFEASIBLE; goto done; }

            if (!isfinite(t2)) {
                /* dual step only, drop constraint l */
                for (int k = 0; k < q; ++k) u[k] -= t * r[k];
                up += t;
            } else {
                for (int i = 0; i < n; ++i) x[i] += t * z[i];
                for (int k = 0; k < q; ++k) u[k] -= t * r[k];
                up += t;
                if (t == t2) {
                    /* ---- full step: add constraint p (Givens on d[q..n-1], update J, R) ---- */
                    for (int c = n - 1; c > q; --c) {
                        double g1 = d[c - 1], g2 = d[c];
                        if (g2 == 0.0) continue;
                        double h = hypot2(g1, g2);
                        double cs = g1 / h, sn = g2 / h;
                        d[c - 1] = h; d[c] = 0.0;
                        double *ja = J + (size_t)(c - 1) * n, *jb = J + (size_t)c * n;
                        for (int i = 0; i < n; ++i) {
                            double t_a = ja[i], t_b = jb[i];
                            ja[i] = cs * t_a + sn * t_b;
                            jb[i] = -sn * t_a + cs * t_b;
                        }
                    }
                    for (int i = 0; i <= q; ++i) R[i + (size_t)q * n] = d[i];
                    if (fabs(d[q]) <= eps * (cnorm[p] > 1.0 ? cnorm[p] : 1.0) * 1e-3) {
                        /* linearly dependent on the active set: cannot add */
                        status = QP_INFEASIBLE; goto done;
                    }
                    u[q] = up;
                    A[q] = p;
                    isact[p] = 1;
                    ++q; ++n_add;
                    break;                                 /* back to step 1 */
                }
                /* partial step: recompute s_p for the moved x */
                double tt = -b[p];
                for (int i = 0; i < n; ++i) tt += np_[i] * x[i];
                s[p] = tt;
            }
            /* ---- drop constraint l from the active set ---- */
            {
                isact[A[l]] = 0;
                for (int k = l; k < q - 1; ++k) {
                    A[k] = A[k + 1];
                    u[k] = u[k + 1];
                    for (int i = 0; i <= k + 1; ++i) R[i + (size_t)k * n] = R[i + (size_t)(k + 1) * n];
                }
                /* restore triangularity: column k (k >= l) has a subdiagonal entry at row k+1 */
                for (int k = l; k < q - 1; ++k) {
                    double g1 = R[k + (size_t)k * n], g2 = R[(k + 1) + (size_t)k * n];
                    if (g2 == 0.0) continue;
                    double h = hypot2(g1, g2);
                    double cs = g1 / h, sn = g2 / h;
                    R[k + (size_t)k * n] = h; R[(k + 1) + (size_t)k * n] = 0.0;
                    for (int c = k + 1; c < q - 1; ++c) {
                        double t_a = R[k + (size_t)c * n], t_b = R[(k + 1) + (size_t)c * n];
                        R[k + (size_t)c * n] = cs * t_a + sn * t_b;
                        R[(k + 1) + (size_t)c * n] = -sn * t_a + cs * t_b;
                    }
                    double *ja = J + (size_t)k * n, *jb = J + (size_t)(k + 1) * n;
                    for (int i = 0; i < n; ++i) {
                        double t_a = ja[i], t_b = jb[i];
                        ja[i] = cs * t_a + sn * t_b;
                        jb[i] = -sn * t_a + cs * t_b;
                    }
                }
                for (int i = 0; i < q; ++i) R[i + (size_t)(q - 1) * n] = 0.0;
                u[q - 1] = 0.0;
                --q; ++n_del;
            }
        }
        if (it == max_iter - 1) status = QP_MAXITER;
    }

    for (int j = 0; j < m; ++j) lagr[j] = 0.0;
    for (int k = 0; k < q; ++k) { lagr[A[k]] = u[k]; iact[k] = A[k] + 1; }
    *nact = q;
    iters[0] = n_add; iters[1] = n_del;
    {
        double fv = 0.0;
        for (int i = 0; i < n; ++i) {
            double t = 0.0;
            for (int k = 0; k < n; ++k) t += Gin[i + (size_t)k * n] * x[k];
            fv += x[i] * (0.5 * t - a[i]);
        }
        *fval = fv;
    }
done:
    free(L); free(J); free(R); free(d); free(z); free(r); free(u); free(s); free(cnorm); free(A); free(isact);
    return status;
}
