// K2b -- box-constrained QP solve  min 1/2 a^T H a + f^T a,  lb <= a <= ub  by a Mehrotra
// predictor-corrector primal-dual interior-point method (replaces quadprog.solve_qp inside
// tph.opt_min_curv, call site /root/reference/main_globaltraj.py:264-271).
//
// One CTA of two warps per QP instance.  H is the cyclic band (half-bandwidth 32) assembled by K2a.
// Every interior-point iteration factorises M = H + D (D diagonal, from the barrier) with a
// block-cyclic Cholesky on 32x32 blocks:
//   chain blocks I = 0..nb-1 (nodes 0..n-33, padded), separator S = last 32 nodes (closes the cycle)
//     A'_I   = A_I + D_I - T_I T_I^T                 T_I   = L_{I,I-1}
//     L_II   = chol(A'_I),  Linv_I = L_II^{-1}       (explicit inverse: sweeps become mat-vecs)
//     T_{I+1}= B_I Linv_I^T                          B_I   = M[block I+1, block I]
//     F_I    = (Y_I - F_{I-1} T_I^T) Linv_I^T        Y_I   = M[S, block I]   (fill row of the separator)
//     S     -= F_I F_I^T
// warp 0 owns the chain (A', chol, inverse, T), warp 1 owns the separator row (F, S); both keep one
// 32-entry row per lane in registers and read the other operand as broadcast from swizzled
// shared-memory tiles.  Tiles of the factor (Linv_I, T_I, F_I) go to the instance's HBM slab in
// column-major 8 KB tiles; the triangular sweeps are sequences of 32x32 mat-vecs over them.
#include "mincurv_ws.cuh"

namespace mc {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PD_THREADS = 64;

// 32x32 shared tile, row-major with a 16-byte pair swizzle: conflict-free for "lane = row" vector stores
// and for broadcast pair loads with compile-time (row, pair).
__device__ __forceinline__ int tidx(int r, int k) { return r * 32 + 2 * ((k >> 1) ^ (r & 15)) + (k & 1); }
__device__ __forceinline__ double2 tpair(const double *t, int r, int p) {
    return *reinterpret_cast<const double2 *>(t + r * 32 + 2 * (p ^ (r & 15)));
}
// store this lane's row x[0..31] as row `lane` of tile t
__device__ __forceinline__ void tile_store_row(double *t, const double (&x)[32], int lane) {
#pragma unroll
    for (int p = 0; p < 16; ++p)
        *reinterpret_cast<double2 *>(t + lane * 32 + 2 * (p ^ (lane & 15))) = make_double2(x[2 * p], x[2 * p + 1]);
}

// acc[c] -= sum_k x[k] * Y[c][k]   (C -= X Y^T, X rows in registers, Y broadcast from shared)
__device__ __forceinline__ void gemm_sub_xyT(double (&acc)[32], const double (&x)[32], const double *Y) {
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const double2 y = tpair(Y, c, p);
            s0 = fma(x[2 * p], y.x, s0);
            s1 = fma(x[2 * p + 1], y.y, s1);
        }
        acc[c] -= (s0 + s1);
    }
}
// out[c] = sum_{k <= c} x[k] * Y[c][k]   (X Y^T with Y lower triangular); in place (descending c)
__device__ __forceinline__ void trmm_inplace_xLT(double (&x)[32], const double *Y) {
#pragma unroll
    for (int c = 31; c >= 0; --c) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int p = 0; p <= (c >> 1); ++p) {
            const double2 y = tpair(Y, c, p);
            s0 = fma(x[2 * p], y.x, s0);
            if (2 * p + 1 <= c) s1 = fma(x[2 * p + 1], y.y, s1);
        }
        x[c] = s0 + s1;
    }
}

// Cholesky of a 32x32 SPD block, one row per lane (entries c <= lane meaningful).
// On exit a[c] = L[lane][c] for c <= lane (entries c > lane are garbage). lcol: 32 doubles of shared.
__device__ __forceinline__ bool chol32(double (&a)[32], double *lcol, int lane) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const double d = __shfl_sync(FULL, a[k], k);
        if (!(d > 0.0)) ok = false;
        const double l = a[k] * rsqrt(d);
        a[k] = l;
        lcol[lane] = l;
        __syncwarp();
#pragma unroll
        for (int c = k + 1; c < 32; ++c) a[c] = fma(-l, lcol[c], a[c]);
        __syncwarp();
    }
    return ok;
}

// lane c computes column c of L^{-1}: x[r] = Linv[r][c].  Ls: swizzled tile of L (lower part valid),
// dinv[r] = 1 / L[r][r] (shared).
__device__ __forceinline__ void trinv32(double (&x)[32], const double *Ls, const double *dinv, int lane) {
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        double s0 = (r == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int p = 0; p < (r + 1) / 2; ++p) {
            const double2 y = tpair(Ls, r, p);
            if (p & 1) {
                s2 = fma(-y.x, x[2 * p], s2);
                if (2 * p + 1 < r) s3 = fma(-y.y, x[2 * p + 1], s3);
            } else {
                s0 = fma(-y.x, x[2 * p], s0);
                if (2 * p + 1 < r) s1 = fma(-y.y, x[2 * p + 1], s1);
            }
        }
        x[r] = ((s0 + s1) + (s2 + s3)) * dinv[r];
    }
}

// entry M[i][j] of the cyclic band (real node indices, i != j allowed in any order), 0 outside the band
__device__ __forceinline__ double hentry(const double *HB, int n, int i, int j) {
    int k = j - i;
    if (k < 0) k += n;
    if (k <= HBW) return HB[(size_t)i * HB_PITCH + k];
    k = n - k;
    if (k <= HBW) return HB[(size_t)j * HB_PITCH + k];
    return 0.0;
}

struct PdShared {
    double band[32 * HB_PITCH];   // band rows of the current chain block
    double Ls[1024];              // L_II (swizzled)
    double Li[1024];              // Linv_I (swizzled, row-major: Li[c][k] = Linv[c][k])
    double Ts[1024];              // T_I = L_{I,I-1} (swizzled)
    double Fs[1024];              // F_I (swizzled)
    double Ss[1024];              // separator Schur complement, Ss[c * 32 + lane] = S[lane][c]
    double lcol[2][32];
    double dinv[32];
    double vbuf[4][32];
    double red[32];
    int flag;
};

// ------------------------------------------------------------------------------------------------
// factorisation of M = H + diag(DD); returns false on a non-positive pivot
// ------------------------------------------------------------------------------------------------
__device__ bool factor(PdShared &sh, const double *__restrict__ HB, const double *__restrict__ DD,
                       double *__restrict__ tiles, int n, int nb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NA = n - 32;
    bool ok = true;
    double t[32];    // warp 0: row `lane` of T_I ; warp 1: row `lane` of F_{I-1}
#pragma unroll
    for (int c = 0; c < 32; ++c) t[c] = 0.0;

    if (warp == 1) {   // separator diagonal block C + D
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            double v = hentry(HB, n, NA + lane, NA + c);
            if (c == lane) v += DD[NA + lane];
            sh.Ss[c * 32 + lane] = v;
        }
    }

    for (int I = 0; I < nb; ++I) {
        const int base = 32 * I;
        // ---- stage the band rows of block I (coalesced) ----
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * HB_PITCH; e += PD_THREADS) {
            const int row = base + e / HB_PITCH;
            sh.band[e] = (row < n) ? HB[(size_t)base * HB_PITCH + e] : 0.0;
        }
        __syncthreads();
        // =========================== phase A ===========================
        if (warp == 0) {
            double a[32];
            const int node = base + lane;
            const bool real = node < NA;
            // m1: row `lane` of A_I + D_I (lower part incl. diagonal is what chol32 needs; fill all)
            const double dd = real ? DD[node] : 0.0;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int lo = (c < lane) ? c : lane, dist = (c < lane) ? lane - c : c - lane;
                double v = sh.band[lo * HB_PITCH + dist];
                const bool creal = (base + c) < NA;
                if (!(real && creal)) v = (c == lane) ? 1.0 : 0.0;
                else if (c == lane) v += dd;
                a[c] = v;
            }
            // m2: A' = A - T_I T_I^T
            if (I > 0) gemm_sub_xyT(a, t, sh.Ts);
            // m3: Cholesky
            ok = chol32(a, sh.lcol[0], lane) && ok;
            // m4: L_II -> shared (zero the strict upper part)
#pragma unroll
            for (int c = 0; c < 32; ++c) if (c > lane) a[c] = 0.0;
            tile_store_row(sh.Ls, a, lane);
#pragma unroll
            for (int c = 0; c < 32; ++c) if (c == lane) sh.dinv[lane] = 1.0 / a[c];
            __syncwarp();
            // m5: explicit inverse, lane = column
            double xi[32];
            trinv32(xi, sh.Ls, sh.dinv, lane);
            // Linv -> shared row-major (Li[r][lane] = xi[r]) and HBM tile (column-major: [lane*32 + r])
            double *gt = tiles + (size_t)(3 * I + 0) * 1024 + lane * 32;
#pragma unroll
            for (int r = 0; r < 32; ++r) sh.Li[tidx(r, lane)] = xi[r];
#pragma unroll
            for (int r = 0; r < 32; r += 2) *reinterpret_cast<double2 *>(gt + r) = make_double2(xi[r], xi[r + 1]);
        } else {
            // f4 (previous block): S -= F_{I-1} F_{I-1}^T
            if (I > 0) {
                double s[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) s[c] = sh.Ss[c * 32 + lane];
                gemm_sub_xyT(s, t, sh.Fs);
#pragma unroll
                for (int c = 0; c < 32; ++c) sh.Ss[c * 32 + lane] = s[c];
            }
            // f2: FW = Y_I - F_{I-1} T_I^T   (Y_I nonzero only next to the separator / across the wrap)
            double fw[32];
            const bool hasY = (I == 0) || (base + 31 >= NA - 32);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                double v = 0.0;
                if (hasY && (base + c) < NA) v = hentry(HB, n, NA + lane, base + c);
                fw[c] = v;
            }
            if (I > 0) gemm_sub_xyT(fw, t, sh.Ts);
#pragma unroll
            for (int c = 0; c < 32; ++c) t[c] = fw[c];
        }
        __syncthreads();
        // =========================== phase B ===========================
        if (warp == 0) {
            if (I + 1 < nb) {
                // m6: T_{I+1} = B_I Linv_I^T, B_I[r][k] = M[base+32+r][base+k] = band[k][32 + r - k] (r <= k)
                const bool real = (base + 32 + lane) < NA;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    double v = 0.0;
                    if (lane <= k && real) v = sh.band[k * HB_PITCH + 32 + lane - k];
                    t[k] = v;
                }
                trmm_inplace_xLT(t, sh.Li);
                double *gt = tiles + (size_t)(3 * (I + 1) + 1) * 1024;
#pragma unroll
                for (int c = 0; c < 32; ++c) gt[c * 32 + lane] = t[c];
                // Ts is only read in phase A: safe to publish T_{I+1} now
                tile_store_row(sh.Ts, t, lane);
            }
        } else {
            // f3: F_I = FW Linv_I^T (in place in t)
            trmm_inplace_xLT(t, sh.Li);
            double *gt = tiles + (size_t)(3 * I + 2) * 1024;
#pragma unroll
            for (int c = 0; c < 32; ++c) gt[c * 32 + lane] = t[c];
            tile_store_row(sh.Fs, t, lane);   // Fs is only read in phase A (by this warp)
        }
    }
    __syncthreads();
    // ---- separator: S -= F_{nb-1} F_{nb-1}^T, chol, inverse (warp 1) ----
    if (warp == 1) {
        double s[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) s[c] = sh.Ss[c * 32 + lane];
        gemm_sub_xyT(s, t, sh.Fs);
        ok = chol32(s, sh.lcol[1], lane) && ok;
#pragma unroll
        for (int c = 0; c < 32; ++c) if (c > lane) s[c] = 0.0;
        tile_store_row(sh.Ls, s, lane);
#pragma unroll
        for (int c = 0; c < 32; ++c) if (c == lane) sh.dinv[lane] = 1.0 / s[c];
        __syncwarp();
        double xi[32];
        trinv32(xi, sh.Ls, sh.dinv, lane);
        double *gt = tiles + (size_t)(3 * nb) * 1024 + lane * 32;
#pragma unroll
        for (int r = 0; r < 32; r += 2) *reinterpret_cast<double2 *>(gt + r) = make_double2(xi[r], xi[r + 1]);
    }
    if (!ok) sh.flag = 1;
    __syncthreads();
    return sh.flag == 0;
}

// y = X v (lane = row), X column-major HBM tile, v broadcast from shared
__device__ __forceinline__ double tile_mv(const double *__restrict__ X, const double *v, int lane) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        s0 = fma(X[(c + 0) * 32 + lane], v[c + 0], s0);
        s1 = fma(X[(c + 1) * 32 + lane], v[c + 1], s1);
        s2 = fma(X[(c + 2) * 32 + lane], v[c + 2], s2);
        s3 = fma(X[(c + 3) * 32 + lane], v[c + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
// y = X^T v (lane = column of X), X column-major HBM tile: this lane's column is 32 contiguous doubles
__device__ __forceinline__ double tile_mtv(const double *__restrict__ X, const double *v, int lane) {
    const double2 *col = reinterpret_cast<const double2 *>(X + lane * 32);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int p = 0; p < 16; p += 2) {
        const double2 a = col[p], b2 = col[p + 1];
        s0 = fma(a.x, v[2 * p], s0);
        s1 = fma(a.y, v[2 * p + 1], s1);
        s2 = fma(b2.x, v[2 * p + 2], s2);
        s3 = fma(b2.y, v[2 * p + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// ------------------------------------------------------------------------------------------------
// solve M x = g with the stored factor.  g, x: real-indexed vectors (length n) in the slab.
// ypad / zpad: padded scratch vectors (>= 32 nb + 32 doubles).
// ------------------------------------------------------------------------------------------------
__device__ void solve(PdShared &sh, const double *__restrict__ tiles, const double *__restrict__ g,
                      double *__restrict__ x, double *__restrict__ ypad, double *__restrict__ zpad, int n, int nb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NA = n - 32;
    double *yb = sh.vbuf[0], *tb = sh.vbuf[1], *gs = sh.vbuf[2], *xs = sh.vbuf[3];
    __syncthreads();
    // ---- forward chain (warp 0): y_I = Linv_I (g_I - T_I y_{I-1}) ----
    if (warp == 0) {
        for (int I = 0; I < nb; ++I) {
            const int node = 32 * I + lane;
            double v = (node < NA) ? g[node] : 0.0;
            if (I > 0) v -= tile_mv(tiles + (size_t)(3 * I + 1) * 1024, yb, lane);
            __syncwarp();
            tb[lane] = v;
            __syncwarp();
            const double y = tile_mv(tiles + (size_t)(3 * I + 0) * 1024, tb, lane);
            yb[lane] = y;
            ypad[node] = y;
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- separator right-hand side: gS = g_S - sum_I F_I y_I (both warps, blocks interleaved) ----
    {
        double acc = 0.0;
        for (int I = warp; I < nb; I += 2) {
            const double *Ft = tiles + (size_t)(3 * I + 2) * 1024;
            const double *yv = ypad + 32 * I;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                s0 = fma(Ft[c * 32 + lane], yv[c], s0);
                s1 = fma(Ft[(c + 1) * 32 + lane], yv[c + 1], s1);
            }
            acc += s0 + s1;
        }
        if (warp == 1) tb[lane] = acc;
        __syncthreads();
        if (warp == 0) {
            gs[lane] = g[NA + lane] - acc - tb[lane];
            __syncwarp();
            const double *LS = tiles + (size_t)(3 * nb) * 1024;
            const double ys = tile_mv(LS, gs, lane);
            __syncwarp();
            tb[lane] = ys;
            __syncwarp();
            const double xv = tile_mtv(LS, tb, lane);
            xs[lane] = xv;
            x[NA + lane] = xv;
        }
        __syncthreads();
    }
    // ---- z_I = F_I^T x_S for all chain blocks (both warps) ----
    for (int I = warp; I < nb; I += 2) zpad[32 * I + lane] = tile_mtv(tiles + (size_t)(3 * I + 2) * 1024, xs, lane);
    __syncthreads();
    // ---- backward chain (warp 0): x_I = Linv_I^T (y_I - z_I - T_{I+1}^T x_{I+1}) ----
    if (warp == 0) {
        for (int I = nb - 1; I >= 0; --I) {
            const int node = 32 * I + lane;
            double v = ypad[node] - zpad[node];
            if (I + 1 < nb) v -= tile_mtv(tiles + (size_t)(3 * (I + 1) + 1) * 1024, yb, lane);
            __syncwarp();
            tb[lane] = v;
            __syncwarp();
            const double xv = tile_mtv(tiles + (size_t)(3 * I + 0) * 1024, tb, lane);
            yb[lane] = xv;
            if (node < NA) x[node] = xv;
            __syncwarp();
        }
    }
    __syncthreads();
}

// banded cyclic mat-vec out = H v (real-indexed), both warps
__device__ void band_matvec(const double *__restrict__ HB, const double *__restrict__ v, double *__restrict__ out, int n) {
    for (int i = threadIdx.x; i < n; i += PD_THREADS) {
        const double *row = HB + (size_t)i * HB_PITCH;
        double s = row[0] * v[i];
        int j = i;
        for (int k = 1; k <= HBW; ++k) {
            j = (j + 1 == n) ? 0 : j + 1;
            s = fma(row[k], v[j], s);
        }
        j = i;
        for (int k = 1; k <= HBW; ++k) {
            j = (j == 0) ? n - 1 : j - 1;
            s = fma(HB[(size_t)j * HB_PITCH + k], v[j], s);
        }
        out[i] = s;
    }
}

__global__ void __launch_bounds__(PD_THREADS)
mincurv_pdip_kernel(int B, int n_max, const int32_t *__restrict__ n_pts, double *__restrict__ ws, Layout L,
                    PdipParams prm, double *__restrict__ alpha_out, int32_t *__restrict__ status,
                    int32_t *__restrict__ iters_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PdShared &sh = *reinterpret_cast<PdShared *>(smem_raw);

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const int n = n_pts ? n_pts[b] : n_max;
        double *aout = alpha_out + (size_t)b * n_max;
        __syncthreads();
        if (status[b] != 0) {
            for (int i = threadIdx.x; i < n_max; i += PD_THREADS) aout[i] = 0.0;
            if (iters_out && threadIdx.x == 0) iters_out[b] = 0;
            continue;
        }
        double *slab = ws + (size_t)b * L.stride;
        const double *HB = slab + L.o_hb;
        double *tiles = slab + L.o_tiles;
        const double *LB = vec(slab, L, V_LB), *UB = vec(slab, L, V_UB), *F = vec(slab, L, V_F);
        double *AL = vec(slab, L, V_ALPHA), *LU = vec(slab, L, V_LU), *LL = vec(slab, L, V_LL), *RD = vec(slab, L, V_RD);
        double *RHS = vec(slab, L, V_RHS), *DX = vec(slab, L, V_DX), *DD = vec(slab, L, V_DD);
        double *TU = vec(slab, L, V_DLU), *TL = vec(slab, L, V_DLL), *SU = vec(slab, L, V_SU), *SL = vec(slab, L, V_SL);
        double *YP = vec(slab, L, V_T4), *ZP = vec(slab, L, V_T5), *G0 = vec(slab, L, V_T0);
        const int nb = (n - 32 + 31) / 32;
        if (threadIdx.x == 0) sh.flag = 0;

        // ---------------- initial point: box centre, multipliers from the gradient ----------------
        for (int i = threadIdx.x; i < n; i += PD_THREADS) AL[i] = 0.5 * (LB[i] + UB[i]);
        __syncthreads();
        band_matvec(HB, AL, G0, n);
        __syncthreads();
        double gmax = 0.0, fmaxv = 0.0;
        for (int i = threadIdx.x; i < n; i += PD_THREADS) {
            const double gi = G0[i] + F[i];
            G0[i] = gi;
            gmax = fmax(gmax, fabs(gi));
            fmaxv = fmax(fmaxv, fabs(F[i]));
        }
        gmax = block_reduce<1>(gmax, sh.red);
        fmaxv = block_reduce<1>(fmaxv, sh.red);
        const double lam0 = 1e-2 * gmax + 1e-300;
        double musum = 0.0;
        for (int i = threadIdx.x; i < n; i += PD_THREADS) {
            const double gi = G0[i];
            const double lu = fmax(-gi, 0.0) + lam0, ll = fmax(gi, 0.0) + lam0;
            LU[i] = lu; LL[i] = ll;
            RD[i] = gi + lu - ll;
            const double a = AL[i];
            // slacks are carried as variables of their own: recomputing ub - alpha loses them to
            // cancellation once s << eps |alpha| (late iterations), see DESIGN.md
            const double su = UB[i] - a, sl = a - LB[i];
            SU[i] = su; SL[i] = sl;
            musum += su * lu + sl * ll;
        }
        musum = block_reduce<0>(musum, sh.red);
        const double mu0 = musum / (2.0 * n);
        const double rd_tol = prm.rd_rel * (fmaxv + gmax) + 1e-300;
        double mu = mu0;
        int it = 0;
        int result = 2;   // max-iter unless we converge

        for (it = 0; it < prm.max_iter; ++it) {
            // ---- barrier diagonal and affine right-hand side ----
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double su = SU[i], sl = SL[i], lu = LU[i], ll = LL[i];
                DD[i] = lu / su + ll / sl;
                RHS[i] = -RD[i] + lu - ll;
            }
            __syncthreads();
            if (!factor(sh, HB, DD, tiles, n, nb)) { result = 3; break; }
            solve(sh, tiles, RHS, DX, YP, ZP, n, nb);
            // ---- affine step lengths, centring parameter ----
            double ap = 1.0, ad = 1.0;
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double su = SU[i], sl = SL[i], lu = LU[i], ll = LL[i], dx = DX[i];
                const double dlu = -lu + lu * dx / su, dll = -ll - ll * dx / sl;
                if (dx > 0.0) ap = fmin(ap, su / dx);
                if (dx < 0.0) ap = fmin(ap, -sl / dx);
                if (dlu < 0.0) ad = fmin(ad, -lu / dlu);
                if (dll < 0.0) ad = fmin(ad, -ll / dll);
            }
            ap = block_reduce<2>(ap, sh.red);
            ad = block_reduce<2>(ad, sh.red);
            double mua = 0.0;
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double su = SU[i], sl = SL[i], lu = LU[i], ll = LL[i], dx = DX[i];
                const double dlu = -lu + lu * dx / su, dll = -ll - ll * dx / sl;
                mua += (su - ap * dx) * (lu + ad * dlu) + (sl + ap * dx) * (ll + ad * dll);
            }
            mua = block_reduce<0>(mua, sh.red) / (2.0 * n);
            double sigma = mua / mu;
            sigma = sigma * sigma * sigma;
            const double smu = sigma * mu;
            // ---- corrector right-hand side ----
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double su = SU[i], sl = SL[i], lu = LU[i], ll = LL[i], dx = DX[i];
                const double dlu = -lu + lu * dx / su, dll = -ll - ll * dx / sl;
                const double tu = smu - su * lu - (-dx) * dlu;
                const double tl = smu - sl * ll - dx * dll;
                TU[i] = tu; TL[i] = tl;
                RHS[i] = -RD[i] - tu / su + tl / sl;
            }
            __syncthreads();
            solve(sh, tiles, RHS, DX, YP, ZP, n, nb);
            ap = 1e300; ad = 1e300;
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double su = SU[i], sl = SL[i], lu = LU[i], ll = LL[i], dx = DX[i];
                const double dlu = (TU[i] + lu * dx) / su, dll = (TL[i] - ll * dx) / sl;
                if (dx > 0.0) ap = fmin(ap, su / dx);
                if (dx < 0.0) ap = fmin(ap, -sl / dx);
                if (dlu < 0.0) ad = fmin(ad, -lu / dlu);
                if (dll < 0.0) ad = fmin(ad, -ll / dll);
            }
            ap = fmin(1.0, prm.eta * block_reduce<2>(ap, sh.red));
            ad = fmin(1.0, prm.eta * block_reduce<2>(ad, sh.red));
            double musum2 = 0.0, rdmax = 0.0;
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double su = SU[i], sl = SL[i], lu = LU[i], ll = LL[i], dx = DX[i];
                const double dlu = (TU[i] + lu * dx) / su, dll = (TL[i] - ll * dx) / sl;
                const double an = AL[i] + ap * dx, lun = lu + ad * dlu, lln = ll + ad * dll;
                const double sun = su - ap * dx, sln = sl + ap * dx;
                // H dx = rhs - D dx  (M dx = rhs)
                const double rdn = RD[i] + ap * (RHS[i] - DD[i] * dx) + ad * (dlu - dll);
                AL[i] = an; LU[i] = lun; LL[i] = lln; RD[i] = rdn; SU[i] = sun; SL[i] = sln;
                musum2 += sun * lun + sln * lln;
                rdmax = fmax(rdmax, fabs(rdn));
            }
            mu = block_reduce<0>(musum2, sh.red) / (2.0 * n);
            rdmax = block_reduce<1>(rdmax, sh.red);
            if (mu <= prm.mu_rel * mu0 && rdmax <= rd_tol) { result = 0; ++it; break; }
            if (mu <= 1e-4 * prm.mu_rel * mu0) { result = (rdmax <= 1e3 * rd_tol) ? 0 : 2; ++it; break; }   // complementarity exhausted
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_max; i += PD_THREADS) aout[i] = (i < n) ? AL[i] : 0.0;
        if (threadIdx.x == 0) {
            status[b] = result;
            if (iters_out) iters_out[b] = it;
        }
    }
}

size_t pdip_smem_bytes() { return sizeof(PdShared); }

int launch_mincurv_pdip(int B, int n_max, const int32_t *n_pts, double *ws, const Layout &L, const PdipParams &prm,
                        double *alpha, int32_t *status, int32_t *iters, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mincurv_pdip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(PdShared));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    mincurv_pdip_kernel<<<grid, PD_THREADS, sizeof(PdShared), stream>>>(B, n_max, n_pts, ws, L, prm, alpha, status, iters);
    return 0;
}

}  // namespace mc
