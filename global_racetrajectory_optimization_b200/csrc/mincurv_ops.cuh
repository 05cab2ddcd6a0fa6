// Block-cooperative device pieces shared by the assembly kernel (K2a) and the curvature-row phase of the
// interior-point kernel (K2b'): the weighted band assembly  H_w = E^T diag(1 + w) E  in O(N b) and the O(N)
// operator forms of E and E^T (see mincurv_setup.cu for the derivation).  All vectors live in the instance slab.
#pragma once
#include "mincurv_ws.cuh"

namespace mc {

constexpr int BD = 34;               // B_t band: offsets d = c' - c in [0, 34]
constexpr int BB_T = BD + 2;         // 35 doubles per weight, pitch 36: runs of four offsets are 32 B sectors
static_assert(3 * BB_T <= ZB_PITCH, "B band does not fit the slab pitch");
constexpr int STG_T = 32 * 5;         // per-warp store staging: [row][4 offsets], padded to 5
__host__ __device__ constexpr int hband_win_doubles(int threads) { return 8 * (threads + BD + 1) + (threads / 32) * 3 * STG_T; }


// Band of H_w = E^T diag(1 + wk) E into the slab's HB (wk == nullptr: plain H = E^T E).
// Uses V_T0..V_T5 and the B-band scratch; win: shared scratch of hband_win_doubles(blockDim.x) doubles.
// Ends with the band written but NOT synchronised.
__device__ inline void assemble_hband(double *slab, const Layout &L, int n, const double *__restrict__ wk, double *win) {
    const double *IH = vec(slab, L, V_IH), *TII = vec(slab, L, V_TII), *RHOP = vec(slab, L, V_RHOP), *RHOM = vec(slab, L, V_RHOM);
    const double *NX = vec(slab, L, V_NX), *NY = vec(slab, L, V_NY), *SX = vec(slab, L, V_SX), *SY = vec(slab, L, V_SY);
    // ---- P6: tail sums U_t (towards +), V_t (towards -) of the three weights, chunked with warm-up ----
    double *U0 = vec(slab, L, V_T0), *U1 = vec(slab, L, V_T1), *U2 = vec(slab, L, V_T2);
    double *V0 = vec(slab, L, V_T3), *V1 = vec(slab, L, V_T4), *V2 = vec(slab, L, V_T5);
    for (int c0 = threadIdx.x * TRI_CHUNK; c0 < n; c0 += blockDim.x * TRI_CHUNK) {
        const int c1 = min(c0 + TRI_CHUNK, n);
        int i = wrapi(c0 - TRI_WARM, n);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, rprev = 0.0;       // V_c = w_c + rho-_{c-1}^2 V_{c-1}
#pragma unroll 8
        for (int s = 0; s < TRI_WARM; ++s) {                     // warm-up: no stores, loads of eight steps in flight
            const double ww = wk ? 1.0 + wk[i] : 1.0;
            const double sx = SX[i], sy = ww * SY[i], r2 = rprev * rprev;
            a0 = fma(r2, a0, sy * SY[i]); a1 = fma(r2, a1, sx * sy); a2 = fma(r2, a2, ww * sx * sx);
            rprev = RHOM[i];
            i = (i + 1 == n) ? 0 : i + 1;
        }
        for (int s = c0; s < c1; ++s) {
            const double ww = wk ? 1.0 + wk[i] : 1.0;
            const double sx = SX[i], sy = ww * SY[i], r2 = rprev * rprev;
            a0 = fma(r2, a0, sy * SY[i]); a1 = fma(r2, a1, sx * sy); a2 = fma(r2, a2, ww * sx * sx);
            V0[i] = a0; V1[i] = a1; V2[i] = a2;
            rprev = RHOM[i];
            i = (i + 1 == n) ? 0 : i + 1;
        }
        i = wrapi(c1 - 1 + TRI_WARM, n);
        a0 = a1 = a2 = 0.0;
        double rnext = 0.0;                                        // U_c = w_c + rho+_{c+1}^2 U_{c+1}
#pragma unroll 8
        for (int s = 0; s < TRI_WARM; ++s) {
            const double ww = wk ? 1.0 + wk[i] : 1.0;
            const double sx = SX[i], sy = ww * SY[i], r2 = rnext * rnext;
            a0 = fma(r2, a0, sy * SY[i]); a1 = fma(r2, a1, sx * sy); a2 = fma(r2, a2, ww * sx * sx);
            rnext = RHOP[i];
            i = (i == 0) ? n - 1 : i - 1;
        }
        for (int s = c1 - 1; s >= c0; --s) {
            const double ww = wk ? 1.0 + wk[i] : 1.0;
            const double sx = SX[i], sy = ww * SY[i], r2 = rnext * rnext;
            a0 = fma(r2, a0, sy * SY[i]); a1 = fma(r2, a1, sx * sy); a2 = fma(r2, a2, ww * sx * sx);
            U0[i] = a0; U1[i] = a1; U2[i] = a2;
            rnext = RHOP[i];
            i = (i == 0) ? n - 1 : i - 1;
        }
    }
    __syncthreads();
    // ---- P7: band of B_t[c][c + d], d = 0..34.  One thread per row c, a serial recurrence along d that reads
    //      eight vectors at c + d: a sliding window, staged in shared memory per chunk of blockDim.x rows (from
    //      global memory the 34 dependent steps ran at L2 latency and were 56 % of the assembly kernel) ----
    double *BB = slab + L.o_zb;
    {
        const int CH = blockDim.x, WL = CH + BD + 1;
        double *w_rp = win, *w_ti = win + WL, *w_u0 = win + 2 * WL, *w_u1 = win + 3 * WL, *w_u2 = win + 4 * WL;
        double *w_yy = win + 5 * WL, *w_xy = win + 6 * WL, *w_xx = win + 7 * WL;     // ww sy^2 t, ww sx sy t, ww sx^2 t
        for (int cbase = 0; cbase < n; cbase += CH) {
            __syncthreads();
            for (int e = threadIdx.x; e < WL; e += blockDim.x) {
                const int idx = (cbase + e) % n;
                const double ww = wk ? 1.0 + wk[idx] : 1.0;
                const double sx = SX[idx], sy = SY[idx], ti = TII[idx], wt = ww * ti;
                w_rp[e] = RHOP[idx]; w_ti[e] = ti; w_u0[e] = U0[idx]; w_u1[e] = U1[idx]; w_u2[e] = U2[idx];
                w_yy[e] = sy * sy * wt; w_xy[e] = sx * sy * wt; w_xx[e] = sx * sx * wt;
            }
            __syncthreads();
            // every lane runs the recurrence (rows past n compute on window data and are masked at the store);
            // the results leave through a per-warp staging tile so that a store instruction writes 8 rows x 32 B
            // (8 sectors) instead of 32 rows x 8 B (32 sectors): the direct stores were LSU-tag bound
            const int c = cbase + threadIdx.x, l = threadIdx.x, lane = threadIdx.x & 31;
            double *stg = win + 8 * WL + (threadIdx.x >> 5) * 3 * STG_T;
            const int cw = cbase + (threadIdx.x & ~31);           // first row of this warp
            const double tc = w_ti[l];
            const double v0 = (c < n) ? V0[c] : 0.0, v1 = (c < n) ? V1[c] : 0.0, v2 = (c < n) ? V2[c] : 0.0;
            stg[lane * 5] = tc * (tc * (w_u0[l] + v0) - w_yy[l]);
            stg[STG_T + lane * 5] = tc * (tc * (w_u1[l] + v1) - w_xy[l]);
            stg[2 * STG_T + lane * 5] = tc * (tc * (w_u2[l] + v2) - w_xx[l]);
            double P = tc, m0 = 0.0, m1 = 0.0, m2 = 0.0;
#pragma unroll 1
            for (int d0 = 0; d0 <= BD; d0 += 4) {
#pragma unroll
                for (int s = (d0 == 0) ? 1 : 0; s < 4; ++s) {
                    const int d = d0 + s;
                    if (d <= BD) {
                        const double rp = w_rp[l + d];
                        if (d >= 2) {
                            m0 = rp * fma(w_yy[l + d - 1], P, m0);
                            m1 = rp * fma(w_xy[l + d - 1], P, m1);
                            m2 = rp * fma(w_xx[l + d - 1], P, m2);
                        }
                        P *= rp;
                        const double tp = w_ti[l + d];
                        stg[lane * 5 + s] = P * fma(tp, w_u0[l + d], tc * v0) + m0;
                        stg[STG_T + lane * 5 + s] = P * fma(tp, w_u1[l + d], tc * v1) + m1;
                        stg[2 * STG_T + lane * 5 + s] = P * fma(tp, w_u2[l + d], tc * v2) + m2;
                    }
                }
                __syncwarp();
                const int s = lane & 3;
                if (d0 + s <= BD) {
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int r = (lane >> 2) + 8 * k;
                            if (cw + r < n) BB[(size_t)(cw + r) * ZB_PITCH + t * BB_T + d0 + s] = stg[t * STG_T + r * 5 + s];
                        }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    // ---- P8: band of H: HB[i][k] = H[i][i + k], k = 0..32 (9-point stencil on the three B bands) ----
    double *HB = slab + L.o_hb;
    const int tot = n * (HBW + 1);
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        const int i = e / (HBW + 1), k = e - i * (HBW + 1);
        int j = i + k; if (j >= n) j -= n;
        const int im1 = (i == 0) ? n - 1 : i - 1, jm1 = (j == 0) ? n - 1 : j - 1;
        const double ihi = IH[i], ihim = IH[im1], ihj = IH[j], ihjm = IH[jm1];
        const double ei[3] = {ihim, -(ihim + ihi), ihi};       // 6 D2[c][i] / 6, c = i-1, i, i+1
        const double ej[3] = {ihjm, -(ihjm + ihj), ihj};
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int dc = -1; dc <= 1; ++dc) {
            int c = i + dc; c = (c < 0) ? c + n : ((c >= n) ? c - n : c);
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                int cp = j + dj; cp = (cp < 0) ? cp + n : ((cp >= n) ? cp - n : cp);
                const int d = k + dj - dc;                       // c' - c  (in [-2, 34])
                const double *row = (d >= 0) ? BB + (size_t)c * ZB_PITCH + d : BB + (size_t)cp * ZB_PITCH - d;
                const double co = ei[dc + 1] * ej[dj + 1];
                a0 = fma(co, row[0], a0);
                a1 = fma(co, row[BB_T], a1);
                a2 = fma(co, row[2 * BB_T], a2);
            }
        }
        const double nyi = NY[i], nxi = NX[i], nyj = NY[j], nxj = NX[j];
        HB[(size_t)i * HB_PITCH + k] = 36.0 * (nyi * nyj * a0 - (nyi * nxj + nxi * nyj) * a1 + nxi * nxj * a2);
    }
}

// out = E v = S_y Z (n_y v) - S_x Z (n_x v)   (r0, r1, y0, y1, z0, z1: scratch vectors; out may alias none of them)
__device__ inline void apply_E(double *slab, const Layout &L, int n, const double *__restrict__ v, double *__restrict__ out,
                               double *r0, double *r1, double *y0, double *y1, double *z0, double *z1) {
    const double *H = vec(slab, L, V_H), *LFW = vec(slab, L, V_LFW), *INVD = vec(slab, L, V_INVD);
    const double *NX = vec(slab, L, V_NX), *NY = vec(slab, L, V_NY), *SX = vec(slab, L, V_SX), *SY = vec(slab, L, V_SY);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int im1 = (i == 0) ? n - 1 : i - 1, ip1 = (i + 1 == n) ? 0 : i + 1;
        const double hi = H[i], him = H[im1];
        const double vx = NX[i] * v[i], vxm = NX[im1] * v[im1], vxp = NX[ip1] * v[ip1];
        const double vy = NY[i] * v[i], vym = NY[im1] * v[im1], vyp = NY[ip1] * v[ip1];
        r0[i] = 6.0 * ((vxp - vx) / hi - (vx - vxm) / him);
        r1[i] = 6.0 * ((vyp - vy) / hi - (vy - vym) / him);
    }
    __syncthreads();
    tri_solve2(LFW, INVD, H, r0, r1, y0, y1, z0, z1, n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = SY[i] * z1[i] - SX[i] * z0[i];
    __syncthreads();
}

// out = E^T v = N_y Z^T (S_y v) - N_x Z^T (S_x v),  Z^T = 6 D2 Tri^{-1}
__device__ inline void apply_Et(double *slab, const Layout &L, int n, const double *__restrict__ v, double *__restrict__ out,
                                double *r0, double *r1, double *y0, double *y1, double *z0, double *z1) {
    const double *H = vec(slab, L, V_H), *LFW = vec(slab, L, V_LFW), *INVD = vec(slab, L, V_INVD);
    const double *NX = vec(slab, L, V_NX), *NY = vec(slab, L, V_NY), *SX = vec(slab, L, V_SX), *SY = vec(slab, L, V_SY);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { r0[i] = SX[i] * v[i]; r1[i] = SY[i] * v[i]; }
    __syncthreads();
    tri_solve2(LFW, INVD, H, r0, r1, y0, y1, z0, z1, n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int im1 = (i == 0) ? n - 1 : i - 1, ip1 = (i + 1 == n) ? 0 : i + 1;
        const double hi = H[i], him = H[im1];
        const double zx = 6.0 * ((z0[ip1] - z0[i]) / hi - (z0[i] - z0[im1]) / him);
        const double zy = 6.0 * ((z1[ip1] - z1[i]) / hi - (z1[i] - z1[im1]) / him);
        out[i] = NY[i] * zy - NX[i] * zx;
    }
    __syncthreads();
}

}  // namespace mc
