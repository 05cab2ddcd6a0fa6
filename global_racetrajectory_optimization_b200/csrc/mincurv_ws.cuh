// Per-instance scratch layout of the minimum-curvature path (all float64, one slab per QP instance).
//
// HBM layout: slab b starts at ws + b * stride doubles and holds
//   [NUM_VEC][np]        O(N) vectors (geometry, linearisation, interior-point iterates)
//   [n_max][ZB_PITCH]    bands of B_t = Ti diag(w_t) Ti, t = 0..2  (assembly scratch, mincurv_setup.cu)
//   [np][HB_PITCH]       band of H = E^T E              (row i: H[i][i .. i+32], cyclic)
//   [np][42] + [np][34]  bordered LDL^T factor of H + D: chain rows [Q - I | L21 | - | w] per panel of eight columns,
//                        fill rows [G | z | w] (mincurv_ipm.cu)
#pragma once
#include "common.cuh"

namespace mc {

enum Vec : int {
    V_H = 0, V_DIAG, V_DFW, V_DBW, V_LFW, V_INVD, V_TII, V_RHOP, V_RHOM,
    V_PX, V_PY, V_NX, V_NY, V_MX, V_MY, V_XP, V_YP, V_SX, V_SY, V_KREF,
    V_LB, V_UB, V_F,
    V_T0, V_T1, V_T2, V_T3, V_T4, V_T5,
    V_ALPHA, V_LU, V_LL, V_RD, V_RHS, V_DX, V_DD, V_DLU, V_DLL, V_SU, V_SL,
    V_ISU, V_ISL, V_YPAD,                                   // reciprocal slacks, padded forward-solve vector
    V_S3, V_S4, V_L3, V_L4, V_KL, V_WK, V_EDX, V_T3K, V_T4K, V_VV,   // curvature-row phase (K2b')
    V_IH,                                                   // 1 / h
    NUM_VEC
};

struct Layout {
    int n_max;
    int np;          // padded vector length (multiple of 32, >= n_max + 64)
    int nb_max;      // (n_max - 32) / 32 rounded up (kept for the Python mirror of the layout)
    size_t o_zb, o_hb, o_tiles, stride;   // in doubles
};

__host__ __device__ inline Layout make_layout(int n_max) {
    Layout L;
    L.n_max = n_max;
    L.np = ((n_max + 31) / 32) * 32 + 64;
    L.nb_max = (n_max - 32 + 31) / 32;
    if (L.nb_max < 1) L.nb_max = 1;
    size_t o = (size_t)NUM_VEC * L.np;
    L.o_zb = o;
    o += (size_t)n_max * ZB_PITCH;
    L.o_hb = o;
    o += (size_t)L.np * HB_PITCH;
    L.o_tiles = o;
    o += (size_t)L.np * 76;
    L.stride = (o + 15) & ~(size_t)15;
    return L;
}

struct PdipParams {
    int max_iter;
    double mu_rel;     // stop when mu <= mu_rel * mu0 ...
    double rd_rel;     // ... and |r_d|_inf <= rd_rel * (|f|_inf + |g0|_inf)
    double eta;        // fraction to the boundary
    double dx_rel;     // ... and the last step moved alpha by <= dx_rel * max(|alpha|_inf, 0.01 m)  (0: not checked)
    double lam0_rel;   // initial multipliers: max(-+g, 0) + lam0_rel * |g|_inf
};

__device__ __forceinline__ double *vec(double *slab, const Layout &L, int v) { return slab + (size_t)v * L.np; }

}  // namespace mc
