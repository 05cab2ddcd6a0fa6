// K2b -- box-constrained QP solve  min 1/2 a^T H a + f^T a,  lb <= a <= ub  by a Mehrotra
// predictor-corrector primal-dual interior-point method (replaces quadprog.solve_qp inside
// tph.opt_min_curv, call site /root/reference/main_globaltraj.py:264-271).
//
// One CTA of four warps per QP instance.  H is the cyclic band (half-bandwidth 32) assembled by K2a.
// Every interior-point iteration factorises M = H + D (D diagonal, from the barrier) with a
// block-cyclic Cholesky on 32x32 blocks:
//   chain blocks I = 0..nb-1 (nodes 0..n-33, padded), separator S = last 32 nodes (closes the cycle)
//     A'_I   = A_I + D_I - T_I T_I^T                 T_I   = L_{I,I-1}      (upper triangular)
//     L_II   = chol(A'_I),  Linv_I = L_II^{-1}       (explicit inverse: sweeps become mat-vecs)
//     T_{I+1}= B_I Linv_I^T                          B_I   = M[block I+1, block I]
//     F_I    = (Y_I - F_{I-1} T_I^T) Linv_I^T        Y_I   = M[S, block I]   (fill row of the separator)
//     S     -= F_I F_I^T
// The 32x32x32 block products are dense fp64 contractions and run on the FP64 tensor cores
// (mma.sync.m8n8k4.f64, SASS DMMA) on 8x8 sub-blocks, skipping the sub-blocks that the triangular
// structure of T / Linv makes zero (40 instead of 128 DMMA for the chain products).  The sequential parts
// (32x32 Cholesky, triangular inverse) are rolled loops over shared-memory tiles, which keeps the whole
// kernel inside the instruction cache (the first, fully unrolled version of this kernel was 442 KB of
// SASS and instruction-fetch bound: profiles/r01_v1_pdip_ncu_summary.json).
// warp 0 owns the chain (A', chol, inverse, T); warps 1/2 own the two 16-row halves of the separator row
// (F, S); warp 3 stages the band rows of the next block.  The factor goes to the instance's HBM slab as one
// packed 32x33 tile per block (Linv_I lower | T_I upper) plus one 32x32 tile F_I; the triangular sweeps
// are sequences of 32x32 mat-vecs over them.
#include "mincurv_ws.cuh"

namespace mc {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PD_THREADS = 128;
constexpr int TP = 36;             // shared tile pitch (doubles): conflict-free DMMA fragment loads
constexpr int LTP = 33;            // pitch of the packed (Linv | T) tile in HBM
constexpr int LT_TILE = 32 * LTP;  // 1056 doubles
constexpr int F_TILE = 32 * LTP;   // F_I, row-major pitch 33 (odd pitch: conflict-free lane=row and lane=column access)
constexpr int BLK_TILES = LT_TILE + F_TILE;   // doubles per chain block in the slab (one contiguous bulk copy)
constexpr unsigned BLK_BYTES = BLK_TILES * sizeof(double);
constexpr unsigned LT_BYTES = LT_TILE * sizeof(double);

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// ---- TMA (bulk async copy) + mbarrier helpers: the triangular sweeps stream the factor tiles HBM -> shared ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// entry M[i][j] of the cyclic band (real node indices), 0 outside the band
__device__ __forceinline__ double hentry(const double *HB, int n, int i, int j) {
    int k = j - i;
    if (k < 0) k += n;
    if (k <= HBW) return HB[(size_t)i * HB_PITCH + k];
    k = n - k;
    if (k <= HBW) return HB[(size_t)j * HB_PITCH + k];
    return 0.0;
}

struct PdShared {
    double band[2][32 * HB_PITCH];   // band rows of the current / next chain block
    // The five factorisation tiles; during the triangular sweeps the same 46 KB hold the two 16.5 KB
    // staging buffers of the TMA tile pipeline (stage s at &As[0] + s * BLK_TILES).
    double As[32 * TP];              // A'_I -> L_II, column-major: As[c * TP + r]
    double Li[32 * TP];              // Linv_I, row-major
    double Ts[32 * TP];              // T_I, row-major (upper triangular)
    double Fa[32 * TP];              // F tiles, ping
    double Fb[32 * TP];              // pong
    uint64_t full_bar[2], empty_bar[2];
    double dinv[32];
    double vbuf[8][32];
    double red[32];
    int flag;
};

// 32x32 Cholesky, lane = row, tile column-major in shared (As[c*TP + r]); in place, lower triangle only.
// Left-looking over four 8-column panels: the contributions of the finished columns are subtracted with
// eight independent accumulators (rolled loop, two broadcast LDS.128 + one LDS per step), then the 8-column
// panel is factored in registers with warp shuffles.  Also writes dinv[r] = 1 / L[r][r].
__device__ __forceinline__ bool chol32_smem(double *As, double *dinv, int lane) {
    bool ok = true;
#pragma unroll 1
    for (int J = 0; J < 4; ++J) {
        const int c0 = 8 * J;
        double a[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) a[t] = As[(c0 + t) * TP + lane];
#pragma unroll 2
        for (int j = 0; j < c0; ++j) {
            const double lr = -As[j * TP + lane];
            const double2 b0 = *reinterpret_cast<const double2 *>(&As[j * TP + c0]);
            const double2 b1 = *reinterpret_cast<const double2 *>(&As[j * TP + c0 + 2]);
            const double2 b2 = *reinterpret_cast<const double2 *>(&As[j * TP + c0 + 4]);
            const double2 b3 = *reinterpret_cast<const double2 *>(&As[j * TP + c0 + 6]);
            a[0] = fma(lr, b0.x, a[0]); a[1] = fma(lr, b0.y, a[1]);
            a[2] = fma(lr, b1.x, a[2]); a[3] = fma(lr, b1.y, a[3]);
            a[4] = fma(lr, b2.x, a[4]); a[5] = fma(lr, b2.y, a[5]);
            a[6] = fma(lr, b3.x, a[6]); a[7] = fma(lr, b3.y, a[7]);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const double d = __shfl_sync(FULL, a[t], c0 + t);
            if (!(d > 0.0)) ok = false;
            const double rs = rsqrt(d);
            const double l = a[t] * rs;
            a[t] = l;
#pragma unroll
            for (int u = t + 1; u < 8; ++u) a[u] = fma(-l, __shfl_sync(FULL, l, c0 + u), a[u]);
            if (lane == c0 + t) dinv[c0 + t] = rs;
        }
#pragma unroll
        for (int t = 0; t < 8; ++t)
            if (lane >= c0 + t) As[(c0 + t) * TP + lane] = a[t];
        __syncwarp();
    }
    return ok;
}

// Linv = L^{-1}: lane = column c; L column-major in As, Linv row-major in Li (full tile written, zeros above the
// diagonal).  Four 8-row panels: contributions of the finished rows with eight independent accumulators, then
// the 8x8 triangular solve of the panel.
__device__ __forceinline__ void trinv32_smem(const double *As, const double *dinv, double *Li, int lane) {
#pragma unroll 1
    for (int R = 0; R < 4; ++R) {
        const int r0 = 8 * R;
        double sacc[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) sacc[t] = (r0 + t == lane) ? 1.0 : 0.0;
#pragma unroll 2
        for (int k = 0; k < r0; ++k) {
            const double x = -Li[k * TP + lane];
            const double2 b0 = *reinterpret_cast<const double2 *>(&As[k * TP + r0]);
            const double2 b1 = *reinterpret_cast<const double2 *>(&As[k * TP + r0 + 2]);
            const double2 b2 = *reinterpret_cast<const double2 *>(&As[k * TP + r0 + 4]);
            const double2 b3 = *reinterpret_cast<const double2 *>(&As[k * TP + r0 + 6]);
            sacc[0] = fma(b0.x, x, sacc[0]); sacc[1] = fma(b0.y, x, sacc[1]);
            sacc[2] = fma(b1.x, x, sacc[2]); sacc[3] = fma(b1.y, x, sacc[3]);
            sacc[4] = fma(b2.x, x, sacc[4]); sacc[5] = fma(b2.y, x, sacc[5]);
            sacc[6] = fma(b3.x, x, sacc[6]); sacc[7] = fma(b3.y, x, sacc[7]);
        }
        double xv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            double acc = sacc[t];
#pragma unroll
            for (int u = 0; u < t; ++u) acc = fma(-As[(r0 + u) * TP + r0 + t], xv[u], acc);
            xv[t] = acc * dinv[r0 + t];
            Li[(r0 + t) * TP + lane] = xv[t];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// factorisation of M = H + diag(DD); returns false on a non-positive pivot
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ bool factor(PdShared &sh, const double *__restrict__ HB, const double *__restrict__ DD,
                       double *__restrict__ tiles, int n, int nb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int NA = n - 32;
    bool ok = true;
    double sacc[2][4][2];          // warps 1,2: their 16 rows of the separator Schur complement (C fragments)
    double *Fcur = sh.Fa, *Fnxt = sh.Fb;

    // separator diagonal block C + D_S: gathered into As by all threads, then picked up as C fragments
#pragma unroll 1
    for (int e = threadIdx.x; e < 1024; e += PD_THREADS) {
        const int r = e >> 5, c = e & 31;
        double v = hentry(HB, n, NA + r, NA + c);
        if (r == c) v += DD[NA + r];
        sh.As[c * TP + r] = v;
    }
    __syncthreads();
    if (warp == 1 || warp == 2) {
        const int ib = 2 * (warp - 1);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) sacc[i][j][e] = sh.As[(8 * j + 2 * q + e) * TP + 8 * (ib + i) + g];
    }
    __syncthreads();
    // stage the band rows of block 0
    for (int e = threadIdx.x; e < 32 * HB_PITCH; e += PD_THREADS) sh.band[0][e] = HB[e];
    __syncthreads();

    for (int I = 0; I < nb; ++I) {
        const int base = 32 * I;
        const double *band = sh.band[I & 1];
        // =========================== phase A ===========================
        if (warp == 0) {
            // ---- A' = A_I + D_I - T_I T_I^T (lower 8x8 blocks), written column-major into As ----
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    double c2[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int r = 8 * i + g, c = 8 * j + 2 * q + e;
                        const int lo = (c < r) ? c : r, dist = (c < r) ? r - c : c - r;
                        double v = band[lo * HB_PITCH + dist];
                        const bool real = (base + r) < NA && (base + c) < NA;
                        if (!real) v = (r == c) ? 1.0 : 0.0;
                        else if (r == c) v += DD[base + r];
                        c2[e] = v;
                    }
                    if (I > 0) {
                        double m2[2] = {0.0, 0.0};
#pragma unroll
                        for (int K = i; K < 4; ++K)      // T upper triangular: blocks (i,K), (j,K) nonzero for K >= i >= j
#pragma unroll
                            for (int s = 0; s < 2; ++s)
                                dmma(m2, sh.Ts[(8 * i + g) * TP + 8 * K + 4 * s + q], sh.Ts[(8 * j + g) * TP + 8 * K + 4 * s + q]);
                        c2[0] -= m2[0];
                        c2[1] -= m2[1];
                    }
                    sh.As[(8 * j + 2 * q) * TP + 8 * i + g] = c2[0];
                    sh.As[(8 * j + 2 * q + 1) * TP + 8 * i + g] = c2[1];
                }
            }
            __syncwarp();
            ok = chol32_smem(sh.As, sh.dinv, lane) && ok;
            trinv32_smem(sh.As, sh.dinv, sh.Li, lane);
        } else if (warp == 3) {
            if (I + 1 < nb) {      // stage the band rows of the next block
                double *dst = sh.band[(I + 1) & 1];
                const double *src = HB + (size_t)(base + 32) * HB_PITCH;
                for (int e = lane; e < 32 * HB_PITCH; e += 32) dst[e] = src[e];
            }
        } else {
            const int ib = 2 * (warp - 1);
            // ---- S -= F_{I-1} F_{I-1}^T (own 16 rows) ----
            if (I > 0) {
#pragma unroll
                for (int K = 0; K < 4; ++K)
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        double a[2], b[4];
#pragma unroll
                        for (int i = 0; i < 2; ++i) a[i] = -Fcur[(8 * (ib + i) + g) * TP + 8 * K + 4 * s + q];
#pragma unroll
                        for (int j = 0; j < 4; ++j) b[j] = Fcur[(8 * j + g) * TP + 8 * K + 4 * s + q];
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) dmma(sacc[i][j], a[i], b[j]);
                    }
            }
            // ---- FW = Y_I - F_{I-1} T_I^T (own rows) -> Fnxt ----
            const bool hasY = (I == 0) || (base + 31 >= NA - 32);
            if (hasY) {      // Y_I = M[S, block I] is nonzero only across the wrap (I = 0) and next to the separator
#pragma unroll 1
                for (int e = lane; e < 512; e += 32) {
                    const int r = 16 * (warp - 1) + (e >> 5), c = e & 31;
                    Fnxt[r * TP + c] = ((base + c) < NA) ? hentry(HB, n, NA + r, base + c) : 0.0;
                }
                __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double c2[2] = {0.0, 0.0};
                    if (hasY) {
                        const double2 y2 = *reinterpret_cast<const double2 *>(&Fnxt[(8 * (ib + i) + g) * TP + 8 * j + 2 * q]);
                        c2[0] = y2.x;
                        c2[1] = y2.y;
                    }
                    if (I > 0) {
                        double m2[2] = {0.0, 0.0};
#pragma unroll
                        for (int K = j; K < 4; ++K)      // T[c][k] != 0 for k >= c
#pragma unroll
                            for (int s = 0; s < 2; ++s)
                                dmma(m2, Fcur[(8 * (ib + i) + g) * TP + 8 * K + 4 * s + q], sh.Ts[(8 * j + g) * TP + 8 * K + 4 * s + q]);
                        c2[0] -= m2[0];
                        c2[1] -= m2[1];
                    }
                    __syncwarp();
                    *reinterpret_cast<double2 *>(&Fnxt[(8 * (ib + i) + g) * TP + 8 * j + 2 * q]) = make_double2(c2[0], c2[1]);
                }
        }
        __syncthreads();
        // =========================== phase B ===========================
        if (warp == 0) {
            // Linv_I -> packed HBM tile (lower part, row-major pitch 33)
            double *gt = tiles + (size_t)I * BLK_TILES;
            for (int r = 0; r < 32; ++r)
                if (lane <= r) gt[r * LTP + lane] = sh.Li[r * TP + lane];
            if (I + 1 < nb) {
                // ---- T_{I+1} = B_I Linv_I^T : upper-triangular blocks (i <= j), K = i..j ----
                double *gn = tiles + (size_t)(I + 1) * BLK_TILES;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double c2[2] = {0.0, 0.0};
                        if (j >= i) {
#pragma unroll
                            for (int K = i; K <= j; ++K)
#pragma unroll
                                for (int s = 0; s < 2; ++s) {
                                    // B_I[r][k] = M[base+32+r][base+k] = band[k][32 + r - k] for r <= k
                                    const int r = 8 * i + g, k = 8 * K + 4 * s + q;
                                    double a = 0.0;
                                    if (r <= k && (base + 32 + r) < NA) a = band[k * HB_PITCH + 32 + r - k];
                                    dmma(c2, a, sh.Li[(8 * j + g) * TP + k]);
                                }
                        }
                        const int r = 8 * i + g, c = 8 * j + 2 * q;
                        *reinterpret_cast<double2 *>(&sh.Ts[r * TP + c]) = make_double2(c2[0], c2[1]);
                        if (c >= r) gn[r * LTP + c + 1] = c2[0];
                        if (c + 1 >= r) gn[r * LTP + c + 2] = c2[1];
                    }
            }
        } else if (warp == 1 || warp == 2) {
            const int ib = 2 * (warp - 1);
            // ---- F_I = FW Linv_I^T (own rows; Linv lower triangular: K <= j), in place in Fnxt ----
            double a[2][8];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) a[i][ks] = Fnxt[(8 * (ib + i) + g) * TP + 4 * ks + q];
            __syncwarp();
            double *gf = tiles + (size_t)I * BLK_TILES + LT_TILE;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double c2[2] = {0.0, 0.0};
#pragma unroll
                    for (int K = 0; K <= j; ++K)
#pragma unroll
                        for (int s = 0; s < 2; ++s) dmma(c2, a[i][2 * K + s], sh.Li[(8 * j + g) * TP + 8 * K + 4 * s + q]);
                    const int r = 8 * (ib + i) + g, c = 8 * j + 2 * q;
                    *reinterpret_cast<double2 *>(&Fnxt[r * TP + c]) = make_double2(c2[0], c2[1]);
                    gf[r * LTP + c] = c2[0];
                    gf[r * LTP + c + 1] = c2[1];
                }
        }
        { double *t = Fcur; Fcur = Fnxt; Fnxt = t; }
        __syncthreads();
    }
    // ---- separator: S -= F_{nb-1} F_{nb-1}^T, handed over through As; chol + inverse by warp 0 ----
    if (warp == 1 || warp == 2) {
        const int ib = 2 * (warp - 1);
#pragma unroll
        for (int K = 0; K < 4; ++K)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                double a[2], b[4];
#pragma unroll
                for (int i = 0; i < 2; ++i) a[i] = -Fcur[(8 * (ib + i) + g) * TP + 8 * K + 4 * s + q];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Fcur[(8 * j + g) * TP + 8 * K + 4 * s + q];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma(sacc[i][j], a[i], b[j]);
            }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sh.As[(8 * j + 2 * q) * TP + 8 * (ib + i) + g] = sacc[i][j][0];
                sh.As[(8 * j + 2 * q + 1) * TP + 8 * (ib + i) + g] = sacc[i][j][1];
            }
    }
    __syncthreads();
    if (warp == 0) {
        ok = chol32_smem(sh.As, sh.dinv, lane) && ok;
        trinv32_smem(sh.As, sh.dinv, sh.Li, lane);
        double *gt = tiles + (size_t)nb * BLK_TILES;     // separator inverse: row-major pitch 33 (lower part)
        for (int r = 0; r < 32; ++r)
            if (lane <= r) gt[r * LTP + lane] = sh.Li[r * TP + lane];
        if (!ok) sh.flag = 1;
    }
    __syncthreads();
    return sh.flag == 0;
}

// ---- mat-vecs over one staged block in shared memory: LT = packed (Linv | T), pitch 33; F pitch 33 ----
// (odd pitch: both "lane = row" and "lane = column" access patterns are bank-conflict free; v is broadcast)
__device__ __forceinline__ double sm_t_mv(const double *LT, const double *v, int lane) {     // sum_{c >= r} T[r][c] v[c]
    const double *row = LT + lane * LTP + 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
    for (int c = 0; c < 32; c += 4) {
        s0 = fma((c + 0 >= lane) ? row[c + 0] : 0.0, v[c + 0], s0);
        s1 = fma((c + 1 >= lane) ? row[c + 1] : 0.0, v[c + 1], s1);
        s2 = fma((c + 2 >= lane) ? row[c + 2] : 0.0, v[c + 2], s2);
        s3 = fma((c + 3 >= lane) ? row[c + 3] : 0.0, v[c + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ double sm_linv_mv(const double *LT, const double *v, int lane) {  // sum_{c <= r} Linv[r][c] v[c]
    const double *row = LT + lane * LTP;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
    for (int c = 0; c < 32; c += 4) {
        s0 = fma((c + 0 <= lane) ? row[c + 0] : 0.0, v[c + 0], s0);
        s1 = fma((c + 1 <= lane) ? row[c + 1] : 0.0, v[c + 1], s1);
        s2 = fma((c + 2 <= lane) ? row[c + 2] : 0.0, v[c + 2], s2);
        s3 = fma((c + 3 <= lane) ? row[c + 3] : 0.0, v[c + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ double sm_linv_mtv(const double *LT, const double *v, int lane) { // sum_{r >= c} Linv[r][c] v[r]
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
    for (int r = 0; r < 32; r += 4) {
        s0 = fma((r + 0 >= lane) ? LT[(r + 0) * LTP + lane] : 0.0, v[r + 0], s0);
        s1 = fma((r + 1 >= lane) ? LT[(r + 1) * LTP + lane] : 0.0, v[r + 1], s1);
        s2 = fma((r + 2 >= lane) ? LT[(r + 2) * LTP + lane] : 0.0, v[r + 2], s2);
        s3 = fma((r + 3 >= lane) ? LT[(r + 3) * LTP + lane] : 0.0, v[r + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ double sm_t_mtv(const double *LT, const double *v, int lane) {    // sum_{r <= c} T[r][c] v[r]
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
    for (int r = 0; r < 32; r += 4) {
        s0 = fma((r + 0 <= lane) ? LT[(r + 0) * LTP + lane + 1] : 0.0, v[r + 0], s0);
        s1 = fma((r + 1 <= lane) ? LT[(r + 1) * LTP + lane + 1] : 0.0, v[r + 1], s1);
        s2 = fma((r + 2 <= lane) ? LT[(r + 2) * LTP + lane + 1] : 0.0, v[r + 2], s2);
        s3 = fma((r + 3 <= lane) ? LT[(r + 3) * LTP + lane + 1] : 0.0, v[r + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ double sm_f_mv(const double *F, const double *v, int lane) {      // sum_c F[r][c] v[c]
    const double *row = F + lane * LTP;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
    for (int c = 0; c < 32; c += 4) {
        s0 = fma(row[c + 0], v[c + 0], s0);
        s1 = fma(row[c + 1], v[c + 1], s1);
        s2 = fma(row[c + 2], v[c + 2], s2);
        s3 = fma(row[c + 3], v[c + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ double sm_f_mtv(const double *F, const double *v, int lane) {     // sum_r F[r][c] v[r]
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
    for (int r = 0; r < 32; r += 4) {
        s0 = fma(F[(r + 0) * LTP + lane], v[r + 0], s0);
        s1 = fma(F[(r + 1) * LTP + lane], v[r + 1], s1);
        s2 = fma(F[(r + 2) * LTP + lane], v[r + 2], s2);
        s3 = fma(F[(r + 3) * LTP + lane], v[r + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// ------------------------------------------------------------------------------------------------
// solve M x = g with the stored factor.  g, x: real-indexed vectors (length n) in the slab; ypad: padded scratch.
// The 2 nb + 1 factor blocks [0 .. nb-1, S, nb-1 .. 0] are streamed HBM -> shared by one elected thread of
// warp 3 with cp.async.bulk (TMA, one 16.5 KB copy per block) into a two-stage ring guarded by full/empty
// mbarriers; warp 0 consumes them: per block three 32x32 mat-vecs from shared memory.
// `fill` counts the ring fills of this CTA so far (parity bookkeeping); the new count is returned.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ unsigned solve(PdShared &sh, const double *__restrict__ tiles, const double *__restrict__ g,
                                       double *__restrict__ x, double *__restrict__ ypad, int n, int nb, unsigned fill) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NA = n - 32;
    double *stage0 = sh.As;
    const unsigned nfill = 2u * nb + 1u;
    fence_proxy_async();        // the staging area was last written through the generic proxy (factor tiles)
    __syncthreads();
    if (warp == 3) {
        if (lane == 0) {
            for (unsigned i = 0; i < nfill; ++i) {
                const unsigned f = fill + i, st = f & 1u, k = f >> 1;
                if (k > 0) mbar_wait(&sh.empty_bar[st], (k - 1) & 1u);
                const unsigned blk = (i < (unsigned)nb) ? i : ((i == (unsigned)nb) ? (unsigned)nb : 2u * nb - i);
                const unsigned bytes = (i == (unsigned)nb) ? LT_BYTES : BLK_BYTES;
                mbar_expect_tx(&sh.full_bar[st], bytes);
                tma_load_1d(stage0 + st * BLK_TILES, tiles + (size_t)blk * BLK_TILES, bytes, &sh.full_bar[st]);
            }
        }
    } else if (warp == 0) {
        double *yb = sh.vbuf[0], *tb = sh.vbuf[1], *xs = sh.vbuf[3];
        // ---- forward: y_I = Linv_I (g_I - T_I y_{I-1}),  gS -= F_I y_I ----
        double gacc = 0.0;
        double gnext = (lane < NA) ? g[lane] : 0.0;
        yb[lane] = 0.0;
        __syncwarp();
        for (int I = 0; I < nb; ++I) {
            const unsigned f = fill + I, st = f & 1u, k = f >> 1;
            const double gv = gnext;
            if (I + 1 < nb) { const int nd = 32 * (I + 1) + lane; gnext = (nd < NA) ? g[nd] : 0.0; }
            mbar_wait(&sh.full_bar[st], k & 1u);
            const double *LT = stage0 + st * BLK_TILES, *Ft = LT + LT_TILE;
            double v = gv;
            if (I > 0) v -= sm_t_mv(LT, yb, lane);
            __syncwarp();
            tb[lane] = v;
            __syncwarp();
            const double y = sm_linv_mv(LT, tb, lane);
            yb[lane] = y;
            ypad[32 * I + lane] = y;
            __syncwarp();
            gacc += sm_f_mv(Ft, yb, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty_bar[st]);
        }
        // ---- separator: x_S = LinvS^T LinvS (g_S - sum F_I y_I) ----
        {
            const unsigned f = fill + nb, st = f & 1u, k = f >> 1;
            const double gs = g[NA + lane] - gacc;
            mbar_wait(&sh.full_bar[st], k & 1u);
            const double *LS = stage0 + st * BLK_TILES;
            tb[lane] = gs;
            __syncwarp();
            const double ys = sm_linv_mv(LS, tb, lane);
            __syncwarp();
            yb[lane] = ys;
            __syncwarp();
            const double xv = sm_linv_mtv(LS, yb, lane);
            xs[lane] = xv;
            x[NA + lane] = xv;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty_bar[st]);
        }
        // ---- backward: x_I = Linv_I^T (y_I - F_I^T x_S - T_{I+1}^T x_{I+1}) ----
        double u = 0.0;                               // (T_{I+1}^T x_{I+1})[lane]
        double ynext = ypad[32 * (nb - 1) + lane];
        for (int i = 0; i < nb; ++i) {
            const int I = nb - 1 - i;
            const unsigned f = fill + nb + 1 + i, st = f & 1u, k = f >> 1;
            const double yv = ynext;
            if (I > 0) ynext = ypad[32 * (I - 1) + lane];
            mbar_wait(&sh.full_bar[st], k & 1u);
            const double *LT = stage0 + st * BLK_TILES, *Ft = LT + LT_TILE;
            const double v = yv - sm_f_mtv(Ft, xs, lane) - u;
            __syncwarp();
            tb[lane] = v;
            __syncwarp();
            const double xv = sm_linv_mtv(LT, tb, lane);
            yb[lane] = xv;
            if (32 * I + lane < NA) x[32 * I + lane] = xv;
            __syncwarp();
            u = (I > 0) ? sm_t_mtv(LT, yb, lane) : 0.0;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty_bar[st]);
        }
    }
    __syncthreads();
    return fill + nfill;
}

// banded cyclic mat-vec out = H v (real-indexed)
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

__global__ void __launch_bounds__(PD_THREADS, 3)
mincurv_pdip_kernel(int B, int n_max, const int32_t *__restrict__ n_pts, double *__restrict__ ws, Layout L,
                    PdipParams prm, double *__restrict__ alpha_out, int32_t *__restrict__ status,
                    int32_t *__restrict__ iters_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PdShared &sh = *reinterpret_cast<PdShared *>(smem_raw);
    if (threadIdx.x == 0) {
        mbar_init(&sh.full_bar[0], 1); mbar_init(&sh.full_bar[1], 1);
        mbar_init(&sh.empty_bar[0], 1); mbar_init(&sh.empty_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned fill = 0;      // ring fills so far (uniform across the CTA)

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
        double *YP = vec(slab, L, V_T4), *G0 = vec(slab, L, V_T0);
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
            fill = solve(sh, tiles, RHS, DX, YP, n, nb, fill);
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
            fill = solve(sh, tiles, RHS, DX, YP, n, nb, fill);
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
