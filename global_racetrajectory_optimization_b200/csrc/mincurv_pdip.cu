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
#include "mincurv_ops.cuh"

namespace mc {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PD_THREADS = 96;
constexpr int TP = 36;             // shared tile pitch (doubles): conflict-free DMMA fragment loads
constexpr int LTP = 33;            // pitch of the packed (Linv | T) tile in HBM
constexpr int LT_TILE = 32 * LTP;  // packed (Linv lower | T upper, shifted one column right)
constexpr int F_TILE = 32 * LTP;   // F_I, row-major
constexpr int BLK_TILES = LT_TILE + F_TILE;   // doubles per chain block in the slab (one contiguous bulk copy)
constexpr unsigned LT_BYTES = LT_TILE * sizeof(double);
constexpr int RING = 6;            // half-block slots (one packed LT tile or one F tile each) of the sweep ring.  Even, so that
                                   // LT tiles only ever use even slots and F tiles odd ones: every slot has ONE consumer warp,
                                   // which has consumed use k-1 before it waits for use k (a parity wait cannot tell use k from
                                   // use k-2; with 5 slots the consumers alternated and warp 0 could pass a wait on a slot whose
                                   // previous fill had not landed yet)
constexpr int RING_PAD = RING * LT_TILE - (32 * HB_PITCH + 4 * 32 * TP);      // doubles the ring needs beyond band + tiles
static_assert(RING_PAD > 0, "ring padding");

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
// same, for waits that are not on the chain's critical path (TMA producer, warp 2): back off between probes --
// a bare try_wait loop re-issues every ~8 cycles (ncu: 14 % of all executed instructions were these probes,
// each one a shared-memory transaction competing with warp 0)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, unsigned parity) {
    unsigned ok = 0;
    for (;;) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(32);
    }
}
// The factor tiles are a stream (written once, read four times per iteration, ~0.5 MB per instance, far beyond what L2
// can keep for 592 resident instances): copies and stores of tiles and band rows carry an evict-first L2 policy so that
// they do not push the O(N) iterate vectors of the interior-point loop out of L2.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
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

// Phase cycle counters of CTA 0 / warp 0 (debug aid, read with mc_debug_read_profile): cheap enough to stay compiled in.
__device__ unsigned long long g_prof[24];
#define PROF_T0() const long long _pt0 = (blockIdx.x == 0 && (threadIdx.x & 31) == 0) ? clock64() : 0
#define PROF_ADD(slot, t0) do { if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_prof[slot], (unsigned long long)(clock64() - (t0))); } while (0)

struct PdShared {
    double band[32 * HB_PITCH];      // band rows of the current chain block (next block: prefetched in registers)
    // The four factorisation tiles; during the triangular sweeps band + tiles + ringpad (49.5 KB, contiguous) hold the
    // six 8.25 KB slots of the TMA tile ring (slot s at &band[0] + s * LT_TILE).
    double As[32 * TP];              // A'_I -> L_II, column-major: As[c * TP + r]
    double Li[32 * TP];              // Linv_I, row-major
    double Ts[32 * TP];              // T_I, row-major (upper triangular)
    double Fa[32 * TP];              // F_{I-1} -> FW_I -> F_I, row-major (updated in place)
    double ringpad[RING_PAD];        // tail of the sweep ring
    uint64_t full_bar[RING], empty_bar[RING], aux_bar[4];
    double xch[4][32];               // sweeps: y_I (forward) / F_I^T x_S (backward) handed between warps 0 and 2, per visit & 3
    double vbuf[8][32];
    double red[32];
    int flag;
    int next;          // next instance index (dynamic work distribution)
};

// 1/sqrt(d) for a positive, normal d: hardware seed (rsqrt.approx.f64, ~2^-23) + one cubically convergent
// correction y += y e (1/2 + 3/8 e), e = 1 - d y^2  (relative error ~2^-68 before rounding).  Six dependent
// instructions instead of the library routine with its special-case slow path: the pivot chain of the
// Cholesky is latency bound on exactly this.
__device__ __forceinline__ double fast_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ------------------------------------------------------------------------------------------------
// Cholesky + explicit inverse of a 32x32 SPD block on 8x8 sub-blocks (one warp):
//   A (lower triangle) column-major in As (As[c*TP + r]) -> L in As (lower), Linv = L^{-1} row-major in Li
//   (full tile: zeros above the diagonal).
// Panel J (columns 8J..8J+7):
//   (1) A[:, J] -= sum_{K<J} L[:, K] L[J, K]^T               -- FP64 tensor cores (DMMA), 20 per block in total
//   (2) the updated 8x8 diagonal block is broadcast to every lane (shuffles) and factored redundantly in
//       registers, right-looking: the dependent chain per pivot is rsqrt + one multiply + one FMA; every lane
//       applies the same eliminations to its own row (the panel solve comes for free);
//   (3) the inverse of the diagonal block (8x8 triangular, registers, redundant) -> Li diagonal block.
// Off-diagonal blocks of Linv, by block distance d = 1..3 (DMMA):  W = sum_{K=J}^{I-1} L_IK Linv_KJ,
//   Linv_IJ = -Linv_II W   (W passes through its destination slot in Li to change fragment layout).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool chol_inv32(double *As, double *Li, int lane) {
    const int g = lane >> 2, q = lane & 3;
    bool ok = true;
#pragma unroll 1
    for (int J = 0; J < 4; ++J) {
        const int c0 = 8 * J;
        const long long tp0 = clock64();
        // ---- (1) panel update on the tensor cores (row blocks I >= J are independent: issue them interleaved) ----
        if (J > 0) {
            double c2[4][2];
#pragma unroll
            for (int I = 0; I < 4; ++I) {
                c2[I][0] = As[(c0 + 2 * q) * TP + 8 * I + g];
                c2[I][1] = As[(c0 + 2 * q + 1) * TP + 8 * I + g];
            }
#pragma unroll 1
            for (int ks = 0; ks < 2 * J; ++ks) {          // k = 4 ks + q runs over the finished columns 0 .. 8J-1
                const double b = As[(4 * ks + q) * TP + c0 + g];
#pragma unroll
                for (int I = 1; I < 4; ++I)               // (row block 0 is never below a panel J >= 1)
                    if (I >= J) dmma(c2[I], -As[(4 * ks + q) * TP + 8 * I + g], b);
            }
#pragma unroll
            for (int I = 1; I < 4; ++I)
                if (I >= J) {
                    As[(c0 + 2 * q) * TP + 8 * I + g] = c2[I][0];
                    As[(c0 + 2 * q + 1) * TP + 8 * I + g] = c2[I][1];
                }
            __syncwarp();
        }
        const long long tp1 = clock64();
        // ---- (2) panel factorisation: lane = row ----
        double a[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) a[t] = As[(c0 + t) * TP + lane];
        double dg[8][8];
#pragma unroll
        for (int v = 0; v < 8; ++v)
#pragma unroll
            for (int w = 0; w <= v; ++w) dg[v][w] = __shfl_sync(FULL, a[w], c0 + v);
        double rs[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (!(dg[t][t] > 0.0)) ok = false;
            rs[t] = fast_rsqrt(dg[t][t]);
            a[t] *= rs[t];
#pragma unroll
            for (int v = t + 1; v < 8; ++v) dg[v][t] *= rs[t];
#pragma unroll
            for (int v = t + 1; v < 8; ++v) {
                a[v] = fma(-a[t], dg[v][t], a[v]);
#pragma unroll
                for (int w = t + 1; w <= v; ++w) dg[v][w] = fma(-dg[v][t], dg[w][t], dg[v][w]);
            }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t)
            if (lane >= c0 + t) As[(c0 + t) * TP + lane] = a[t];
        const long long tp2 = clock64();
        // ---- (3) inverse of the diagonal block: X[v][t], v >= t  (dg[v][t] = L[v][t] for v > t, rs[t] = 1/L[t][t]) ----
        // computed redundantly in every lane; lane (g, q) then stores row g, columns 2q, 2q+1 of the 8x8 block
        // (selected with branch-free selects: a lane-indexed switch would diverge 32 ways)
        double e0 = 0.0, e1 = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            double x[8];
            x[t] = rs[t];
#pragma unroll
            for (int v = t + 1; v < 8; ++v) {
                double acc = 0.0;
#pragma unroll
                for (int u = t; u < v; ++u) acc = fma(dg[v][u], x[u], acc);
                x[v] = -acc * rs[v];
            }
#pragma unroll
            for (int v = t; v < 8; ++v) {
                const bool mine = (g == v) && (q == (t >> 1));
                if (t & 1) e1 = mine ? x[v] : e1;
                else e0 = mine ? x[v] : e0;
            }
        }
        *reinterpret_cast<double2 *>(&Li[(c0 + g) * TP + c0 + 2 * q]) = make_double2(e0, e1);
        // zero the blocks of Li to the right of the diagonal block (rows c0..c0+7, columns c0+8..31)
        for (int e = lane; e < 8 * (24 - c0); e += 32) {
            const int rr = e / (24 - c0), cc = e - rr * (24 - c0);
            Li[(c0 + rr) * TP + c0 + 8 + cc] = 0.0;
        }
        __syncwarp();
        if (blockIdx.x == 0 && lane == 0) {
            const long long tp3 = clock64();
            atomicAdd(&g_prof[8], (unsigned long long)(tp3 - tp2));
        }
    }
    const long long tq0 = clock64();
    // ---- off-diagonal blocks of the inverse, by block distance (blocks of one distance are independent) ----
#pragma unroll
    for (int d = 1; d < 4; ++d) {
        double w2[3][2];
#pragma unroll
        for (int J = 0; J + d < 4; ++J) {
            const int I = J + d;
            w2[J][0] = 0.0; w2[J][1] = 0.0;
#pragma unroll
            for (int ks = 2 * J; ks < 2 * I; ++ks)       // k = 4 ks + q over blocks K = J .. I-1
                dmma(w2[J], As[(4 * ks + q) * TP + 8 * I + g], Li[(4 * ks + q) * TP + 8 * J + g]);
        }
#pragma unroll
        for (int J = 0; J + d < 4; ++J)
            *reinterpret_cast<double2 *>(&Li[(8 * (J + d) + g) * TP + 8 * J + 2 * q]) = make_double2(w2[J][0], w2[J][1]);
        __syncwarp();
        double x2[3][2];
#pragma unroll
        for (int J = 0; J + d < 4; ++J) {
            const int I = J + d;
            const double b0 = Li[(8 * I + q) * TP + 8 * J + g], b1 = Li[(8 * I + 4 + q) * TP + 8 * J + g];
            x2[J][0] = 0.0; x2[J][1] = 0.0;
            dmma(x2[J], -Li[(8 * I + g) * TP + 8 * I + q], b0);
            dmma(x2[J], -Li[(8 * I + g) * TP + 8 * I + 4 + q], b1);
        }
        __syncwarp();
#pragma unroll
        for (int J = 0; J + d < 4; ++J)
            *reinterpret_cast<double2 *>(&Li[(8 * (J + d) + g) * TP + 8 * J + 2 * q]) = make_double2(x2[J][0], x2[J][1]);
        __syncwarp();
    }
    if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_prof[2], (unsigned long long)(clock64() - tq0));
    return ok;
}

// S[own 16 rows][:] -= F[own rows][:] F^T  (rolled over the four 8-wide k blocks to keep the code small).  S is
// symmetric and chol_inv32 reads its lower triangle only: the blocks right of the diagonal are skipped (10 of the
// 16 block products; their accumulators keep stale values that nobody reads).
__device__ __forceinline__ void s_update(double (&sacc)[2][4][2], const double *Ft, int ib, int g, int q) {
#pragma unroll 1
    for (int ks = 0; ks < 8; ++ks) {
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = -Ft[(8 * (ib + i) + g) * TP + 4 * ks + q];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Ft[(8 * j + g) * TP + 4 * ks + q];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j <= ib + i) dmma(sacc[i][j], a[i], b[j]);
    }
}


// CTA-wide barrier between the warp roles of the factorisation.  The roles sit in different (noinline) functions, so the
// barrier is reached at different program points: a named barrier with an explicit thread count states exactly that
// (every warp arrives convergently; __syncthreads() in role-divergent code is flagged by compute-sanitizer synccheck).
__device__ __forceinline__ void role_sync() { named_bar_sync(10, PD_THREADS); }

// ---- X Linv^T for nrb row blocks starting at rb0 (Linv lower triangular: k blocks K <= j), in place in the shared tile X
//      and to the HBM tile gx (pitch LTP).  Used for F_I = FW_I Linv_I^T.  The 20 B fragments of Linv are the same for
//      every row block: loaded once. ----
__device__ __forceinline__ void rows_times_linvT(double *X, const double *Li, double *__restrict__ gx, int rb0, int nrb, int g, int q) {
    double b[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int ks = 0; ks < 2 * (j + 1); ++ks) b[j][ks] = Li[(8 * j + g) * TP + 4 * ks + q];
#pragma unroll 1
    for (int rb = rb0; rb < rb0 + nrb; ++rb) {
        double a[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) a[ks] = X[(8 * rb + g) * TP + 4 * ks + q];
        __syncwarp();
        double c2[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) { c2[j][0] = 0.0; c2[j][1] = 0.0; }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ks < 2 * (j + 1)) dmma(c2[j], a[ks], b[j][ks]);      // the four column blocks are independent chains
        const int r = 8 * rb + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 8 * j + 2 * q;
            *reinterpret_cast<double2 *>(&X[r * TP + c]) = make_double2(c2[j][0], c2[j][1]);
            __stcs(&gx[r * LTP + c], c2[j][0]);
            __stcs(&gx[r * LTP + c + 1], c2[j][1]);
        }
    }
}

// ---- warp 0: the chain.  Rounds I = 0..nb-1 are the chain blocks, round nb the separator block.
//      Phase A: Cholesky + inverse of the block the fill warps assembled into As.
//      Phase B: Linv_I and F_I = FW_I Linv_I^T to the HBM tiles (the fill warps build T_{I+1} and A'_{I+1}). ----
__device__ __noinline__ bool factor_chain(PdShared &sh, double *__restrict__ tiles, int n, int nb) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    bool ok = true;
    role_sync();                      // the fill warps have assembled A'_0 into As
    for (int I = 0; I <= nb; ++I) {
        // =========================== phase A ===========================
        if (I == nb) named_bar_sync(3, PD_THREADS);      // the fill warps have put the separator Schur complement into As
        const long long tB = clock64();
        ok = chol_inv32(sh.As, sh.Li, lane) && ok;
        const long long tC = clock64();
        if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_prof[1], (unsigned long long)(tC - tB));
        role_sync();
        const long long tE = clock64();
        PROF_ADD(3, tC);
        // =========================== phase B ===========================
        {
            // Linv_I -> packed HBM tile (lower part, row-major pitch LTP); round nb: the separator's slot
            double *gt = tiles + (size_t)I * BLK_TILES;
#pragma unroll 4
            for (int r = 0; r < 32; ++r)
                if (lane <= r) __stcs(&gt[r * LTP + lane], sh.Li[r * TP + lane]);
            if (I < nb) {
                // ---- F_I = FW_I Linv_I^T, in place in the shared F tile and to HBM ----
                double *gf = gt + LT_TILE;
                rows_times_linvT(sh.Fa, sh.Li, gf, 0, 4, g, q);
            }
        }
        const long long tS2 = clock64();
        PROF_ADD(4, tE);
        role_sync();
        PROF_ADD(5, tS2);
    }
    return ok;
}

// ---- fill warps, phase B: A'_J = A_J + D_J - T_J T_J^T (lower 8x8 blocks of row block i) from the staged band rows
//      of block J (diagonal already carries D) and the T tile, column-major into As ----
template <int i>
__device__ __forceinline__ void assemble_rowblock(PdShared &sh, int baseJ, int NA, bool hasT, int g, int q) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
        double c2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int r = 8 * i + g, c = 8 * j + 2 * q + e;
            const int lo = (c < r) ? c : r, dist = (c < r) ? r - c : c - r;
            double v = sh.band[lo * HB_PITCH + dist];
            const bool real = (baseJ + r) < NA && (baseJ + c) < NA;
            if (!real) v = (r == c) ? 1.0 : 0.0;
            c2[e] = v;
        }
        if (hasT) {
#pragma unroll
            for (int K = i; K < 4; ++K)      // T upper triangular: blocks (i,K), (j,K) nonzero for K >= i >= j
#pragma unroll
                for (int s = 0; s < 2; ++s)
                    dmma(c2, -sh.Ts[(8 * i + g) * TP + 8 * K + 4 * s + q], sh.Ts[(8 * j + g) * TP + 8 * K + 4 * s + q]);
        }
        sh.As[(8 * j + 2 * q) * TP + 8 * i + g] = c2[0];
        sh.As[(8 * j + 2 * q + 1) * TP + 8 * i + g] = c2[1];
    }
}

// ---- fill warps, phase B: row block i of T_{I+1} = B_I Linv_I^T (upper-triangular blocks j >= i, K = i..j) into Ts and
//      the HBM tile gn of block I+1 (shifted one column right, next to that block's Linv) ----
template <int i>
__device__ __forceinline__ void coupling_rowblock(PdShared &sh, double *__restrict__ gn, int base, int NA, int g, int q) {
#pragma unroll
    for (int j = i; j < 4; ++j) {
        double c2[2] = {0.0, 0.0};
#pragma unroll
        for (int K = i; K <= j; ++K)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                // B_I[r][k] = M[base+32+r][base+k] = band[k][32 + r - k] for r <= k
                const int r = 8 * i + g, k = 8 * K + 4 * s + q;
                double av = 0.0;
                if (r <= k && (base + 32 + r) < NA) av = sh.band[k * HB_PITCH + 32 + r - k];
                dmma(c2, av, sh.Li[(8 * j + g) * TP + k]);
            }
        const int r = 8 * i + g, c = 8 * j + 2 * q;
        *reinterpret_cast<double2 *>(&sh.Ts[r * TP + c]) = make_double2(c2[0], c2[1]);
        if (c >= r) __stcs(&gn[r * LTP + c + 1], c2[0]);
        if (c + 1 >= r) __stcs(&gn[r * LTP + c + 2], c2[1]);
    }
}

// ---- warps 1, 2: the two 16-row halves of the separator fill row (FW, S) in phase A; in phase B the coupling block
//      T_{I+1}, the band rows of block I+1 and the next diagonal block A'_{I+1} (so warp 0 only ever factors) ----
__device__ __noinline__ void factor_fill(PdShared &sh, const double *__restrict__ HB, const double *__restrict__ DD,
                                        double *__restrict__ tiles, int n, int nb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int NA = n - 32;
    const int ib = 2 * (warp - 1);
    double sacc[2][4][2];          // own 16 rows of the separator Schur complement (C fragments)
    double breg[17];               // half of the next block's band rows (1088 doubles = 2 x 32 x 17)
    double dreg = 0.0;             // ... and of its barrier diagonal D (lanes 0..15: one row each)
    double *Ft = sh.Fa;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) sacc[i][j][e] = sh.As[(8 * j + 2 * q + e) * TP + 8 * (ib + i) + g];
    named_bar_sync(1, 64);         // both fill warps have picked the separator block up: As is free
    // A'_0 (no coupling term) from the band rows of block 0 staged by factor(); prefetch the rows of block 1
    if (warp == 1) { assemble_rowblock<0>(sh, 0, NA, false, g, q); assemble_rowblock<1>(sh, 0, NA, false, g, q); }
    else           { assemble_rowblock<2>(sh, 0, NA, false, g, q); assemble_rowblock<3>(sh, 0, NA, false, g, q); }
    if (1 < nb) {
        const double *src = HB + (size_t)32 * HB_PITCH + (warp - 1) * 544;
#pragma unroll
        for (int t = 0; t < 17; ++t) breg[t] = __ldcs(&src[32 * t + lane]);
        const int nd = 32 + 16 * (warp - 1) + lane;
        dreg = (lane < 16 && nd < NA) ? DD[nd] : 0.0;
    }
    role_sync();
    for (int I = 0; I <= nb; ++I) {
        const int base = 32 * I;
        // =========================== phase A ===========================
        if (I == nb) {
            // last round: the pending S -= F_{nb-1} F_{nb-1}^T, then the separator is handed to warp 0 through As
            s_update(sacc, Ft, ib, g, q);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    sh.As[(8 * j + 2 * q) * TP + 8 * (ib + i) + g] = sacc[i][j][0];
                    sh.As[(8 * j + 2 * q + 1) * TP + 8 * (ib + i) + g] = sacc[i][j][1];
                }
            __syncwarp();
            named_bar_arrive(3, PD_THREADS);
        } else {
            // ---- S -= F_{I-1} F_{I-1}^T (own 16 rows, reads the whole F tile) ----
            if (I > 0) s_update(sacc, Ft, ib, g, q);
            // ---- FW = Y_I - F_{I-1} T_I^T, in place in the F tile (own rows) ----
            double a[2][8];
            if (I > 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) a[i][ks] = Ft[(8 * (ib + i) + g) * TP + 4 * ks + q];
            }
            named_bar_sync(1, 64);       // both fill warps are done reading F_{I-1}
            const bool hasY = (I == 0) || (base + 31 >= NA - 32);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    double c2[2] = {0.0, 0.0};
                    if (hasY) {      // Y_I = M[S, block I]: nonzero only across the wrap (I = 0) and next to the separator
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int r = 8 * (ib + i) + g, c = 8 * j + 2 * q + e;
                            if ((base + c) < NA) c2[e] = hentry(HB, n, NA + r, base + c);
                        }
                    }
                    if (I > 0) {
                        double m2[2] = {0.0, 0.0};
                        for (int K = j; K < 4; ++K) {      // T[c][k] != 0 for k >= c
                            const double ak0 = (K == 0) ? a[i][0] : (K == 1) ? a[i][2] : (K == 2) ? a[i][4] : a[i][6];
                            const double ak1 = (K == 0) ? a[i][1] : (K == 1) ? a[i][3] : (K == 2) ? a[i][5] : a[i][7];
                            dmma(m2, ak0, sh.Ts[(8 * j + g) * TP + 8 * K + q]);
                            dmma(m2, ak1, sh.Ts[(8 * j + g) * TP + 8 * K + 4 + q]);
                        }
                        c2[0] -= m2[0];
                        c2[1] -= m2[1];
                    }
                    *reinterpret_cast<double2 *>(&Ft[(8 * (ib + i) + g) * TP + 8 * j + 2 * q]) = make_double2(c2[0], c2[1]);
                }
            }
        }
        role_sync();
        // =========================== phase B ===========================
        if (I + 1 < nb) {
            // ---- T_{I+1} = B_I Linv_I^T (band rows of block I, coupling part) ----
            double *gn = tiles + (size_t)(I + 1) * BLK_TILES;
            const long long tw1 = clock64();
            if (warp == 1) coupling_rowblock<0>(sh, gn, base, NA, g, q);
            else { coupling_rowblock<1>(sh, gn, base, NA, g, q); coupling_rowblock<2>(sh, gn, base, NA, g, q); coupling_rowblock<3>(sh, gn, base, NA, g, q); }
            named_bar_sync(1, 64);       // T complete; nobody reads the band rows of block I any more
            const long long tw2 = clock64();
            if (blockIdx.x == 0 && threadIdx.x == 32) atomicAdd(&g_prof[13], (unsigned long long)(tw2 - tw1));
#pragma unroll
            for (int t = 0; t < 17; ++t) sh.band[(warp - 1) * 544 + 32 * t + lane] = breg[t];
            __syncwarp();
            if (lane < 16) sh.band[(warp - 1) * 544 + lane * HB_PITCH] += dreg;      // + barrier diagonal D
            named_bar_sync(1, 64);       // band rows of block I+1 staged
            if (I + 2 < nb) {
                const double *src = HB + (size_t)(base + 64) * HB_PITCH + (warp - 1) * 544;
#pragma unroll
                for (int t = 0; t < 17; ++t) breg[t] = __ldcs(&src[32 * t + lane]);
                const int nd = base + 64 + 16 * (warp - 1) + lane;
                dreg = (lane < 16 && nd < NA) ? DD[nd] : 0.0;
            }
            // ---- A'_{I+1} = A + D - T T^T ----
            if (warp == 1) { assemble_rowblock<0>(sh, base + 32, NA, true, g, q); assemble_rowblock<1>(sh, base + 32, NA, true, g, q); }
            else           { assemble_rowblock<2>(sh, base + 32, NA, true, g, q); assemble_rowblock<3>(sh, base + 32, NA, true, g, q); }
            if (blockIdx.x == 0 && threadIdx.x == 32) atomicAdd(&g_prof[14], (unsigned long long)(clock64() - tw2));
        }
        role_sync();
    }
}

__device__ __noinline__ bool factor(PdShared &sh, const double *__restrict__ HB, const double *__restrict__ DD,
                                    double *__restrict__ tiles, int n, int nb) {
    const int warp = threadIdx.x >> 5;
    const int NA = n - 32;
    // separator diagonal block C + D_S: gathered into As by all threads, then picked up as C fragments by the fill warps
#pragma unroll 1
    for (int e = threadIdx.x; e < 1024; e += PD_THREADS) {
        const int r = e >> 5, c = e & 31;
        double v = hentry(HB, n, NA + r, NA + c);
        if (r == c) v += DD[NA + r];
        sh.As[c * TP + r] = v;
    }
    // band rows of block 0
    for (int e = threadIdx.x; e < 32 * HB_PITCH; e += PD_THREADS) {
        const int r = e / HB_PITCH;
        double v = HB[e];
        if (e == r * HB_PITCH && r < NA) v += DD[r];           // diagonal: H_ii + D_i
        sh.band[e] = v;
    }
    __syncthreads();
    if (warp == 0) {
        if (!factor_chain(sh, tiles, n, nb)) sh.flag = 1;
    } else {
        factor_fill(sh, HB, DD, tiles, n, nb);
    }
    __syncthreads();
    return sh.flag == 0;
}

// ---- mat-vecs over one staged block in shared memory: LT = packed (Linv | T), pitch 33; F pitch 33 ----
// (odd pitch: both "lane = row" and "lane = column" access patterns are bank-conflict free; v is broadcast)
__device__ __forceinline__ double sm_t_mv(const double *LT, const double *v, int lane) {     // sum_{c >= r} T[r][c] v[c]
    const double *row = LT + lane * LTP + 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
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
#pragma unroll 4
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
#pragma unroll 4
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
#pragma unroll 4
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
#pragma unroll 4
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
#pragma unroll 4
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
// The factor blocks are visited in the order [0 .. nb-1, S, nb-1 .. 0]; every visit streams two tiles HBM -> shared
// with cp.async.bulk (TMA): the packed (Linv | T) tile for warp 0, which runs the chain recurrences (two triangular
// 32x32 mat-vecs per block), and the F tile for warp 2, which runs the separator-row products F_I y_I / F_I^T x_S
// next to it.  The tiles go through a ring of RING = 6 slots (three blocks of prefetch: the sweeps are bound by the
// HBM fetch latency of a tile, not by the mat-vecs) guarded by per-slot full/empty mbarriers; an elected lane of warp 1
// is the producer.  `fill` counts the visits of this CTA so far (slot and parity bookkeeping): visit v uses fill
// numbers 2v (LT) and 2v+1 (F), slot = fill number % RING.  The new count is returned.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ unsigned solve(PdShared &sh, const double *__restrict__ tiles, const double *__restrict__ g,
                                       double *__restrict__ x, double *__restrict__ ypad, int n, int nb, unsigned fill) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NA = n - 32;
    double *ring = sh.band;
    const unsigned nvis = 2u * nb + 1u;
    double *tb = sh.vbuf[1], *xb = sh.vbuf[0], *xs = sh.vbuf[3], *gsum = sh.vbuf[4];
    fence_proxy_async();        // the ring area was last written through the generic proxy (band rows, factor tiles)
    __syncthreads();
    if (warp == 1) {
        // ---- producer: one elected thread issues the bulk copies ----
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            for (unsigned i = 0; i < nvis; ++i) {
                const unsigned blk = (i < (unsigned)nb) ? i : ((i == (unsigned)nb) ? (unsigned)nb : 2u * nb - i);
                for (unsigned h = 0; h < 2; ++h) {      // (the separator has no F tile: its slot is copied but never read)
                    const unsigned f = 2u * (fill + i) + h, sl = f % RING, k = f / RING;
                    if (k > 0) mbar_wait_relaxed(&sh.empty_bar[sl], (k - 1) & 1u);
                    mbar_expect_tx(&sh.full_bar[sl], LT_BYTES);
                    tma_load_1d(ring + sl * LT_TILE, tiles + (size_t)blk * BLK_TILES + h * LT_TILE, LT_BYTES, &sh.full_bar[sl], pol);
                }
            }
        }
    } else if (warp == 2) {
        // ---- separator-row mat-vecs, off the chain: forward gS -= F_I y_I (behind warp 0), backward c_I = F_I^T x_S
        //      (ahead of warp 0).  aux_bar[v & 3] / xch[v & 3] hand y_I / c_I over per visit v. ----
        double gacc = 0.0;
        for (unsigned i = 0; i < nvis; ++i) {
            const unsigned v = fill + i, f = 2u * v + 1u, sl = f % RING, k = f / RING, a = v & 3u;
            mbar_wait_relaxed(&sh.full_bar[sl], k & 1u);
            const double *Ft = ring + sl * LT_TILE;
            if (i < (unsigned)nb) {
                mbar_wait_relaxed(&sh.aux_bar[a], (v >> 2) & 1u);
                gacc += sm_f_mv(Ft, sh.xch[a], lane);
                if (i + 1 == (unsigned)nb) {
                    gsum[lane] = gacc;
                    __syncwarp();
                    named_bar_arrive(9, 64);
                }
            } else if (i == (unsigned)nb) {
                mbar_wait_relaxed(&sh.aux_bar[a], (v >> 2) & 1u);          // x_S is in xs
            } else {
                sh.xch[a][lane] = sm_f_mtv(Ft, xs, lane);
            }
            __syncwarp();
            if (lane == 0) {
                if (i > (unsigned)nb) mbar_arrive(&sh.aux_bar[a]);
                mbar_arrive(&sh.empty_bar[sl]);
            }
        }
    } else {
        // ---- forward: y_I = Linv_I (g_I - T_I y_{I-1}) ----
        const long long tfw = clock64();
        double gnext = (lane < NA) ? g[lane] : 0.0;
        for (int I = 0; I < nb; ++I) {
            const unsigned v = fill + I, f = 2u * v, sl = f % RING, k = f / RING, a = v & 3u;
            const double gv = gnext;
            if (I + 1 < nb) { const int nd = 32 * (I + 1) + lane; gnext = (nd < NA) ? g[nd] : 0.0; }
            const long long tw = clock64();
            mbar_wait(&sh.full_bar[sl], k & 1u);
            PROF_ADD(7, tw);
            const double *LT = ring + sl * LT_TILE;
            double vv = gv;
            if (I > 0) vv -= sm_t_mv(LT, sh.xch[(v - 1u) & 3u], lane);
            tb[lane] = vv;
            __syncwarp();
            const double y = sm_linv_mv(LT, tb, lane);
            sh.xch[a][lane] = y;
            ypad[32 * I + lane] = y;
            __syncwarp();
            if (lane == 0) { mbar_arrive(&sh.aux_bar[a]); mbar_arrive(&sh.empty_bar[sl]); }
        }
        PROF_ADD(21, tfw);
        const long long tsep = clock64();
        // ---- separator: x_S = LinvS^T LinvS (g_S - sum F_I y_I) ----
        {
            const unsigned v = fill + nb, f = 2u * v, sl = f % RING, k = f / RING, a = v & 3u;
            const double gs0 = g[NA + lane];
            mbar_wait(&sh.full_bar[sl], k & 1u);
            named_bar_sync(9, 64);
            const double *LS = ring + sl * LT_TILE;
            tb[lane] = gs0 - gsum[lane];
            __syncwarp();
            const double ys = sm_linv_mv(LS, tb, lane);
            xb[lane] = ys;
            __syncwarp();
            const double xv = sm_linv_mtv(LS, xb, lane);
            xs[lane] = xv;
            x[NA + lane] = xv;
            __syncwarp();
            if (lane == 0) { mbar_arrive(&sh.aux_bar[a]); mbar_arrive(&sh.empty_bar[sl]); }
        }
        PROF_ADD(22, tsep);
        const long long tbw = clock64();
        // ---- backward: x_I = Linv_I^T (y_I - F_I^T x_S - T_{I+1}^T x_{I+1}) ----
        double u = 0.0;                               // (T_{I+1}^T x_{I+1})[lane]
        double ynext = ypad[32 * (nb - 1) + lane];
        for (int i = 0; i < nb; ++i) {
            const int I = nb - 1 - i;
            const unsigned v = fill + nb + 1 + i, f = 2u * v, sl = f % RING, k = f / RING, a = v & 3u;
            const double yv = ynext;
            if (I > 0) ynext = ypad[32 * (I - 1) + lane];
            const long long tw = clock64();
            mbar_wait(&sh.full_bar[sl], k & 1u);
            mbar_wait(&sh.aux_bar[a], (v >> 2) & 1u);
            PROF_ADD(19, tw);
            const double *LT = ring + sl * LT_TILE;
            tb[lane] = yv - sh.xch[a][lane] - u;
            __syncwarp();
            const double xv = sm_linv_mtv(LT, tb, lane);
            xb[lane] = xv;
            if (32 * I + lane < NA) x[32 * I + lane] = xv;
            __syncwarp();
            u = (I > 0) ? sm_t_mtv(LT, xb, lane) : 0.0;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty_bar[sl]);
        }
        PROF_ADD(20, tbw);
    }
    const long long tend = clock64();
    __syncthreads();
    PROF_ADD(23, tend);
    return fill + nvis;
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

__global__ void __launch_bounds__(PD_THREADS, 4)
mincurv_pdip_kernel(int B, int n_max, const int32_t *__restrict__ n_pts, double *__restrict__ ws, Layout L,
                    PdipParams prm, double *__restrict__ alpha_out, int32_t *__restrict__ status,
                    int32_t *__restrict__ iters_out, int *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PdShared &sh = *reinterpret_cast<PdShared *>(smem_raw);
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) { mbar_init(&sh.full_bar[i], 1); mbar_init(&sh.empty_bar[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&sh.aux_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned fill = 0;      // ring fills so far (uniform across the CTA)

    for (;;) {
        // instances are handed out dynamically (iteration counts differ between instances)
        __syncthreads();
        if (threadIdx.x == 0) sh.next = atomicAdd(work_counter, 1);
        __syncthreads();
        const int b = sh.next;
        if (b >= B) break;
        const int n = n_pts ? n_pts[b] : n_max;
        double *aout = alpha_out + (size_t)b * n_max;
        if (status[b] != 0) {
            for (int i = threadIdx.x; i < n_max; i += PD_THREADS) aout[i] = 0.0;
            if (iters_out && threadIdx.x == 0) iters_out[b] = 0;
            continue;
        }
        const long long tq0 = clock64();
        double *slab = ws + (size_t)b * L.stride;
        const double *HB = slab + L.o_hb;
        double *tiles = slab + L.o_tiles;
        const double *__restrict__ LB = vec(slab, L, V_LB), *__restrict__ UB = vec(slab, L, V_UB), *__restrict__ F = vec(slab, L, V_F);
        double *__restrict__ AL = vec(slab, L, V_ALPHA), *__restrict__ LU = vec(slab, L, V_LU), *__restrict__ LL = vec(slab, L, V_LL), *__restrict__ RD = vec(slab, L, V_RD);
        double *__restrict__ RHS = vec(slab, L, V_RHS), *__restrict__ DX = vec(slab, L, V_DX), *__restrict__ DD = vec(slab, L, V_DD);
        double *__restrict__ TU = vec(slab, L, V_DLU), *__restrict__ TL = vec(slab, L, V_DLL), *__restrict__ SU = vec(slab, L, V_SU), *__restrict__ SL = vec(slab, L, V_SL);
        double *__restrict__ ISU = vec(slab, L, V_ISU), *__restrict__ ISL = vec(slab, L, V_ISL);
        double *YP = vec(slab, L, V_YPAD), *G0 = vec(slab, L, V_T0);
        const int nb = (n - 32 + 31) / 32;
        if (threadIdx.x == 0) sh.flag = 0;

        // ---------------- initial point: box centre, multipliers from the gradient ----------------
        for (int i = threadIdx.x; i < n; i += PD_THREADS) AL[i] = 0.5 * (LB[i] + UB[i]);
        __syncthreads();
        band_matvec(HB, AL, G0, n);
        __syncthreads();
        double gmax = 0.0, fmaxv = 0.0;
#pragma unroll 1
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
#pragma unroll 1
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
            ISU[i] = 1.0 / su; ISL[i] = 1.0 / sl;
            musum += su * lu + sl * ll;
        }
        musum = block_reduce<0>(musum, sh.red);
        const double mu0 = musum / (2.0 * n);
        const double rd_tol = prm.rd_rel * (fmaxv + gmax) + 1e-300;
        double mu = mu0;
        int it = 0;
        int result = 2;   // max-iter unless we converge

        // Vector phases: the reciprocals 1/s_u, 1/s_l are state (ISU, ISL), so one interior-point iteration
        // costs four divisions per variable; step lengths come from max-ratios (no division per element).
        // Every thread owns the elements i = tid + k * 96; the loops take them in groups of VG with all loads of a
        // group issued before the first use (the vectors live in L2/HBM: one element at a time, each of the ~10
        // trips of a loop paid the full memory latency, ~9 % of the kernel).
        constexpr int VG = 4, VS = VG * PD_THREADS;
        for (it = 0; it < prm.max_iter; ++it) {
            // ---- barrier diagonal and affine right-hand side ----
            const long long tv1 = clock64();
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS) {
                double lu[VG], ll[VG], isu[VG], isl[VG], rd[VG];
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = min(i0 + k * PD_THREADS, n - 1);
                    lu[k] = LU[i]; ll[k] = LL[i]; isu[k] = ISU[i]; isl[k] = ISL[i]; rd[k] = RD[i];
                }
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = i0 + k * PD_THREADS;
                    if (i < n) { DD[i] = lu[k] * isu[k] + ll[k] * isl[k]; RHS[i] = -rd[k] + lu[k] - ll[k]; }
                }
            }
            __syncthreads();
            PROF_ADD(16, tv1);
            const long long tf0 = clock64();
            if (!factor(sh, HB, DD, tiles, n, nb)) { result = 3; break; }
            PROF_ADD(10, tf0);
            const long long ts0 = clock64();
            fill = solve(sh, tiles, RHS, DX, YP, n, nb, fill);
            PROF_ADD(6, ts0);
            // ---- affine direction: step lengths 1 / max-ratio; mu_aff as a polynomial in (ap, ad) ----
            const long long tv2 = clock64();
            double rp = 0.0, rdl = 0.0, c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS) {
                double su[VG], sl[VG], lu[VG], ll[VG], dx[VG], isu[VG], isl[VG];
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = min(i0 + k * PD_THREADS, n - 1);
                    su[k] = SU[i]; sl[k] = SL[i]; lu[k] = LU[i]; ll[k] = LL[i]; dx[k] = DX[i]; isu[k] = ISU[i]; isl[k] = ISL[i];
                }
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    if (i0 + k * PD_THREADS < n) {
                        const double p = dx[k] * isu[k], m = dx[k] * isl[k];
                        rp = fmax(rp, fmax(p, -m));                 // s_u - a dx >= 0, s_l + a dx >= 0
                        rdl = fmax(rdl, fmax(1.0 - p, 1.0 + m));    // dlu / lu = -1 + p, dll / ll = -1 - m
                        const double dlu = lu[k] * (p - 1.0), dll = -ll[k] * (1.0 + m);
                        c00 += su[k] * lu[k] + sl[k] * ll[k];
                        c01 += su[k] * dlu + sl[k] * dll;           // coefficient of ad
                        c10 += dx[k] * (ll[k] - lu[k]);             // coefficient of ap
                        c11 += dx[k] * (dll - dlu);                 // coefficient of ap * ad
                    }
                }
            }
            rp = block_reduce<1>(rp, sh.red);
            rdl = block_reduce<1>(rdl, sh.red);
            double ap = (rp > 1.0) ? 1.0 / rp : 1.0, ad = (rdl > 1.0) ? 1.0 / rdl : 1.0;
            c00 = block_reduce<0>(c00, sh.red); c01 = block_reduce<0>(c01, sh.red);
            c10 = block_reduce<0>(c10, sh.red); c11 = block_reduce<0>(c11, sh.red);
            const double mua = (c00 + ad * c01 + ap * c10 + ap * ad * c11) / (2.0 * n);
            double sigma = mua / mu;
            sigma = sigma * sigma * sigma;
            const double smu = sigma * mu;
            PROF_ADD(17, tv2);
            const long long tv3 = clock64();
            // ---- corrector right-hand side ----
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS) {
                double su[VG], sl[VG], lu[VG], ll[VG], dx[VG], isu[VG], isl[VG], rd[VG];
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = min(i0 + k * PD_THREADS, n - 1);
                    su[k] = SU[i]; sl[k] = SL[i]; lu[k] = LU[i]; ll[k] = LL[i]; dx[k] = DX[i]; isu[k] = ISU[i]; isl[k] = ISL[i]; rd[k] = RD[i];
                }
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = i0 + k * PD_THREADS;
                    if (i < n) {
                        const double dlu = lu[k] * (dx[k] * isu[k] - 1.0), dll = -ll[k] * (1.0 + dx[k] * isl[k]);
                        const double tu = smu - su[k] * lu[k] + dx[k] * dlu;
                        const double tl = smu - sl[k] * ll[k] - dx[k] * dll;
                        TU[i] = tu; TL[i] = tl;
                        RHS[i] = -rd[k] - tu * isu[k] + tl * isl[k];
                    }
                }
            }
            __syncthreads();
            PROF_ADD(18, tv3);
            const long long ts1 = clock64();
            fill = solve(sh, tiles, RHS, DX, YP, n, nb, fill);
            PROF_ADD(6, ts1);
            const long long tv4 = clock64();
            rp = 0.0; rdl = 0.0;
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS) {
                double lu[VG], ll[VG], dx[VG], isu[VG], isl[VG], tu[VG], tl[VG];
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = min(i0 + k * PD_THREADS, n - 1);
                    lu[k] = LU[i]; ll[k] = LL[i]; dx[k] = DX[i]; isu[k] = ISU[i]; isl[k] = ISL[i]; tu[k] = TU[i]; tl[k] = TL[i];
                }
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    if (i0 + k * PD_THREADS < n) {
                        const double dlu = (tu[k] + lu[k] * dx[k]) * isu[k], dll = (tl[k] - ll[k] * dx[k]) * isl[k];
                        rp = fmax(rp, fmax(dx[k] * isu[k], -dx[k] * isl[k]));
                        rdl = fmax(rdl, fmax(-dlu / lu[k], -dll / ll[k]));
                    }
                }
            }
            rp = block_reduce<1>(rp, sh.red);
            rdl = block_reduce<1>(rdl, sh.red);
            ap = (prm.eta < rp) ? prm.eta / rp : 1.0;       // min(1, eta / max-ratio)
            ad = (prm.eta < rdl) ? prm.eta / rdl : 1.0;
            PROF_ADD(15, tv4);
            const long long tv5 = clock64();
            double musum2 = 0.0, rdmax = 0.0, dxmax = 0.0, amax = 0.0;
            constexpr int VG2 = 2, VS2 = VG2 * PD_THREADS;        // 13 input vectors: groups of two
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS2) {
                double su[VG2], sl[VG2], lu[VG2], ll[VG2], dx[VG2], isu[VG2], isl[VG2], tu[VG2], tl[VG2], al[VG2], rd[VG2], rh[VG2], dd[VG2];
#pragma unroll
                for (int k = 0; k < VG2; ++k) {
                    const int i = min(i0 + k * PD_THREADS, n - 1);
                    su[k] = SU[i]; sl[k] = SL[i]; lu[k] = LU[i]; ll[k] = LL[i]; dx[k] = DX[i]; isu[k] = ISU[i]; isl[k] = ISL[i];
                    tu[k] = TU[i]; tl[k] = TL[i]; al[k] = AL[i]; rd[k] = RD[i]; rh[k] = RHS[i]; dd[k] = DD[i];
                }
#pragma unroll
                for (int k = 0; k < VG2; ++k) {
                    const int i = i0 + k * PD_THREADS;
                    if (i < n) {
                        const double dlu = (tu[k] + lu[k] * dx[k]) * isu[k], dll = (tl[k] - ll[k] * dx[k]) * isl[k];
                        const double an = al[k] + ap * dx[k], lun = lu[k] + ad * dlu, lln = ll[k] + ad * dll;
                        const double sun = su[k] - ap * dx[k], sln = sl[k] + ap * dx[k];
                        // H dx = rhs - D dx  (M dx = rhs)
                        const double rdn = rd[k] + ap * (rh[k] - dd[k] * dx[k]) + ad * (dlu - dll);
                        AL[i] = an; LU[i] = lun; LL[i] = lln; RD[i] = rdn; SU[i] = sun; SL[i] = sln;
                        ISU[i] = 1.0 / sun; ISL[i] = 1.0 / sln;
                        musum2 += sun * lun + sln * lln;
                        rdmax = fmax(rdmax, fabs(rdn));
                        dxmax = fmax(dxmax, fabs(dx[k]));
                        amax = fmax(amax, fabs(an));
                    }
                }
            }
            mu = block_reduce<0>(musum2, sh.red) / (2.0 * n);
            rdmax = block_reduce<1>(rdmax, sh.red);
            PROF_ADD(0, tv5);
            // weakly active bounds converge like sqrt(mu): also require that the step itself has become small
            bool settled = true;
            if (prm.dx_rel > 0.0 && mu <= prm.mu_rel * mu0) {
                dxmax = block_reduce<1>(dxmax, sh.red);
                amax = block_reduce<1>(amax, sh.red);
                settled = ap * dxmax <= prm.dx_rel * fmax(amax, 0.01);
            }
            if (mu <= prm.mu_rel * mu0 && rdmax <= rd_tol && settled) { result = 0; ++it; break; }
            if (mu <= 1e-4 * prm.mu_rel * mu0) { result = (rdmax <= 1e3 * rd_tol) ? 0 : 2; ++it; break; }   // complementarity exhausted
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_max; i += PD_THREADS) aout[i] = (i < n) ? AL[i] : 0.0;
        if (threadIdx.x == 0) {
            status[b] = result;
            if (iters_out) iters_out[b] = it;
        }
        PROF_ADD(9, tq0);
        if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&g_prof[11], 1ull); atomicAdd(&g_prof[12], (unsigned long long)it); }
    }
}

size_t pdip_smem_bytes() { return sizeof(PdShared); }

int debug_read_profile(unsigned long long *host_out, int reset) {
    if (cudaMemcpyFromSymbol(host_out, g_prof, sizeof(unsigned long long) * 24) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[24] = {0};
        if (cudaMemcpyToSymbol(g_prof, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}

int launch_mincurv_pdip(int B, int n_max, const int32_t *n_pts, double *ws, const Layout &L, const PdipParams &prm,
                        double *alpha, int32_t *status, int32_t *iters, int grid, int *work_counter, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mincurv_pdip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(PdShared));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    cudaMemsetAsync(work_counter, 0, sizeof(int), stream);
    mincurv_pdip_kernel<<<grid, PD_THREADS, sizeof(PdShared), stream>>>(B, n_max, n_pts, ws, L, prm, alpha, status, iters, work_counter);
    return 0;
}

// ================================================================================================
// K2b' -- the full QP of tph.opt_min_curv including the curvature rows |k_ref + E a| <= kappa_bound, for the
// instances whose box-only optimum violates them (status 4 after K2c).  Same Mehrotra iteration; the rows enter
// with slacks s3, s4 (infeasible start allowed) and multipliers l3, l4:
//   M = H + D_box + E^T W E = E^T (I + W) E + D_box,  W = l3/s3 + l4/s4      -> weighted band assembly per iteration
//   rhs = -(f + lu - ll) - tu/su + tl/sl - E^T v,  v = (kl - k_ref) + l3 - l4 + (t3 + l3 rp3)/s3 - (t4 + l4 rp4)/s4
// with kl = k_ref + E a carried incrementally and E, E^T applied in O(N) operator form (mincurv_ops.cuh).
// ================================================================================================
__global__ void __launch_bounds__(PD_THREADS, 4)
mincurv_pdip_kappa_kernel(int B, int n_max, const int32_t *__restrict__ n_pts, double *__restrict__ ws, Layout L,
                          PdipParams prm, double kb, double *__restrict__ alpha_out, int32_t *__restrict__ status,
                          int32_t *__restrict__ iters_out, int *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PdShared &sh = *reinterpret_cast<PdShared *>(smem_raw);
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) { mbar_init(&sh.full_bar[i], 1); mbar_init(&sh.empty_bar[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&sh.aux_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned fill = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh.next = atomicAdd(work_counter, 1);
        __syncthreads();
        const int b = sh.next;
        if (b >= B) break;
        if (status[b] != 4) continue;
        const int n = n_pts ? n_pts[b] : n_max;
        double *aout = alpha_out + (size_t)b * n_max;
        double *slab = ws + (size_t)b * L.stride;
        const double *HB = slab + L.o_hb;
        double *tiles = slab + L.o_tiles;
        const double *__restrict__ LB = vec(slab, L, V_LB), *__restrict__ UB = vec(slab, L, V_UB), *__restrict__ F = vec(slab, L, V_F);
        const double *__restrict__ KR = vec(slab, L, V_KREF);
        double *__restrict__ AL = vec(slab, L, V_ALPHA), *__restrict__ LU = vec(slab, L, V_LU), *__restrict__ LL = vec(slab, L, V_LL);
        double *__restrict__ RHS = vec(slab, L, V_RHS), *__restrict__ DX = vec(slab, L, V_DX), *__restrict__ DD = vec(slab, L, V_DD);
        double *__restrict__ TU = vec(slab, L, V_DLU), *__restrict__ TL = vec(slab, L, V_DLL), *__restrict__ SU = vec(slab, L, V_SU), *__restrict__ SL = vec(slab, L, V_SL);
        double *__restrict__ ISU = vec(slab, L, V_ISU), *__restrict__ ISL = vec(slab, L, V_ISL), *YP = vec(slab, L, V_YPAD);
        double *__restrict__ S3 = vec(slab, L, V_S3), *__restrict__ S4 = vec(slab, L, V_S4), *__restrict__ L3 = vec(slab, L, V_L3), *__restrict__ L4 = vec(slab, L, V_L4);
        double *__restrict__ KL = vec(slab, L, V_KL), *__restrict__ WK = vec(slab, L, V_WK), *__restrict__ EDX = vec(slab, L, V_EDX);
        double *__restrict__ T3 = vec(slab, L, V_T3K), *__restrict__ T4 = vec(slab, L, V_T4K), *__restrict__ VV = vec(slab, L, V_VV), *__restrict__ ETV = vec(slab, L, V_RD);
        double *t0 = vec(slab, L, V_T0), *t1 = vec(slab, L, V_T1), *t2 = vec(slab, L, V_T2), *t3 = vec(slab, L, V_T3), *t4 = vec(slab, L, V_T4), *t5 = vec(slab, L, V_T5);
        const int nb = (n - 32 + 31) / 32;
        const double m4 = 4.0 * n;
        if (threadIdx.x == 0) sh.flag = 0;

        // ---- start: box centre; kl = k_ref + E a; gradient g = E^T (E a) + f ----
        for (int i = threadIdx.x; i < n; i += PD_THREADS) AL[i] = 0.5 * (LB[i] + UB[i]);
        __syncthreads();
        apply_E(slab, L, n, AL, EDX, t0, t1, t2, t3, t4, t5);
        apply_Et(slab, L, n, EDX, ETV, t0, t1, t2, t3, t4, t5);
        double gmax = 0.0, fmaxv = 0.0;
        for (int i = threadIdx.x; i < n; i += PD_THREADS) {
            gmax = fmax(gmax, fabs(ETV[i] + F[i]));
            fmaxv = fmax(fmaxv, fabs(F[i]));
        }
        gmax = block_reduce<1>(gmax, sh.red);
        fmaxv = block_reduce<1>(fmaxv, sh.red);
        const double lam0 = 1e-2 * gmax + 1e-300;
        double musum = 0.0;
        for (int i = threadIdx.x; i < n; i += PD_THREADS) {
            const double gi = ETV[i] + F[i];
            const double lu = fmax(-gi, 0.0) + lam0, ll = fmax(gi, 0.0) + lam0;
            const double a = AL[i], su = UB[i] - a, sl = a - LB[i];
            const double kl = KR[i] + EDX[i];
            const double s3 = fmax(kb - kl, 1e-2 * kb), s4 = fmax(kb + kl, 1e-2 * kb);
            LU[i] = lu; LL[i] = ll; SU[i] = su; SL[i] = sl; ISU[i] = 1.0 / su; ISL[i] = 1.0 / sl;
            KL[i] = kl; S3[i] = s3; S4[i] = s4; L3[i] = lam0; L4[i] = lam0;
            musum += su * lu + sl * ll + (s3 + s4) * lam0;
        }
        musum = block_reduce<0>(musum, sh.red);
        const double mu0 = musum / m4;
        const double rd_tol = prm.rd_rel * (fmaxv + gmax) + 1e-300;
        double mu = mu0;
        int it = 0, result = 2;
        double rdmax = 1e300, rpmax = 1e300;
        for (it = 0; it < prm.max_iter + 20; ++it) {
            // ---- weights, barrier diagonal, affine right-hand side ----
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double s3 = S3[i], s4 = S4[i], l3 = L3[i], l4 = L4[i], kl = KL[i];
                const double rp3 = kl + s3 - kb, rp4 = -kl + s4 - kb;
                WK[i] = l3 / s3 + l4 / s4;
                DD[i] = LU[i] * ISU[i] + LL[i] * ISL[i];
                VV[i] = (kl - KR[i]) + l3 * rp3 / s3 - l4 * rp4 / s4;       // affine: t3 = -s3 l3, t4 = -s4 l4
            }
            __syncthreads();
            assemble_hband(slab, L, n, WK, sh.As);     // the tile area is free between the sweeps and the next factorisation
            __syncthreads();
            apply_Et(slab, L, n, VV, ETV, t0, t1, t2, t3, t4, t5);
            for (int i = threadIdx.x; i < n; i += PD_THREADS) RHS[i] = -F[i] - ETV[i];
            __syncthreads();
            if (!factor(sh, HB, DD, tiles, n, nb)) {
                // E^T W E with W = l/s -> 1e12 and beyond is no longer numerically SPD: accept a late iterate, else give up
                result = (mu <= 1e-7 * mu0 && rdmax <= 1e3 * rd_tol && rpmax <= 1e-6 * kb) ? 0 : 3;
                break;
            }
            fill = solve(sh, tiles, RHS, DX, YP, n, nb, fill);
            apply_E(slab, L, n, DX, EDX, t0, t1, t2, t3, t4, t5);
            // ---- affine step lengths and centring ----
            double rp = 0.0, rdl = 0.0, c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double su = SU[i], sl = SL[i], lu = LU[i], ll = LL[i], dx = DX[i];
                const double s3 = S3[i], s4 = S4[i], l3 = L3[i], l4 = L4[i], kl = KL[i], ed = EDX[i];
                const double p = dx * ISU[i], m = dx * ISL[i];
                const double ds3 = -(kl + s3 - kb) - ed, ds4 = -(-kl + s4 - kb) + ed;
                const double dlu = lu * (p - 1.0), dll = -ll * (1.0 + m);
                const double dl3 = -l3 - l3 * ds3 / s3, dl4 = -l4 - l4 * ds4 / s4;
                rp = fmax(rp, fmax(fmax(p, -m), fmax(-ds3 / s3, -ds4 / s4)));
                rdl = fmax(rdl, fmax(fmax(1.0 - p, 1.0 + m), fmax(-dl3 / l3, -dl4 / l4)));
                c00 += su * lu + sl * ll + s3 * l3 + s4 * l4;
                c01 += su * dlu + sl * dll + s3 * dl3 + s4 * dl4;
                c10 += dx * (ll - lu) + ds3 * l3 + ds4 * l4;
                c11 += dx * (dll - dlu) + ds3 * dl3 + ds4 * dl4;
            }
            rp = block_reduce<1>(rp, sh.red);
            rdl = block_reduce<1>(rdl, sh.red);
            double ap = (rp > 1.0) ? 1.0 / rp : 1.0, ad = (rdl > 1.0) ? 1.0 / rdl : 1.0;
            c00 = block_reduce<0>(c00, sh.red); c01 = block_reduce<0>(c01, sh.red);
            c10 = block_reduce<0>(c10, sh.red); c11 = block_reduce<0>(c11, sh.red);
            const double mua = (c00 + ad * c01 + ap * c10 + ap * ad * c11) / m4;
            double sigma = mua / mu;
            sigma = sigma * sigma * sigma;
            const double smu = sigma * mu;
            // ---- corrector right-hand side ----
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double su = SU[i], sl = SL[i], lu = LU[i], ll = LL[i], dx = DX[i], isu = ISU[i], isl = ISL[i];
                const double s3 = S3[i], s4 = S4[i], l3 = L3[i], l4 = L4[i], kl = KL[i], ed = EDX[i];
                const double rp3 = kl + s3 - kb, rp4 = -kl + s4 - kb;
                const double ds3 = -rp3 - ed, ds4 = -rp4 + ed;
                const double dlu = lu * (dx * isu - 1.0), dll = -ll * (1.0 + dx * isl);
                const double dl3 = -l3 - l3 * ds3 / s3, dl4 = -l4 - l4 * ds4 / s4;
                const double tu = smu - su * lu + dx * dlu, tl = smu - sl * ll - dx * dll;
                const double q3 = smu - s3 * l3 - ds3 * dl3, q4 = smu - s4 * l4 - ds4 * dl4;
                TU[i] = tu; TL[i] = tl; T3[i] = q3; T4[i] = q4;
                VV[i] = (kl - KR[i]) + l3 - l4 + (q3 + l3 * rp3) / s3 - (q4 + l4 * rp4) / s4;
                RHS[i] = -(F[i] + lu - ll) - tu * isu + tl * isl;
            }
            __syncthreads();
            apply_Et(slab, L, n, VV, ETV, t0, t1, t2, t3, t4, t5);
            for (int i = threadIdx.x; i < n; i += PD_THREADS) RHS[i] -= ETV[i];
            __syncthreads();
            fill = solve(sh, tiles, RHS, DX, YP, n, nb, fill);
            apply_E(slab, L, n, DX, EDX, t0, t1, t2, t3, t4, t5);
            rp = 0.0; rdl = 0.0;
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double lu = LU[i], ll = LL[i], dx = DX[i], isu = ISU[i], isl = ISL[i];
                const double s3 = S3[i], s4 = S4[i], l3 = L3[i], l4 = L4[i], kl = KL[i], ed = EDX[i];
                const double ds3 = -(kl + s3 - kb) - ed, ds4 = -(-kl + s4 - kb) + ed;
                const double dlu = (TU[i] + lu * dx) * isu, dll = (TL[i] - ll * dx) * isl;
                const double dl3 = (T3[i] - l3 * ds3) / s3, dl4 = (T4[i] - l4 * ds4) / s4;
                rp = fmax(rp, fmax(fmax(dx * isu, -dx * isl), fmax(-ds3 / s3, -ds4 / s4)));
                rdl = fmax(rdl, fmax(fmax(-dlu / lu, -dll / ll), fmax(-dl3 / l3, -dl4 / l4)));
            }
            rp = block_reduce<1>(rp, sh.red);
            rdl = block_reduce<1>(rdl, sh.red);
            ap = (prm.eta < rp) ? prm.eta / rp : 1.0;
            ad = (prm.eta < rdl) ? prm.eta / rdl : 1.0;
            double musum2 = 0.0;
            rpmax = 0.0;
            for (int i = threadIdx.x; i < n; i += PD_THREADS) {
                const double su = SU[i], sl = SL[i], lu = LU[i], ll = LL[i], dx = DX[i];
                const double s3 = S3[i], s4 = S4[i], l3 = L3[i], l4 = L4[i], kl = KL[i], ed = EDX[i];
                const double ds3 = -(kl + s3 - kb) - ed, ds4 = -(-kl + s4 - kb) + ed;
                const double dlu = (TU[i] + lu * dx) * ISU[i], dll = (TL[i] - ll * dx) * ISL[i];
                const double dl3 = (T3[i] - l3 * ds3) / s3, dl4 = (T4[i] - l4 * ds4) / s4;
                const double sun = su - ap * dx, sln = sl + ap * dx, s3n = s3 + ap * ds3, s4n = s4 + ap * ds4;
                const double lun = lu + ad * dlu, lln = ll + ad * dll, l3n = l3 + ad * dl3, l4n = l4 + ad * dl4;
                const double kln = kl + ap * ed;
                AL[i] += ap * dx; SU[i] = sun; SL[i] = sln; ISU[i] = 1.0 / sun; ISL[i] = 1.0 / sln;
                LU[i] = lun; LL[i] = lln; S3[i] = s3n; S4[i] = s4n; L3[i] = l3n; L4[i] = l4n; KL[i] = kln;
                VV[i] = (kln - KR[i]) + l3n - l4n;
                musum2 += sun * lun + sln * lln + s3n * l3n + s4n * l4n;
                rpmax = fmax(rpmax, fmax(fabs(kln + s3n - kb), fabs(-kln + s4n - kb)));
            }
            mu = block_reduce<0>(musum2, sh.red) / m4;
            rpmax = block_reduce<1>(rpmax, sh.red);
            __syncthreads();
            // dual residual r_d = E^T (kl - k_ref + l3 - l4) + f + lu - ll   (exact every iteration, O(N))
            apply_Et(slab, L, n, VV, ETV, t0, t1, t2, t3, t4, t5);
            rdmax = 0.0;
            for (int i = threadIdx.x; i < n; i += PD_THREADS) rdmax = fmax(rdmax, fabs(ETV[i] + F[i] + LU[i] - LL[i]));
            rdmax = block_reduce<1>(rdmax, sh.red);
            if (mu <= prm.mu_rel * mu0 && rdmax <= rd_tol && rpmax <= 1e-8 * kb) { result = 0; ++it; break; }
            if (mu <= 1e-2 * prm.mu_rel * mu0) { result = (rdmax <= 1e3 * rd_tol && rpmax <= 1e-6 * kb) ? 0 : 2; ++it; break; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_max; i += PD_THREADS) aout[i] = (i < n) ? AL[i] : 0.0;
        if (threadIdx.x == 0) {
            status[b] = result;
            if (iters_out) iters_out[b] += it;
        }
    }
}

int launch_mincurv_pdip_kappa(int B, int n_max, const int32_t *n_pts, double *ws, const Layout &L, const PdipParams &prm,
                              double kappa_bound, double *alpha, int32_t *status, int32_t *iters, int grid, int *work_counter,
                              cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mincurv_pdip_kappa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(PdShared));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    cudaMemsetAsync(work_counter, 0, sizeof(int), stream);
    mincurv_pdip_kappa_kernel<<<grid, PD_THREADS, sizeof(PdShared), stream>>>(B, n_max, n_pts, ws, L, prm, kappa_bound, alpha,
                                                                             status, iters, work_counter);
    return 0;
}

}  // namespace mc
