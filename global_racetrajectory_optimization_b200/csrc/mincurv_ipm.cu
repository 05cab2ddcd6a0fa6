// K2b -- box-constrained QP solve  min 1/2 a^T H a + f^T a,  lb <= a <= ub  by a Mehrotra predictor-corrector
// primal-dual interior-point method (replaces quadprog.solve_qp inside tph.opt_min_curv, call site
// /root/reference/main_globaltraj.py:264-271), and K2b' -- the same iteration with tph's curvature rows.
//
// One CTA of two warps per QP instance, eight CTAs per SM.  H is the cyclic band (half-bandwidth 32) assembled by K2a.
// Every interior-point iteration factorises M = H + D (D diagonal, from the barrier) as a bordered LDL^T:
//   chain nodes 0..NA-1 (NA = n - 32):  A = M[chain, chain] = L D L^T   (banded, L unit lower, 32 sub-diagonals)
//   separator = the last 32 nodes (they close the cycle):  G = Y L^-T,  Y = M[sep, chain]  (the fill row of the separator)
//                                                          S = M[sep, sep] - G D^-1 G^T = L_S D_S L_S^T
// column by column, right-looking, with the whole active window in REGISTERS:
//   warp 0 (chain): lane = row (rows k+1..k+32 of the window, circular), 32 accumulators per lane = the updates of
//           the row's 32 window columns.  Step k: v = A[row][k] - acc[0]; the column is broadcast through shared memory;
//           w = 1/d_k; l = v w; acc[c] <- acc[c+1] + l v_c (32 independent DFMAs: the register file is the sliding window,
//           the slide is the operand index, no moves).  The forward substitution of the predictor's right-hand side rides
//           along (one more accumulator).  The pivot chain per column is DADD -> STS/LDS -> rcp -> DMUL -> DFMA.
//   warp 1 (fill):  lane = separator row, 32 accumulators = the updates of G[row][k+1..k+32]; consumes warp 0's columns
//           from a shared-memory ring eight columns behind, and accumulates S -= (G w) G^T on the FP64 tensor cores
//           (mma.sync.m8n8k4.f64, SASS DMMA) once per eight columns.
// The first version of this kernel ran a 32x32-block Cholesky on DMMA with explicit block inverses: 10 % of the fp64
// pipe, latency-bound on the in-block pivot chains and on the hand-offs between three warp roles
// (profiles/r01_final_pdip_ncu_summary.json).  Here every step offers 64 independent DFMAs per instance and eight
// instances share an SM, so the fp64 pipe and the issue slots are what is busy.
//
// The factor (34 doubles per column of L and of G: 32 entries + the sweeps' per-column scalars) goes to the instance's HBM slab and
// is streamed back by the triangular sweeps through a shared-memory ring of 8-column units filled by cp.async.bulk
// (TMA, 1-D) on mbarriers -- the band rows of H reach warp 0 the same way.  Per iteration: factor written once, read
// three times (the predictor's forward sweep is fused into the factorisation).
#include <cstdio>
#include <cstdlib>
#include "mincurv_ops.cuh"

namespace mc {

constexpr unsigned FULL = 0xffffffffu;
constexpr int IP_THREADS = 64;
constexpr int SUB = 8;                 // columns per hand-off / streaming unit
constexpr int VBP = 12;                // pitch of the row-major panel buffers (conflict-free DMMA fragment loads, 16-byte rows)
constexpr int FROW = 34;               // pitch of a fill row: [g (32), z, w]
constexpr int HB_SLOTS = 2;            // band-row units of warp 0 (the next panel's is in flight)
constexpr int HO_SLOTS = 2;            // panels between warp 0 and warp 1
constexpr int LROW = 42;               // pitch of a chain-factor row: [Q - I (8), L21 (32), t, w]
#ifndef MC_LT_SLOTS
#define MC_LT_SLOTS 3
#endif
constexpr int LT_SLOTS = MC_LT_SLOTS;  // sweep rings: warp 0 streams the chain rows (LT), warp 1 the fill rows (GT), concurrently
constexpr int GT_SLOTS = 6 - LT_SLOTS;
constexpr int RING_UNITS = LT_SLOTS + GT_SLOTS;       // mbarriers: [0, LT_SLOTS) warp 0, [LT_SLOTS, RING_UNITS) warp 1
constexpr int QDEPTH = 16;             // units of z (forward) / t (backward) in flight between the two warps
constexpr int LT_UNIT_DOUBLES = SUB * LROW;           // 336
constexpr int GT_UNIT_DOUBLES = SUB * FROW;           // 272
constexpr unsigned HB_UNIT_BYTES = SUB * HB_PITCH * sizeof(double);   // 2176

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// ---- TMA (bulk async copy) + mbarrier helpers ----
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
#ifndef MC_WAIT_MODE
#define MC_WAIT_MODE 1
#endif
#ifndef MC_RELAXED_BACKOFF_NS
#define MC_RELAXED_BACKOFF_NS 250
#endif

// MC_WAIT_MODE 0: try_wait with a long suspend hint (the thread is parked until the phase completes);
//              1: test_wait spin (non-blocking probe): the waiter resumes within a few cycles of the arrival.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
#if MC_WAIT_MODE == 0
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
                 " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#else
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}
// same, for the warp that is normally AHEAD of its producer (the fill warp, ~20 % of its time): parked by the hardware
// (try_wait with a suspend hint) instead of probing -- the probes of the spin version were 16 % of all executed instructions
// and took issue slots from the chain warps of the same scheduler; runs with them fell into a slow mode now and then.
// -DMC_RELAXED_SPIN_NS=<ns> restores the probing loop with that back-off, for experiments.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, unsigned parity) {
#ifndef MC_RELAXED_SPIN_NS
    // (one try_wait parks the warp for well under a microsecond before it reports "not yet": the back-off between two
    //  of them keeps the probes of a ~1400-cycle wait at a handful instead of ~200)
    unsigned ok = 0;
    for (;;) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(MC_RELAXED_BACKOFF_NS);
    }
#else
    unsigned ok = 0;
    for (;;) {
        asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(MC_RELAXED_SPIN_NS);
    }
#endif
}
// The factor is a stream (written once, read three times per iteration, 0.5 MB per instance, far beyond what L2 can
// keep for 1184 resident instances): its copies and stores carry an evict-first L2 policy so that they do not push the
// O(N) iterate vectors of the interior-point loop out of L2.
template <typename T>
__device__ __forceinline__ void st_stream(T *p, T v) { __stcs(p, v); }      // factor rows: streaming (evict-first) stores
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

#ifdef MC_PROFILE
__device__ unsigned long long g_prof[24];
#define PROF_T0(name) const long long name = clock64()
#define PROF_ADD(slot, t0) do { if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_prof[slot], (unsigned long long)(clock64() - (t0))); } while (0)
#define PROF_ADD1(slot, t0) do { if (blockIdx.x == 0 && threadIdx.x == 32) atomicAdd(&g_prof[slot], (unsigned long long)(clock64() - (t0))); } while (0)
__device__ unsigned long long g_seg[32];
#ifdef MC_PROFILE_SEG      // per-panel sub-phase marks: ~200 cycles each, only for attributing time inside one panel
#define SEG_BEGIN() long long _seg_t = clock64()
#define SEG(slot) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { const long long _n = clock64(); atomicAdd(&g_seg[slot], (unsigned long long)(_n - _seg_t)); _seg_t = _n; } } while (0)
#else
#define SEG_BEGIN() do { } while (0)
#define SEG(slot) do { } while (0)
#endif
#else
#define SEG_BEGIN() do { } while (0)
#define SEG(slot) do { } while (0)
#define PROF_T0(name) do { } while (0)
#define PROF_ADD(slot, t0) do { } while (0)
#define PROF_ADD1(slot, t0) do { } while (0)
#endif

// 1/d for a positive, normal d: hardware seed (rcp.approx.f64, ~20 bits) + two Newton steps (relative error ~1e-16).
// Four dependent DFMAs instead of the IEEE division routine with its special-case path: the pivot chain of the
// factorisation is latency bound on exactly this.
__device__ __forceinline__ double fast_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}
// x with its high word ANDed with m (m = 0: a denormal of magnitude < 2^-1022, i.e. zero for every purpose here;
// m = ~0: x).  One LOP3 instead of a 64-bit select: resets the accumulators of the lane whose row enters the window.

// one panel (eight columns) handed from warp 0 to warp 1
struct Handoff {
    double vb[32 * VBP];       // L of the window rows k0+8 .. k0+39 x the panel's 8 columns, row-major (DMMA fragments)
    double l11[64];            // the panel's unit-lower diagonal block, row-major [m][j] (m > j used)
    double w[8], y[8];         // 1/d and the forward-substituted right-hand side of the panel's columns
};

struct IpShared {
    union {
        struct {
            double hb[HB_SLOTS][SUB * HB_PITCH];      // band rows of H staged for warp 0 (TMA)
            Handoff ho[HO_SLOTS];                     // warp 0 -> warp 1
            double gb[32 * VBP];                      // warp 1: the panel of the fill rows, row-major (DMMA fragments)
        } f;
        struct {
            double lt[LT_SLOTS][LT_UNIT_DOUBLES];     // sweeps: streamed chain units (warp 0)
            double gt[GT_SLOTS][GT_UNIT_DOUBLES];     //         streamed fill units (warp 1)
            double q[QDEPTH * SUB];                   //         z (forward) / t (backward) handed between the warps
        } sw;
        double win[hband_win_doubles(IP_THREADS)];    // K2b': scratch of the weighted band assembly
    } u;
    union {
        double Ss[32 * 33];                           // separator block -> its L_S (strictly lower, in place)
        double sfrag[20 * 32];                        // during the chain: S accumulators (DMMA C fragments, [block][lane][2])
    } s;
    double wS[32], gs[32], part[2][32];
    double red[32];
    uint64_t hb_full[HB_SLOTS], ho_full[HO_SLOTS], ho_empty[HO_SLOTS], ring_full[RING_UNITS];
    unsigned ring_phase[2];                           // parity bit per ring slot, one word per warp
    int prog[4];                                      // sweep progress (units done): forward w0, w1; backward w0, w1
    int flag;
    int next;                                         // next instance index (dynamic work distribution)
};

// pointers into the instance slab that the factorisation and the sweeps use.  Factor rows in HBM, one bulk copy per unit
// of eight rows (every row 16-byte aligned):
//   LT[k0+j] = [ (Q - I)[0..7][j], L21[0..31][j], -, w_k ]   (pitch 42) the chain factor, one panel of columns k0 .. k0+7
//              at a time: L21 = the 32 rows of L below the panel, Q = L11^-1 the inverse of the panel's unit-lower block, so
//              that the panel's triangular solve in the sweeps is a mat-vec too (forward: y1 = Q a1, a2 -= L21 y1;
//              backward: u = t1 - L21^T x2, x1 = Q^T u).  t = (y - G^T x_S) w: right-hand side of the backward sweep, w = 1/d_k
//   GT[k]    = [ G[0 .. 31][k],  z_k,  w_k ]                  (pitch 34) z = w y: what the separator's forward part needs
// so the sweeps read nothing but the streamed units (no per-column global loads on the serial chains).  Columns
// NA .. 8 ceil(NA / 8) - 1 are padding (pivot 1, nothing else): every unit is a full panel.
struct Factor {
    const double *HB;      // band of H, row i: H[i][i .. i+32], [33]: pivot H_ii + D_i (written by factor())
    const double *DD;      // barrier diagonal
    double *LT, *GT;
    int n, NA;
};

__device__ __forceinline__ Factor make_factor(double *slab, const Layout &L, int n) {
    Factor F;
    F.HB = slab + L.o_hb;
    F.DD = vec(slab, L, V_DD);
    F.LT = slab + L.o_tiles;
    F.GT = F.LT + (size_t)L.np * LROW;
    F.n = n;
    F.NA = n - 32;
    return F;
}

__device__ __forceinline__ int blk(int I, int J) { return (I * (I + 1)) / 2 + J; }      // lower 8x8 block (I >= J) of a 4x4 block grid

// ---- warp 0: LDL^T of the chain, one panel of eight columns at a time, with the forward substitution of g fused in.
//      The 32x32 window of accumulated updates (rows/columns k0 .. k0+31, lower blocks) lives in DMMA C fragments.
//      Panel step: (1) block column 0 -> row layout (lane l = row k0 + l; lanes 0..7 also carry the entering row
//      k0 + 32 + l); (2) eight pivots: d_j by shuffle from lane j, w = 1/d, l = v w, v_{j',j} by shuffle from lane j'
//      for the in-panel updates; (3) trailing update W'[I][J] = W[I+1][J+1] + (L d)(L)^T on the tensor cores, which also
//      slides the window by one block (the block row of the eight entering rows starts from zero). ----
__device__ __noinline__ bool factor_chain(IpShared &sh, const double *__restrict__ HBp, double *__restrict__ LTp, double *__restrict__ GTp, const int NA, const double *__restrict__ g, unsigned tick) {
    const int lane = threadIdx.x & 31;
    const int gq = lane >> 2, q = lane & 3;
    const int nunits = (NA + SUB - 1) / SUB;
    const uint64_t pol = l2_evict_first_policy();
    double W[10][2];
#pragma unroll
    for (int b = 0; b < 10; ++b) { W[b][0] = 0.0; W[b][1] = 0.0; }
    // right-hand side: gv = g - (L y so far) of row k0 + lane; the entering rows take their g from a block of 32 loaded a
    // block ahead (a load issued inside the panel step would be waited for right away: the scoreboard is per warp)
    double gv = (lane < NA) ? g[lane] : 0.0;
    double gcur = (32 + lane < NA) ? g[32 + lane] : 0.0;
    double gnext = (64 + lane < NA) ? g[64 + lane] : 0.0;
    bool ok = true;
    if (lane == 0) {
        mbar_expect_tx(&sh.hb_full[tick % HB_SLOTS], HB_UNIT_BYTES);
        tma_load_1d(sh.u.f.hb[tick % HB_SLOTS], HBp, HB_UNIT_BYTES, &sh.hb_full[tick % HB_SLOTS], pol);
    }
    SEG_BEGIN();
    for (int t = 0; t < nunits; ++t) {
        const unsigned ht = tick + t, hs = ht % HB_SLOTS, ls = ht % HO_SLOTS;
        const int k0 = t * SUB;
        SEG(8);
        if (t > 0 && (t & 3) == 0) {                  // a new block of 32 rows enters over the next four panels
            gcur = gnext;
            const int idx = k0 + 64 + lane;
            gnext = (idx < NA) ? g[idx] : 0.0;
        }
        if (lane == 0 && t + 1 < nunits) {            // band rows of the next panel (slot of panel t-1: all lanes are past it)
            const unsigned h2 = (ht + 1) % HB_SLOTS;
            mbar_expect_tx(&sh.hb_full[h2], HB_UNIT_BYTES);
            tma_load_1d(sh.u.f.hb[h2], HBp + (size_t)(t + 1) * SUB * HB_PITCH, HB_UNIT_BYTES, &sh.hb_full[h2], pol);
        }
        PROF_T0(tw1);
        if (ht >= (unsigned)HO_SLOTS) mbar_wait(&sh.ho_empty[ls], ((ht / HO_SLOTS) - 1u) & 1u);
        PROF_ADD(19, tw1);
        Handoff &ho = sh.u.f.ho[ls];
        SEG(9);
        // ---- (1) block column 0 of the window -> row layout, through the (still free) fragment area of the hand-off slot
#pragma unroll
        for (int I = 0; I < 4; ++I)
            *reinterpret_cast<double2 *>(&ho.vb[(8 * I + gq) * VBP + 2 * q]) = make_double2(W[blk(I, 0)][0], W[blk(I, 0)][1]);
        __syncwarp();
        double p[8], p2[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const double2 uu = *reinterpret_cast<const double2 *>(&ho.vb[lane * VBP + j]);
            p[j] = uu.x; p[j + 1] = uu.y;
        }
        PROF_T0(tw0);
        mbar_wait(&sh.hb_full[hs], (ht / HB_SLOTS) & 1u);
        PROF_ADD(14, tw0);
        const double *hbg = sh.u.f.hb[hs];
        const bool row_ok = (k0 + lane < NA), row2_ok = (lane < 8) && (k0 + 32 + lane < NA);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int d = lane - j;                    // row k0 + lane, column k0 + j: band entry H[k0+j][d]; d = 0: the pivot (with D)
            const bool col_ok = (k0 + j < NA);
            double a1 = 0.0;
            if (d >= 0 && row_ok && col_ok) a1 = hbg[j * HB_PITCH + ((d == 0) ? HB_PITCH - 1 : d)];
            if (d == 0 && !col_ok) a1 = 1.0;           // padding column: unit pivot
            p[j] = (d >= 0) ? a1 - p[j] : 0.0;
            const int d2 = 32 + lane - j;              // entering row k0 + 32 + lane: nonzero for j >= lane only, no updates yet
            p2[j] = (row2_ok && col_ok && d2 <= 32) ? hbg[j * HB_PITCH + d2] : 0.0;
        }
        double gv2 = __shfl_sync(FULL, gcur, ((t & 3) << 3) + (lane & 7));
        if (lane >= 8) gv2 = 0.0;
        __syncwarp();                                  // every lane has read its row of block column 0: the slot can take L now
        SEG(10);
        // ---- (2) the panel ----
        double dsave = 1.0, wsave = 1.0, ysave = 0.0;
        double *ltp = LTp + (size_t)k0 * LROW;
        const int wr = (lane >= 8) ? lane - 8 : 24 + lane;      // window row (after the slide) of the row this lane hands over
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double dj = __shfl_sync(FULL, p[j], j);
            const double yj = __shfl_sync(FULL, gv, j);
            if (!(dj > 0.0)) ok = false;
            const double w = fast_rcp(dj);
            const double lt = p[j] * w, lt2 = p2[j] * w;           // (lanes <= j: lt is not an entry of L; lanes > j: lt2 = 0)
            ho.vb[wr * VBP + j] = (lane >= 8) ? lt : lt2;
            st_stream(ltp + j * LROW + 8 + wr, (lane >= 8) ? lt : lt2);      // L21: the 32 rows below the panel
            if (lane < 8 && lane > j) ho.l11[lane * 8 + j] = lt;
            if (lane == j) { dsave = dj; wsave = w; ysave = yj; }
            gv = fma(-lt, yj, gv);                                 // (lanes <= j: gv is dead)
            gv2 = fma(-lt2, yj, gv2);
#pragma unroll
            for (int jj = j + 1; jj < 8; ++jj) {
                const double vj = __shfl_sync(FULL, p[j], jj);     // v of row k0 + jj in column j (unscaled)
                p[jj] = fma(-lt, vj, p[jj]);
                p2[jj] = fma(-lt2, vj, p2[jj]);
            }
        }
        SEG(11);
        if (lane < 8) {
            ho.w[lane] = wsave;
            ho.y[lane] = ysave;
        }
        __syncwarp();
        // ---- (2b) Q = L11^-1 (the panel's unit-lower block): with it the panel's triangular solves in the sweeps are mat-vecs.
        //      LT[k0+j] = [ (Q - I)[0..7][j], L21[0..31][j] (stored above), t, w ] ----
        {
            const int jq = lane & 7;
            double q8[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                double sacc = (m == jq) ? 1.0 : 0.0;
#pragma unroll
                for (int i = 0; i < m; ++i) sacc = fma(-ho.l11[m * 8 + i], q8[i], sacc);
                q8[m] = sacc;
            }
            if (lane < 8) {
                double *ltq = LTp + (size_t)(k0 + lane) * LROW;
#pragma unroll
                for (int m = 0; m < 8; m += 2)
                    st_stream(reinterpret_cast<double2 *>(ltq + m), make_double2((m == lane) ? 0.0 : q8[m], (m + 1 == lane) ? 0.0 : q8[m + 1]));
                st_stream(ltq + 41, wsave);
            }
        }
        // ---- (3) trailing update on the tensor cores; the window slides by one block ----
        double af[4][2], bf[4][2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const double dk = __shfl_sync(FULL, dsave, 4 * ks + q);
#pragma unroll
            for (int I = 0; I < 4; ++I) {
                bf[I][ks] = ho.vb[(8 * I + gq) * VBP + 4 * ks + q];
                af[I][ks] = bf[I][ks] * dk;
            }
        }
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J) {
                double c[2] = {0.0, 0.0};
                if (I < 3) { c[0] = W[blk(I + 1, J + 1)][0]; c[1] = W[blk(I + 1, J + 1)][1]; }
                dmma(c, af[I][0], bf[J][0]);
                dmma(c, af[I][1], bf[J][1]);
                W[blk(I, J)][0] = c[0]; W[blk(I, J)][1] = c[1];
            }
        // rows slide by eight as well
        {
            const double up = __shfl_sync(FULL, gv, (lane + 8) & 31), en = __shfl_sync(FULL, gv2, (lane - 24) & 31);
            gv = (lane < 24) ? up : en;
        }
        __syncwarp();
        SEG(12);
        if (lane == 0) mbar_arrive(&sh.ho_full[ls]);
    }
    return ok;
}

// ---- warp 1: fill rows G = Y L^-T (lane = separator row), blocked like the chain: the 32 x 32 window of G's updates in
//      DMMA C fragments, panel by forward substitution with the panel's unit-lower block, trailing update
//      G'[:, J] = G[:, J+1] + G_panel L^T and S -= (G_panel w) G_panel^T on the tensor cores; also the separator part of
//      the fused forward substitution  gS -= G (w y) ----
__device__ __noinline__ void factor_fill(IpShared &sh, const double *__restrict__ HBp, double *__restrict__ LTp, double *__restrict__ GTp, const int NA, const double *__restrict__ g, unsigned tick) {
    const int lane = threadIdx.x & 31;
    const int gq = lane >> 2, q = lane & 3;
    const int nunits = (NA + SUB - 1) / SUB;
    double G[4][4][2];
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
        for (int J = 0; J < 4; ++J) { G[I][J][0] = 0.0; G[I][J][1] = 0.0; }
    double gsacc = 0.0;
    double *gb = sh.u.f.gb;
    for (int t = 0; t < nunits; ++t) {
        const unsigned ht = tick + t, ls = ht % HO_SLOTS;
        const int k0 = t * SUB;
        // ---- column block 0 of the window -> row layout ----
#pragma unroll
        for (int I = 0; I < 4; ++I)
            *reinterpret_cast<double2 *>(&gb[(8 * I + gq) * VBP + 2 * q]) = make_double2(G[I][0][0], G[I][0][1]);
        __syncwarp();
        double gp[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const double2 uu = *reinterpret_cast<const double2 *>(&gb[lane * VBP + j]);
            gp[j] = uu.x; gp[j + 1] = uu.y;
        }
        const bool hasY = (k0 < 32) || (k0 + SUB - 1 >= NA - 32);     // Y = M[sep, chain] is nonzero across the wrap and next to the separator
        if (hasY) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + j;
                double yv = 0.0;
                if (k < NA) {
                    if (k <= lane) yv = HBp[(size_t)(NA + lane) * HB_PITCH + (k + 32 - lane)];
                    else if (k >= NA + lane - 32) yv = HBp[(size_t)k * HB_PITCH + (NA + lane - k)];
                }
                gp[j] = yv - gp[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) gp[j] = -gp[j];
        }
        __syncwarp();
        PROF_T0(tw2);
        mbar_wait_relaxed(&sh.ho_full[ls], (ht / HO_SLOTS) & 1u);
        PROF_ADD1(20, tw2);
        const Handoff &ho = sh.u.f.ho[ls];
        // ---- the panel: g_j -= sum_{m<j} g_m L[k0+j][k0+m] ----
#pragma unroll
        for (int m = 0; m < 7; ++m)                 // right-looking: gp[m] is final, the updates of one column are independent
#pragma unroll
            for (int j = m + 1; j < 8; ++j) gp[j] = fma(-gp[m], ho.l11[j * 8 + m], gp[j]);
        double *gtp = GTp + (size_t)k0 * FROW;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            st_stream(gtp + j * FROW + lane, gp[j]);
            gsacc = fma(gp[j], ho.w[j] * ho.y[j], gsacc);
        }
        if (lane < 8) {
            const double w = ho.w[lane];
            *reinterpret_cast<double2 *>(gtp + lane * FROW + 32) = make_double2(w * ho.y[lane], w);      // [z, w]
        }
#pragma unroll
        for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2 *>(&gb[lane * VBP + j]) = make_double2(gp[j], gp[j + 1]);
        __syncwarp();
        PROF_T0(tw3);
        // ---- trailing update: G'[I][J] = G[I][J+1] + G_panel[I] L[J]^T ----
        double af[4][2], bf[4][2], wq[2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            wq[ks] = ho.w[4 * ks + q];
#pragma unroll
            for (int I = 0; I < 4; ++I) {
                af[I][ks] = gb[(8 * I + gq) * VBP + 4 * ks + q];
                bf[I][ks] = ho.vb[(8 * I + gq) * VBP + 4 * ks + q];
            }
        }
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J < 4; ++J) {
                double c[2] = {0.0, 0.0};
                if (J < 3) { c[0] = G[I][J + 1][0]; c[1] = G[I][J + 1][1]; }
                dmma(c, af[I][0], bf[J][0]);
                dmma(c, af[I][1], bf[J][1]);
                G[I][J][0] = c[0]; G[I][J][1] = c[1];
            }
        // ---- S += (G_panel w) G_panel^T (lower blocks; subtracted from M[sep, sep] at the end) ----
#pragma unroll
        for (int I = 0; I < 4; ++I) {
            double c2[4][2];
#pragma unroll
            for (int J = 0; J <= I; ++J) {
                const double2 cc = *reinterpret_cast<const double2 *>(&sh.s.sfrag[(blk(I, J) * 32 + lane) * 2]);
                c2[J][0] = cc.x; c2[J][1] = cc.y;
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const double a = af[I][ks] * wq[ks];
#pragma unroll
                for (int J = 0; J <= I; ++J) dmma(c2[J], a, af[J][ks]);
            }
#pragma unroll
            for (int J = 0; J <= I; ++J)
                *reinterpret_cast<double2 *>(&sh.s.sfrag[(blk(I, J) * 32 + lane) * 2]) = make_double2(c2[J][0], c2[J][1]);
        }
        __syncwarp();
        PROF_ADD1(21, tw3);
        if (lane == 0) mbar_arrive(&sh.ho_empty[ls]);
    }
    sh.gs[lane] = g[NA + lane] - gsacc;
}

// ---- factorisation of M = H + D with the forward substitution of g; returns false on a non-positive pivot.
//      `tick` counts the hand-off units of this CTA so far (slot / parity bookkeeping of the mbarrier rings); the caller
//      advances it by factor_units(n) afterwards. ----
__device__ __forceinline__ unsigned factor_units(int n) { return (unsigned)((n - 32 + SUB - 1) / SUB); }
__device__ __noinline__ bool factor(IpShared &sh, double *slab, const Layout &L, int n, const double *g, unsigned tick) {
    const Factor F = make_factor(slab, L, n);
    const int NA = F.NA;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *HBw = slab + L.o_hb;
#pragma unroll 4
    for (int i = threadIdx.x; i < NA; i += IP_THREADS) HBw[(size_t)i * HB_PITCH + HB_PITCH - 1] = HBw[(size_t)i * HB_PITCH] + F.DD[i];
    for (int e = threadIdx.x; e < 20 * 32; e += IP_THREADS) sh.s.sfrag[e] = 0.0;
    fence_proxy_async();        // the pivots above (generic proxy) are read by the bulk copies (async proxy)
    __syncthreads();
    PROF_T0(tc0);
    if (warp == 0) {
        if (!factor_chain(sh, F.HB, F.LT, F.GT, NA, g, tick)) sh.flag = 1;
        PROF_ADD(1, tc0);
    } else {
        factor_fill(sh, F.HB, F.LT, F.GT, NA, g, tick);
        PROF_ADD1(13, tc0);
    }
    PROF_T0(tc1);
    // ---- separator: S = M[sep, sep] + D_S - G D^-1 G^T, LDL^T in place (warp 1) ----
    double sf[20];
    if (warp == 1) {
#pragma unroll
        for (int e = 0; e < 20; ++e) sf[e] = sh.s.sfrag[((e >> 1) * 32 + lane) * 2 + (e & 1)];
    }
    // (the separator block of M: 16 entries per thread, all loads in flight before the barrier)
    double sv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int e = threadIdx.x + i * IP_THREADS;
        const int r = e >> 5, c = e & 31;
        const int lo = min(r, c), dist = abs(r - c);
        sv[i] = F.HB[(size_t)(NA + lo) * HB_PITCH + dist];
        if (r == c) sv[i] += F.DD[NA + r];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int e = threadIdx.x + i * IP_THREADS;
        sh.s.Ss[(e >> 5) * 33 + (e & 31)] = sv[i];
    }
    __syncthreads();
    if (warp == 1) {
        const int gq = lane >> 2, q = lane & 3;
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J) {
                const int bi = (I * (I + 1)) / 2 + J;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r = 8 * I + gq, c = 8 * J + 2 * q + e;
                    sh.s.Ss[r * 33 + c] -= sf[2 * bi + e];
                    if (I != J) sh.s.Ss[c * 33 + r] -= sf[2 * bi + e];
                }
            }
        __syncwarp();
        // right-looking LDL^T in place in shared memory, lane = row (the full square is kept up to date, so row k is also
        // column k).  A loop, not 32 unrolled steps on a register row: that version was 51 KB of SASS run once per
        // factorisation (same speed, a quarter of the kernel's instruction footprint).
        bool ok = true;
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
            const double d = sh.s.Ss[k * 33 + k];
            if (!(d > 0.0)) ok = false;
            const double w = fast_rcp(d);
            const double l = sh.s.Ss[lane * 33 + k] * w;
            if (lane > k) {
#pragma unroll 4
                for (int c = k + 1; c < 32; ++c) sh.s.Ss[lane * 33 + c] = fma(-l, sh.s.Ss[k * 33 + c], sh.s.Ss[lane * 33 + c]);
                sh.s.Ss[lane * 33 + k] = l;
            }
            if (lane == k) sh.wS[k] = w;
            __syncwarp();
        }
        if (!ok) sh.flag = 1;
    }
    __syncthreads();
    PROF_ADD(5, tc1);
    return sh.flag == 0;
}

// ---- sweep ring: one warp consumes 8-column units streamed HBM -> shared by bulk copies it issues itself ----
template <int SLOTS, int BASE, int UNIT_DOUBLES, int WHO>
struct RingT {
    IpShared &sh;
    double *buf;
    unsigned phase;
    uint64_t pol;
    __device__ RingT(IpShared &s, double *b) : sh(s), buf(b), phase(s.ring_phase[WHO]), pol(l2_evict_first_policy()) {}
    // unit -> slot `sl` (the callers walk the slots cyclically: no modulo on the serial paths)
    __device__ __forceinline__ void issue(int unit, int sl, const double *rows) {
        const unsigned bytes = (unsigned)(UNIT_DOUBLES * sizeof(double));
        mbar_expect_tx(&sh.ring_full[BASE + sl], bytes);
        tma_load_1d(buf + sl * UNIT_DOUBLES, rows + (size_t)unit * UNIT_DOUBLES, bytes, &sh.ring_full[BASE + sl], pol);
    }
    __device__ __forceinline__ const double *wait(int sl) {
#ifdef MC_PROFILE
        const long long tw = clock64();
#endif
        mbar_wait(&sh.ring_full[BASE + sl], (phase >> sl) & 1u);
#ifdef MC_PROFILE
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_prof[22 + WHO], (unsigned long long)(clock64() - tw));
#endif
        phase ^= 1u << sl;
        return buf + sl * UNIT_DOUBLES;
    }
    __device__ __forceinline__ void close() { sh.ring_phase[WHO] = phase; }
};
using LtRing = RingT<LT_SLOTS, 0, LT_UNIT_DOUBLES, 0>;
using GtRing = RingT<GT_SLOTS, LT_SLOTS, GT_UNIT_DOUBLES, 1>;

// progress counters between the two warps of a sweep (shared memory, CTA scope)
__device__ __forceinline__ void prog_publish(int *p, int v) { asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
__device__ __forceinline__ int prog_read(const int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
// wait until *p >= need (bounded: a protocol error must not hang the device; it is reported like a failed pivot)
template <bool BACKOFF = false>
__device__ __forceinline__ void prog_wait(IpShared &sh, const int *p, int need) {
    for (int spin = 0; spin < (1 << 18); ++spin) {
        if (prog_read(p) >= need) return;
        if (BACKOFF) __nanosleep(MC_RELAXED_BACKOFF_NS);      // warp 1's waits: it is the one with slack in both sweeps
    }
    sh.flag = 1;
}

// forward sweep (warp 0): y = L^-1 g, one panel of eight columns at a time, as mat-vecs on the tensor cores.  With the
// panel stored as [Q - I; L21] (Q = L11^-1, L21 the 32 rows below) the step is
//   y1 = a1 + (Q - I) a1,   a[k0+8 .. k0+39] -= L21 y1       (rows k0+32 .. k0+39 enter with g).
// The window a lives in DMMA C fragments (every column of the 8 x 8 tile carries the same vector): lane (gq, q) holds
// a[k0 + 8 I + gq], I = 0..3; B operands (a1, y1 at index 4 h + q) come by shuffle.
// z = w y goes to warp 1 (sweep_sep_rhs runs one unit behind) through the queue sw.q, and into the fill rows (GT[k][32])
// for the backward half of the solve.
__device__ __noinline__ void sweep_forward(IpShared &sh, const double *__restrict__ HBp, double *__restrict__ LTp, double *__restrict__ GTp, const int NA, const double *__restrict__ g) {
    const int lane = threadIdx.x & 31;
    const int gq = lane >> 2, q = lane & 3;
    const int nunits = (NA + SUB - 1) / SUB;
    constexpr int AHEAD = LT_SLOTS - 1;
    LtRing R(sh, &sh.u.sw.lt[0][0]);
    if (lane == 0)
        for (int u = 0; u < AHEAD && u < nunits; ++u) R.issue(u, u, LTp);
    double acc[4];
#pragma unroll
    for (int I = 0; I < 4; ++I) acc[I] = (8 * I + gq < NA) ? g[8 * I + gq] : 0.0;
    double gcur = (32 + lane < NA) ? g[32 + lane] : 0.0;       // g of the rows that enter over the next four panels,
    double gnext = (64 + lane < NA) ? g[64 + lane] : 0.0;      // and the block after it (loaded a block ahead)
    int sl = 0, sn = AHEAD % LT_SLOTS;            // slots of unit u and of unit u + AHEAD
    for (int u = 0; u < nunits; ++u, sl = (sl == LT_SLOTS - 1) ? 0 : sl + 1, sn = (sn == LT_SLOTS - 1) ? 0 : sn + 1) {
        const int k0 = u * SUB;
        if (u > 0 && (u & 3) == 0) {
            gcur = gnext;
            const int idx = k0 + 64 + lane;
            gnext = (idx < NA) ? g[idx] : 0.0;
        }
        if ((u & 7) == 0 && u >= 8) prog_wait(sh, &sh.prog[1], u - 8);      // queue slots of units u .. u+7 are free again
        const double *lt = R.wait(sl) + q * LROW + gq;           // A fragments: column 4 h + q of the panel, row (8 I +) gq
        const double ge = __shfl_sync(FULL, gcur, ((u & 3) << 3) + gq);
        double a[5][2];
#pragma unroll
        for (int I = 0; I < 5; ++I) {
            a[I][0] = lt[8 * I];
            a[I][1] = lt[4 * LROW + 8 * I];
        }
        const double wl = lt[(gq - q) * LROW + 41 - gq];        // w of column k0 + gq
        // y1 = a1 + (Q - I) a1: two independent products
        const double b0 = __shfl_sync(FULL, acc[0], 4 * q), b1 = __shfl_sync(FULL, acc[0], 16 + 4 * q);
        double y0[2] = {acc[0], acc[0]}, y1[2] = {0.0, 0.0};
        dmma(y0, a[0][0], b0);
        dmma(y1, a[0][1], b1);
        const double y = y0[0] + y1[0];
        const double n0 = -__shfl_sync(FULL, y, 4 * q), n1 = -__shfl_sync(FULL, y, 16 + 4 * q);
        double nw[4];
#pragma unroll
        for (int I = 0; I < 4; ++I) {                  // block 0 (rows k0+8 ..) first: it is the next panel's a1
            const double c = (I < 3) ? acc[I + 1] : ge;
            double c0[2] = {c, c}, c1[2] = {0.0, 0.0};
            dmma(c0, a[I + 1][0], n0);
            dmma(c1, a[I + 1][1], n1);
            nw[I] = c0[0] + c1[0];
        }
        if (q == 0) {
            const bool in = (k0 + gq < NA);
            const double z = in ? y * wl : 0.0;
            sh.u.sw.q[(u & (QDEPTH - 1)) * SUB + gq] = z;
            if (in) GTp[(size_t)(k0 + gq) * FROW + 32] = z;
        }
#pragma unroll
        for (int I = 0; I < 4; ++I) acc[I] = nw[I];
        __syncwarp();
        if (lane == 0) {
            prog_publish(&sh.prog[0], u + 1);
            if (u + AHEAD < nunits) R.issue(u + AHEAD, sn, LTp);
        }
    }
    R.close();
    fence_proxy_async();        // z (generic-proxy stores) is read back through bulk copies (async proxy)
}

// separator part of the forward sweep (warp 1, one unit behind warp 0):  gs = g_S - G z
__device__ __noinline__ void sweep_sep_rhs(IpShared &sh, const double *__restrict__ HBp, double *__restrict__ LTp, double *__restrict__ GTp, const int NA, const double *__restrict__ g) {
    const int lane = threadIdx.x & 31;
    const int nunits = (NA + SUB - 1) / SUB;
    constexpr int AHEAD = GT_SLOTS - 1;
    GtRing R(sh, &sh.u.sw.gt[0][0]);
    if (lane == 0)
        for (int u = 0; u < AHEAD && u < nunits; ++u) R.issue(u, u, GTp);
    double s0 = 0.0, s1 = 0.0;
    const double gsl = g[NA + lane];
    int sl = 0, sn = AHEAD % GT_SLOTS;
    for (int u = 0; u < nunits; ++u, sl = (sl == GT_SLOTS - 1) ? 0 : sl + 1, sn = (sn == GT_SLOTS - 1) ? 0 : sn + 1) {
        const double *gt = R.wait(sl) + lane;
        prog_wait<true>(sh, &sh.prog[0], u + 1);
        const double *zq = &sh.u.sw.q[(u & (QDEPTH - 1)) * SUB];       // (padding columns: z = 0, g = 0)
#pragma unroll
        for (int s2 = 0; s2 < SUB; s2 += 2) {
            s0 = fma(gt[s2 * FROW], zq[s2], s0);
            s1 = fma(gt[(s2 + 1) * FROW], zq[s2 + 1], s1);
        }
        __syncwarp();
        if (lane == 0) {
            prog_publish(&sh.prog[1], u + 1);
            if (u + AHEAD < nunits) R.issue(u + AHEAD, sn, GTp);
        }
    }
    R.close();
    sh.gs[lane] = gsl - (s0 + s1);
}

// separator solve and the right-hand side of the backward sweep (warp 1, ahead of warp 0's sweep_backward):
//   x_S = S^-1 gs;   t = (y - G^T x_S) w = z - w G^T x_S, from the last unit down, handed over through the queue sw.q
__device__ __noinline__ void sweep_sep_solve(IpShared &sh, const double *__restrict__ HBp, double *__restrict__ LTp, double *__restrict__ GTp, const int NA, double *__restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int nunits = (NA + SUB - 1) / SUB;
    constexpr int AHEAD = GT_SLOTS - 1;
    GtRing R(sh, &sh.u.sw.gt[0][0]);
    const int U0 = nunits - 1;
    if (lane == 0)
        for (int i = 0; i < AHEAD && i < nunits; ++i) R.issue(U0 - i, i, GTp);
    double a = sh.gs[lane];
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        const double yk = __shfl_sync(FULL, a, k);
        const double l = (lane > k) ? sh.s.Ss[lane * 33 + k] : 0.0;
        a = fma(-l, yk, a);
    }
    a *= sh.wS[lane];
#pragma unroll 4
    for (int k = 31; k >= 0; --k) {
        const double xk = __shfl_sync(FULL, a, k);
        const double l = (lane < k) ? sh.s.Ss[k * 33 + lane] : 0.0;
        a = fma(-l, xk, a);
    }
    x[NA + lane] = a;
    // c_k = sum_r G[r][k] x_S[r]: lane = (column k0 + (lane & 7), quarter lane >> 3 of the separator rows)
    const int kl = lane & 7, qr = lane >> 3;
    double xq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xq[i] = __shfl_sync(FULL, a, 8 * qr + i);
    int sl = 0, sn = AHEAD % GT_SLOTS;
    for (int i = 0; i < nunits; ++i, sl = (sl == GT_SLOTS - 1) ? 0 : sl + 1, sn = (sn == GT_SLOTS - 1) ? 0 : sn + 1) {
        const int k = (U0 - i) * SUB + kl;
        if ((i & 7) == 0 && i >= 8) prog_wait<true>(sh, &sh.prog[2], i - 8);      // queue slots of the next eight units are free again
        const double *gt = R.wait(sl) + kl * FROW;
        const double2 zw = *reinterpret_cast<const double2 *>(&gt[32]);
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const double2 gg = *reinterpret_cast<const double2 *>(&gt[8 * qr + e]);
            c0 = fma(gg.x, xq[e], c0);
            c1 = fma(gg.y, xq[e + 1], c1);
        }
        double c = c0 + c1;
        c += __shfl_xor_sync(FULL, c, 8);
        c += __shfl_xor_sync(FULL, c, 16);
        if (qr == 0) sh.u.sw.q[(i & (QDEPTH - 1)) * SUB + kl] = (k < NA) ? fma(-c, zw.y, zw.x) : 0.0;      // (padding rows: t = 0)
        __syncwarp();
        if (lane == 0) {
            prog_publish(&sh.prog[3], i + 1);
            if (i + AHEAD < nunits) R.issue(U0 - i - AHEAD, sn, GTp);
        }
    }
    R.close();
}

// backward sweep (warp 0): x = L^-T t, one panel at a time from the last one, as mat-vecs on the tensor cores:
//   u = t1 - L21^T x[k0+8 .. k0+39],   x[k0 .. k0+7] = u + (Q - I)^T u.
// The 32 x's below the panel are kept NEGATED as B operands (lane (gq, q): -x[k0 + 8 + 4 i + q], i = 0..7); u and the eight new
// x's come out in C layout (lane (gq, .): row k0 + gq) and move over by shuffle.  Only the two products with the previous
// panel's x and the two with u are on the chain from panel to panel; the other six are issued ahead of them.
// t comes from warp 1 (sweep_sep_solve, running ahead) through the queue sw.q.
__device__ __noinline__ void sweep_backward(IpShared &sh, const double *__restrict__ HBp, double *__restrict__ LTp, double *__restrict__ GTp, const int NA, double *__restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int gq = lane >> 2, q = lane & 3;
    const int nunits = (NA + SUB - 1) / SUB;
    constexpr int AHEAD = LT_SLOTS - 1;
    LtRing R(sh, &sh.u.sw.lt[0][0]);
    const int U0 = nunits - 1;
    if (lane == 0)
        for (int i = 0; i < AHEAD && i < nunits; ++i) R.issue(U0 - i, i, LTp);
    double nbx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) nbx[i] = 0.0;
    int sl = 0, sn = AHEAD % LT_SLOTS;
    for (int i = 0; i < nunits; ++i, sl = (sl == LT_SLOTS - 1) ? 0 : sl + 1, sn = (sn == LT_SLOTS - 1) ? 0 : sn + 1) {
        const int k0 = (U0 - i) * SUB;
        const double *lt = R.wait(sl) + gq * LROW + q;           // A fragments: column gq of the panel, row 4 c + q
        double a[10];
#pragma unroll
        for (int c = 0; c < 10; ++c) a[c] = lt[4 * c];
        double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll
        for (int e = 2; e < 8; e += 2) {
            dmma(c0, a[2 + e], nbx[e]);
            dmma(c1, a[3 + e], nbx[e + 1]);
        }
        dmma(c0, a[2], nbx[0]);
        dmma(c1, a[3], nbx[1]);
        prog_wait(sh, &sh.prog[3], i + 1);
        const double t1 = sh.u.sw.q[(i & (QDEPTH - 1)) * SUB + gq];
        const double uu = t1 + (c0[0] + c1[0]);
        const double u0 = __shfl_sync(FULL, uu, 4 * q), u1 = __shfl_sync(FULL, uu, 16 + 4 * q);
        double d0[2] = {uu, uu}, d1[2] = {0.0, 0.0};
        dmma(d0, a[0], u0);
        dmma(d1, a[1], u1);
        const double x1 = d0[0] + d1[0];
        if (q == 0 && k0 + gq < NA) x[k0 + gq] = x1;
#pragma unroll
        for (int e = 7; e >= 2; --e) nbx[e] = nbx[e - 2];
        nbx[0] = -__shfl_sync(FULL, x1, 4 * q);
        nbx[1] = -__shfl_sync(FULL, x1, 16 + 4 * q);
        __syncwarp();
        if (lane == 0) {
            prog_publish(&sh.prog[2], i + 1);
            if (i + AHEAD < nunits) R.issue(U0 - i - AHEAD, sn, LTp);
        }
    }
    R.close();
}

// x = M^-1 g with the stored factor.  fused: the forward part was done inside factor() (predictor).  Both halves run on the
// two warps concurrently: warp 1's separator reduction trails warp 0's forward sweep, then (after the separator solve)
// warp 1's right-hand side t runs ahead of warp 0's backward sweep.
__device__ __noinline__ void solve(IpShared &sh, double *slab, const Layout &L, int n, const double *g, double *x, bool fused) {
    const Factor F = make_factor(slab, L, n);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 4) sh.prog[threadIdx.x] = 0;
    fence_proxy_async();        // the ring area was last accessed through the generic proxy (factor hand-off buffers)
    __syncthreads();
    if (!fused) {
        PROF_T0(t0);
        if (warp == 0) sweep_forward(sh, F.HB, F.LT, F.GT, F.NA, g);
        else sweep_sep_rhs(sh, F.HB, F.LT, F.GT, F.NA, g);
        __syncthreads();
        PROF_ADD1(2, t0);
    }
    PROF_T0(t3);
    if (warp == 1) sweep_sep_solve(sh, F.HB, F.LT, F.GT, F.NA, x);
    else sweep_backward(sh, F.HB, F.LT, F.GT, F.NA, x);
    __syncthreads();
    PROF_ADD1(8, t3);
}

// banded cyclic mat-vec out = H v (real-indexed)
__device__ void band_matvec(const double *__restrict__ HB, const double *__restrict__ v, double *__restrict__ out, int n) {
    for (int i = threadIdx.x; i < n; i += IP_THREADS) {
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

__device__ __forceinline__ void ip_init_shared(IpShared &sh) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < HB_SLOTS; ++i) mbar_init(&sh.hb_full[i], 1);
        for (int i = 0; i < HO_SLOTS; ++i) { mbar_init(&sh.ho_full[i], 1); mbar_init(&sh.ho_empty[i], 1); }
        for (int i = 0; i < RING_UNITS; ++i) mbar_init(&sh.ring_full[i], 1);
        sh.ring_phase[0] = 0; sh.ring_phase[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

__global__ void __launch_bounds__(IP_THREADS, 8)
mincurv_pdip_kernel(int B, int n_max, const int32_t *__restrict__ n_pts, double *__restrict__ ws, Layout L,
                    PdipParams prm, double *__restrict__ alpha_out, int32_t *__restrict__ status,
                    int32_t *__restrict__ iters_out, int *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IpShared &sh = *reinterpret_cast<IpShared *>(smem_raw);
#ifdef MC_DEBUG_SM_LIMIT       // contention experiments: only the first MC_DEBUG_SM_LIMIT SMs take work (full occupancy on those)
    {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (smid >= MC_DEBUG_SM_LIMIT) return;
    }
#endif
    ip_init_shared(sh);
    unsigned tick = 0;      // hand-off units of the factorisations so far (uniform across the CTA)

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
            for (int i = threadIdx.x; i < n_max; i += IP_THREADS) aout[i] = 0.0;
            if (iters_out && threadIdx.x == 0) iters_out[b] = 0;
            continue;
        }
        PROF_T0(tq0);
        double *slab = ws + (size_t)b * L.stride;
        const double *HB = slab + L.o_hb;
        const double *__restrict__ LB = vec(slab, L, V_LB), *__restrict__ UB = vec(slab, L, V_UB), *__restrict__ F = vec(slab, L, V_F);
        double *__restrict__ AL = vec(slab, L, V_ALPHA), *__restrict__ LU = vec(slab, L, V_LU), *__restrict__ LL = vec(slab, L, V_LL), *__restrict__ RD = vec(slab, L, V_RD);
        double *__restrict__ RHS = vec(slab, L, V_RHS), *__restrict__ DX = vec(slab, L, V_DX), *__restrict__ DD = vec(slab, L, V_DD);
        double *__restrict__ TU = vec(slab, L, V_DLU), *__restrict__ TL = vec(slab, L, V_DLL), *__restrict__ SU = vec(slab, L, V_SU), *__restrict__ SL = vec(slab, L, V_SL);
        double *__restrict__ ISU = vec(slab, L, V_ISU), *__restrict__ ISL = vec(slab, L, V_ISL);
        double *G0 = vec(slab, L, V_T0);
        if (threadIdx.x == 0) sh.flag = 0;

        // ---------------- initial point: box centre, multipliers from the gradient ----------------
        for (int i = threadIdx.x; i < n; i += IP_THREADS) AL[i] = 0.5 * (LB[i] + UB[i]);
        __syncthreads();
        band_matvec(HB, AL, G0, n);
        __syncthreads();
        double gmax = 0.0, fmaxv = 0.0;
#pragma unroll 1
        for (int i = threadIdx.x; i < n; i += IP_THREADS) {
            const double gi = G0[i] + F[i];
            G0[i] = gi;
            gmax = fmax(gmax, fabs(gi));
            fmaxv = fmax(fmaxv, fabs(F[i]));
        }
        gmax = block_reduce<1>(gmax, sh.red);
        fmaxv = block_reduce<1>(fmaxv, sh.red);
        const double lam0 = prm.lam0_rel * gmax + 1e-300;
        double musum = 0.0;
#pragma unroll 1
        for (int i = threadIdx.x; i < n; i += IP_THREADS) {
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
        // Every thread owns the elements i = tid + k * 64; the loops take them in groups of VG with all loads of a
        // group issued before the first use (the vectors live in L2/HBM: one element at a time, each trip of a loop
        // paid the full memory latency).
        constexpr int VG = 4, VS = VG * IP_THREADS;
        for (it = 0; it < prm.max_iter; ++it) {
            // ---- barrier diagonal and affine right-hand side ----
            PROF_T0(tv1);
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS) {
                double lu[VG], ll[VG], isu[VG], isl[VG], rd[VG];
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = min(i0 + k * IP_THREADS, n - 1);
                    lu[k] = LU[i]; ll[k] = LL[i]; isu[k] = ISU[i]; isl[k] = ISL[i]; rd[k] = RD[i];
                }
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = i0 + k * IP_THREADS;
                    if (i < n) { DD[i] = lu[k] * isu[k] + ll[k] * isl[k]; RHS[i] = -rd[k] + lu[k] - ll[k]; }
                }
            }
            __syncthreads();
            PROF_ADD(16, tv1);
            PROF_T0(tf0);
            const bool fok = factor(sh, slab, L, n, RHS, tick);
            tick += factor_units(n);
            if (!fok) { result = 3; break; }
            PROF_ADD(10, tf0);
            PROF_T0(ts0);
            solve(sh, slab, L, n, RHS, DX, true);
            PROF_ADD(6, ts0);
            // ---- affine direction: step lengths 1 / max-ratio; mu_aff as a polynomial in (ap, ad) ----
            PROF_T0(tv2);
            double rp = 0.0, rdl = 0.0, c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS) {
                double su[VG], sl[VG], lu[VG], ll[VG], dx[VG], isu[VG], isl[VG];
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = min(i0 + k * IP_THREADS, n - 1);
                    su[k] = SU[i]; sl[k] = SL[i]; lu[k] = LU[i]; ll[k] = LL[i]; dx[k] = DX[i]; isu[k] = ISU[i]; isl[k] = ISL[i];
                }
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    if (i0 + k * IP_THREADS < n) {
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
            PROF_T0(tv3);
            // ---- corrector right-hand side ----
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS) {
                double su[VG], sl[VG], lu[VG], ll[VG], dx[VG], isu[VG], isl[VG], rd[VG];
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = min(i0 + k * IP_THREADS, n - 1);
                    su[k] = SU[i]; sl[k] = SL[i]; lu[k] = LU[i]; ll[k] = LL[i]; dx[k] = DX[i]; isu[k] = ISU[i]; isl[k] = ISL[i]; rd[k] = RD[i];
                }
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = i0 + k * IP_THREADS;
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
            PROF_T0(ts1);
            solve(sh, slab, L, n, RHS, DX, false);
            PROF_ADD(7, ts1);
            PROF_T0(tv4);
            rp = 0.0; rdl = 0.0;
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS) {
                double lu[VG], ll[VG], dx[VG], isu[VG], isl[VG], tu[VG], tl[VG];
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    const int i = min(i0 + k * IP_THREADS, n - 1);
                    lu[k] = LU[i]; ll[k] = LL[i]; dx[k] = DX[i]; isu[k] = ISU[i]; isl[k] = ISL[i]; tu[k] = TU[i]; tl[k] = TL[i];
                }
#pragma unroll
                for (int k = 0; k < VG; ++k) {
                    if (i0 + k * IP_THREADS < n) {
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
            PROF_T0(tv5);
            double musum2 = 0.0, rdmax = 0.0, dxmax = 0.0, amax = 0.0;
            constexpr int VG2 = 2, VS2 = VG2 * IP_THREADS;        // 13 input vectors: groups of two
#pragma unroll 1
            for (int i0 = threadIdx.x; i0 < n; i0 += VS2) {
                double su[VG2], sl[VG2], lu[VG2], ll[VG2], dx[VG2], isu[VG2], isl[VG2], tu[VG2], tl[VG2], al[VG2], rd[VG2], rh[VG2], dd[VG2];
#pragma unroll
                for (int k = 0; k < VG2; ++k) {
                    const int i = min(i0 + k * IP_THREADS, n - 1);
                    su[k] = SU[i]; sl[k] = SL[i]; lu[k] = LU[i]; ll[k] = LL[i]; dx[k] = DX[i]; isu[k] = ISU[i]; isl[k] = ISL[i];
                    tu[k] = TU[i]; tl[k] = TL[i]; al[k] = AL[i]; rd[k] = RD[i]; rh[k] = RHS[i]; dd[k] = DD[i];
                }
#pragma unroll
                for (int k = 0; k < VG2; ++k) {
                    const int i = i0 + k * IP_THREADS;
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
        for (int i = threadIdx.x; i < n_max; i += IP_THREADS) aout[i] = (i < n) ? AL[i] : 0.0;
        if (threadIdx.x == 0) {
            status[b] = result;
            if (iters_out) iters_out[b] = it;
        }
        PROF_ADD(9, tq0);
#ifdef MC_PROFILE
        if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&g_prof[11], 1ull); atomicAdd(&g_prof[12], (unsigned long long)it); }
#endif
    }
}

size_t pdip_smem_bytes() { return sizeof(IpShared); }

int debug_read_profile(unsigned long long *host_out, int reset) {
#ifdef MC_PROFILE
    if (cudaMemcpyFromSymbol(host_out, g_prof, sizeof(unsigned long long) * 24) != cudaSuccess) return -1;
    if (const char *e = getenv("MC_PROFILE_SEGMENTS")) {       // tools/prof_run.py: sub-phase counters of one panel
        (void)e;
        unsigned long long seg[32];
        if (cudaMemcpyFromSymbol(seg, g_seg, sizeof(seg)) == cudaSuccess) {
            fprintf(stderr, "segments:");
            for (int i = 0; i < 32; ++i) fprintf(stderr, " %d:%llu", i, seg[i]);
            fprintf(stderr, "\n");
        }
    }
    if (reset) {
        unsigned long long z[32] = {0};
        if (cudaMemcpyToSymbol(g_prof, z, sizeof(unsigned long long) * 24) != cudaSuccess) return -1;
        cudaMemcpyToSymbol(g_seg, z, sizeof(z));
    }
    return 0;
#else
    (void)reset;
    for (int i = 0; i < 24; ++i) host_out[i] = 0ull;
    return 0;
#endif
}

// resident CTAs per SM of a solver kernel on the current device (registers and shared memory both count)
template <typename K>
static int ctas_per_sm(K kernel) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, IP_THREADS, sizeof(IpShared)) != cudaSuccess) return 1;
    return nb > 0 ? nb : 1;
}

int pdip_ctas_per_sm() { return ctas_per_sm(mincurv_pdip_kernel); }

int launch_mincurv_pdip(int B, int n_max, const int32_t *n_pts, double *ws, const Layout &L, const PdipParams &prm,
                        double *alpha, int32_t *status, int32_t *iters, int grid, int *work_counter, cudaStream_t stream) {
    // (per launch: the attribute is per device and a process may drive several)
    cudaError_t e = cudaFuncSetAttribute(mincurv_pdip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IpShared));
    if (e != cudaSuccess) return (int)e;
    cudaMemsetAsync(work_counter, 0, sizeof(int), stream);
    mincurv_pdip_kernel<<<grid, IP_THREADS, sizeof(IpShared), stream>>>(B, n_max, n_pts, ws, L, prm, alpha, status, iters, work_counter);
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
__global__ void __launch_bounds__(IP_THREADS, 8)
mincurv_pdip_kappa_kernel(int B, int n_max, const int32_t *__restrict__ n_pts, double *__restrict__ ws, Layout L,
                          PdipParams prm, double kb, double *__restrict__ alpha_out, int32_t *__restrict__ status,
                          int32_t *__restrict__ iters_out, int *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IpShared &sh = *reinterpret_cast<IpShared *>(smem_raw);
    ip_init_shared(sh);
    unsigned tick = 0;
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
        const double *__restrict__ LB = vec(slab, L, V_LB), *__restrict__ UB = vec(slab, L, V_UB), *__restrict__ F = vec(slab, L, V_F);
        const double *__restrict__ KR = vec(slab, L, V_KREF);
        double *__restrict__ AL = vec(slab, L, V_ALPHA), *__restrict__ LU = vec(slab, L, V_LU), *__restrict__ LL = vec(slab, L, V_LL);
        double *__restrict__ RHS = vec(slab, L, V_RHS), *__restrict__ DX = vec(slab, L, V_DX), *__restrict__ DD = vec(slab, L, V_DD);
        double *__restrict__ TU = vec(slab, L, V_DLU), *__restrict__ TL = vec(slab, L, V_DLL), *__restrict__ SU = vec(slab, L, V_SU), *__restrict__ SL = vec(slab, L, V_SL);
        double *__restrict__ ISU = vec(slab, L, V_ISU), *__restrict__ ISL = vec(slab, L, V_ISL);
        double *__restrict__ S3 = vec(slab, L, V_S3), *__restrict__ S4 = vec(slab, L, V_S4), *__restrict__ L3 = vec(slab, L, V_L3), *__restrict__ L4 = vec(slab, L, V_L4);
        double *__restrict__ KL = vec(slab, L, V_KL), *__restrict__ WK = vec(slab, L, V_WK), *__restrict__ EDX = vec(slab, L, V_EDX);
        double *__restrict__ T3 = vec(slab, L, V_T3K), *__restrict__ T4 = vec(slab, L, V_T4K), *__restrict__ VV = vec(slab, L, V_VV), *__restrict__ ETV = vec(slab, L, V_RD);
        double *t0 = vec(slab, L, V_T0), *t1 = vec(slab, L, V_T1), *t2 = vec(slab, L, V_T2), *t3 = vec(slab, L, V_T3), *t4 = vec(slab, L, V_T4), *t5 = vec(slab, L, V_T5);
        const double m4 = 4.0 * n;
        if (threadIdx.x == 0) sh.flag = 0;

        // ---- start: box centre; kl = k_ref + E a; gradient g = E^T (E a) + f ----
        for (int i = threadIdx.x; i < n; i += IP_THREADS) AL[i] = 0.5 * (LB[i] + UB[i]);
        __syncthreads();
        apply_E(slab, L, n, AL, EDX, t0, t1, t2, t3, t4, t5);
        apply_Et(slab, L, n, EDX, ETV, t0, t1, t2, t3, t4, t5);
        double gmax = 0.0, fmaxv = 0.0;
        for (int i = threadIdx.x; i < n; i += IP_THREADS) {
            gmax = fmax(gmax, fabs(ETV[i] + F[i]));
            fmaxv = fmax(fmaxv, fabs(F[i]));
        }
        gmax = block_reduce<1>(gmax, sh.red);
        fmaxv = block_reduce<1>(fmaxv, sh.red);
        const double lam0 = prm.lam0_rel * gmax + 1e-300;
        double musum = 0.0;
        for (int i = threadIdx.x; i < n; i += IP_THREADS) {
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
            for (int i = threadIdx.x; i < n; i += IP_THREADS) {
                const double s3 = S3[i], s4 = S4[i], l3 = L3[i], l4 = L4[i], kl = KL[i];
                const double rp3 = kl + s3 - kb, rp4 = -kl + s4 - kb;
                WK[i] = l3 / s3 + l4 / s4;
                DD[i] = LU[i] * ISU[i] + LL[i] * ISL[i];
                VV[i] = (kl - KR[i]) + l3 * rp3 / s3 - l4 * rp4 / s4;       // affine: t3 = -s3 l3, t4 = -s4 l4
            }
            __syncthreads();
            assemble_hband(slab, L, n, WK, sh.u.win);     // the tile area is free between the sweeps and the next factorisation
            __syncthreads();
            apply_Et(slab, L, n, VV, ETV, t0, t1, t2, t3, t4, t5);
            for (int i = threadIdx.x; i < n; i += IP_THREADS) RHS[i] = -F[i] - ETV[i];
            __syncthreads();
            const bool fok = factor(sh, slab, L, n, RHS, tick);
            tick += factor_units(n);
            if (!fok) {
                // E^T W E with W = l/s -> 1e12 and beyond is no longer numerically SPD: accept a late iterate, else give up
                result = (mu <= 1e-7 * mu0 && rdmax <= 1e3 * rd_tol && rpmax <= 1e-6 * kb) ? 0 : 3;
                break;
            }
            solve(sh, slab, L, n, RHS, DX, true);
            apply_E(slab, L, n, DX, EDX, t0, t1, t2, t3, t4, t5);
            // ---- affine step lengths and centring ----
            double rp = 0.0, rdl = 0.0, c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
            for (int i = threadIdx.x; i < n; i += IP_THREADS) {
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
            for (int i = threadIdx.x; i < n; i += IP_THREADS) {
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
            for (int i = threadIdx.x; i < n; i += IP_THREADS) RHS[i] -= ETV[i];
            __syncthreads();
            solve(sh, slab, L, n, RHS, DX, false);
            apply_E(slab, L, n, DX, EDX, t0, t1, t2, t3, t4, t5);
            rp = 0.0; rdl = 0.0;
            for (int i = threadIdx.x; i < n; i += IP_THREADS) {
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
            for (int i = threadIdx.x; i < n; i += IP_THREADS) {
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
            for (int i = threadIdx.x; i < n; i += IP_THREADS) rdmax = fmax(rdmax, fabs(ETV[i] + F[i] + LU[i] - LL[i]));
            rdmax = block_reduce<1>(rdmax, sh.red);
            if (mu <= prm.mu_rel * mu0 && rdmax <= rd_tol && rpmax <= 1e-8 * kb) { result = 0; ++it; break; }
            if (mu <= 1e-2 * prm.mu_rel * mu0) { result = (rdmax <= 1e3 * rd_tol && rpmax <= 1e-6 * kb) ? 0 : 2; ++it; break; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_max; i += IP_THREADS) aout[i] = (i < n) ? AL[i] : 0.0;
        if (threadIdx.x == 0) {
            status[b] = result;
            if (iters_out) iters_out[b] += it;
        }
    }
}

int launch_mincurv_pdip_kappa(int B, int n_max, const int32_t *n_pts, double *ws, const Layout &L, const PdipParams &prm,
                              double kappa_bound, double *alpha, int32_t *status, int32_t *iters, int grid, int *work_counter,
                              cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(mincurv_pdip_kappa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IpShared));
    if (e != cudaSuccess) return (int)e;
    cudaMemsetAsync(work_counter, 0, sizeof(int), stream);
    mincurv_pdip_kappa_kernel<<<grid, IP_THREADS, sizeof(IpShared), stream>>>(B, n_max, n_pts, ws, L, prm, kappa_bound, alpha,
                                                                             status, iters, work_counter);
    return 0;
}

int pdip_kappa_ctas_per_sm() { return ctas_per_sm(mincurv_pdip_kappa_kernel); }

// ================================================================================================
// debug aid (tests/test_gpu_factor.py): factorise M = H + D of every instance with the fused forward substitution of
// V_RHS, solve into V_DX; then solve M x = V_T0 with the full sweeps into V_T1.  HB, V_DD, V_RHS, V_T0 are inputs.
// ================================================================================================
__global__ void __launch_bounds__(IP_THREADS, 8)
debug_factor_solve_kernel(int B, int n_max, const int32_t *__restrict__ n_pts, double *__restrict__ ws, Layout L,
                          int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IpShared &sh = *reinterpret_cast<IpShared *>(smem_raw);
    ip_init_shared(sh);
    unsigned tick = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const int n = n_pts ? n_pts[b] : n_max;
        double *slab = ws + (size_t)b * L.stride;
        if (threadIdx.x == 0) sh.flag = 0;
        __syncthreads();
        const bool ok = factor(sh, slab, L, n, vec(slab, L, V_RHS), tick);
        tick += factor_units(n);
        solve(sh, slab, L, n, vec(slab, L, V_RHS), vec(slab, L, V_DX), true);
        solve(sh, slab, L, n, vec(slab, L, V_T0), vec(slab, L, V_T1), false);
        // a second factorisation exercises the parity bookkeeping of the rings across calls
        const bool ok2 = factor(sh, slab, L, n, vec(slab, L, V_T0), tick);
        tick += factor_units(n);
        solve(sh, slab, L, n, vec(slab, L, V_T0), vec(slab, L, V_T2), true);
        if (threadIdx.x == 0) status[b] = (ok && ok2) ? 0 : 3;
        __syncthreads();
    }
}

int launch_debug_factor_solve(int B, int n_max, const int32_t *n_pts, double *ws, const Layout &L, int32_t *status, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(debug_factor_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IpShared));
    if (e != cudaSuccess) return (int)e;
    const int grid = B < 1184 ? B : 1184;
    debug_factor_solve_kernel<<<grid, IP_THREADS, sizeof(IpShared), stream>>>(B, n_max, n_pts, ws, L, status);
    return 0;
}

}  // namespace mc
