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
// The factor (L: 32 doubles per column, G: 32 + 1 pad, plus the vectors 1/d and y) goes to the instance's HBM slab and
// is streamed back by the triangular sweeps through a shared-memory ring of 8-column units filled by cp.async.bulk
// (TMA, 1-D) on mbarriers -- the band rows of H reach warp 0 the same way.  Per iteration: factor written once, read
// three times (the predictor's forward sweep is fused into the factorisation).
#include "mincurv_ops.cuh"

namespace mc {

constexpr unsigned FULL = 0xffffffffu;
constexpr int IP_THREADS = 64;
constexpr int SUB = 8;                 // columns per hand-off / streaming unit
constexpr int LTS = 34;                // shared pitch of a factor column staged for the fill warp: [l(32), w, y]
constexpr int GTS = 33;                // pitch of a fill column [g(32), pad] (HBM and shared: odd => conflict-free row-owner reads)
constexpr int LTG = 32;                // HBM pitch of a factor column
constexpr int HB_SLOTS = 3;            // band-row units in flight for warp 0
constexpr int LT_SLOTS = 3;            // factor-column units between warp 0 and warp 1
constexpr int RING_UNITS = 7;          // sweep ring: 32-row window (5 units) + 2 units of prefetch
constexpr int RING_UNIT_DOUBLES = SUB * GTS;          // 264 (a unit of L is 256)
constexpr unsigned HB_UNIT_BYTES = SUB * HB_PITCH * sizeof(double);   // 2176
constexpr unsigned LT_UNIT_BYTES = SUB * LTG * sizeof(double);        // 2048
constexpr unsigned GT_UNIT_BYTES = SUB * GTS * sizeof(double);        // 2112

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
// try_wait suspends the thread for a hardware-defined time before it reports failure: the loop is not a busy spin
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
                 " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// The factor is a stream (written once, read three times per iteration, 0.5 MB per instance, far beyond what L2 can
// keep for 1184 resident instances): its copies and stores carry an evict-first L2 policy so that they do not push the
// O(N) iterate vectors of the interior-point loop out of L2.
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
#else
#define PROF_T0(name) do { } while (0)
#define PROF_ADD(slot, t0) do { } while (0)
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
__device__ __forceinline__ double mask_hi(double x, int m) {
    return __hiloint2double(__double2hiint(x) & m, __double2loint(x));
}

struct IpShared {
    union {
        struct {
            double hb[HB_SLOTS][SUB * HB_PITCH];      // band rows of H staged for warp 0 (TMA)
            double lt[LT_SLOTS][SUB * LTS];           // factor columns, warp 0 -> warp 1
            double gt[SUB * GTS];                     // fill columns of the current unit (warp 1: S update)
        } f;
        double ring[RING_UNITS][RING_UNIT_DOUBLES];   // sweeps: streamed factor units
        double win[hband_win_doubles(IP_THREADS)];    // K2b': scratch of the weighted band assembly
    } u;
    union {
        double Ss[32 * 33];                           // separator block -> its L_S (strictly lower, in place)
        double sfrag[20 * 32];                        // during the chain: S accumulators (DMMA C fragments, lane-major)
    } s;
    double cb[2][36];                                 // warp 0: column broadcast [v(32), pivot, y, -, -], double buffered
    double wS[32], xs[32], gs[32], part[2][32];
    double red[32];
    uint64_t hb_full[HB_SLOTS], lt_full[LT_SLOTS], lt_empty[LT_SLOTS], ring_full[RING_UNITS];
    unsigned ring_phase;                              // parity bit per ring slot
    int flag;
    int next;                                         // next instance index (dynamic work distribution)
};

// pointers into the instance slab that the factorisation and the sweeps use
struct Factor {
    const double *HB;      // band of H, row i: H[i][i .. i+32], [33]: pivot H_ii + D_i (written by factor())
    const double *DD;      // barrier diagonal
    double *LT;            // [NA][32]  unit-lower factor columns: LT[k][rho] = L[k+1+rho][k]
    double *GT;            // [NA][33]  fill columns: GT[k][r] = G[r][k]
    double *WP, *YP, *TP;  // [NA] 1/d_k, forward-substituted rhs, backward rhs (y - G^T x_S) / d
    int n, NA;
};

__device__ __forceinline__ Factor make_factor(double *slab, const Layout &L, int n) {
    Factor F;
    F.HB = slab + L.o_hb;
    F.DD = vec(slab, L, V_DD);
    F.LT = slab + L.o_tiles;
    F.GT = F.LT + (size_t)L.np * LTG;
    F.WP = vec(slab, L, V_WP);
    F.YP = vec(slab, L, V_YPAD);
    F.TP = vec(slab, L, V_TP);
    F.n = n;
    F.NA = n - 32;
    return F;
}

// ---- warp 0: LDL^T of the chain, column by column, with the forward substitution of g fused in ----
__device__ __noinline__ bool factor_chain(IpShared &sh, const Factor F, const double *__restrict__ g, unsigned tick) {
    const int lane = threadIdx.x & 31;
    const int NA = F.NA;
    const int nunits = (NA + SUB - 1) / SUB;
    const uint64_t pol = l2_evict_first_policy();
    double acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.0;
    double gacc = 0.0;
    double gk = (lane < NA) ? g[lane] : 0.0;          // rhs entry of the row this lane finishes next
    bool ok = true;
    if (lane == 0) {
        for (int t = 0; t < 2 && t < nunits; ++t) {
            const unsigned hs = (tick + t) % HB_SLOTS;
            mbar_expect_tx(&sh.hb_full[hs], HB_UNIT_BYTES);
            tma_load_1d(sh.u.f.hb[hs], F.HB + (size_t)t * SUB * HB_PITCH, HB_UNIT_BYTES, &sh.hb_full[hs], pol);
        }
    }
    for (int t = 0; t < nunits; ++t) {
        const unsigned ht = tick + t, hs = ht % HB_SLOTS, ls = ht % LT_SLOTS;
        if (lane == 0 && t + 2 < nunits) {            // slot of unit t-1: every lane has passed that unit's last __syncwarp
            const unsigned h2 = (ht + 2) % HB_SLOTS;
            mbar_expect_tx(&sh.hb_full[h2], HB_UNIT_BYTES);
            tma_load_1d(sh.u.f.hb[h2], F.HB + (size_t)(t + 2) * SUB * HB_PITCH, HB_UNIT_BYTES, &sh.hb_full[h2], pol);
        }
        mbar_wait(&sh.hb_full[hs], (ht / HB_SLOTS) & 1u);
        if (ht >= (unsigned)LT_SLOTS) mbar_wait(&sh.lt_empty[ls], ((ht / LT_SLOTS) - 1u) & 1u);
        const double *hbg = sh.u.f.hb[hs];
        double *ltg = sh.u.f.lt[ls];
        const int k0 = t * SUB;
        const int nst = min(SUB, NA - k0);
        int rho = (lane - k0 - 1) & 31;               // this lane's row is k + 1 + rho; rho == 31: the row k + 32 enters
        double a_cur = hbg[1 + rho], pd_cur = hbg[HB_PITCH - 1];
#pragma unroll 1
        for (int s = 0; s < nst; ++s) {
            const int k = k0 + s;
            const bool isk = (rho == 31);             // lane == k & 31: holds the pivot row k and receives row k + 32
            const double a = (k + 1 + rho < NA) ? a_cur : 0.0;
            const double t0 = (isk ? pd_cur : a) - acc[0];
            const double v = isk ? a : t0;            // (the entering row has no updates yet)
            double *cb = sh.cb[k & 1];
            cb[rho] = v;
            if (isk) *reinterpret_cast<double2 *>(&cb[32]) = make_double2(t0, gk - gacc);
            __syncwarp();
            if (s + 1 < nst) {                        // band entries of the next column (off the pivot chain)
                a_cur = hbg[(s + 1) * HB_PITCH + 1 + ((rho - 1) & 31)];
                pd_cur = hbg[(s + 1) * HB_PITCH + HB_PITCH - 1];
            }
            const double2 dy = *reinterpret_cast<const double2 *>(&cb[32]);
            if (!(dy.x > 0.0)) ok = false;
            const double w = fast_rcp(dy.x);
            const double lm = v * w;
            const int m = isk ? 0 : -1;
            // the window slides by one column: acc[c] <- acc[c+1] + l v_c   (slot c: column k + 1 + c)
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                const double2 cc = *reinterpret_cast<const double2 *>(&cb[c]);
                acc[c] = fma(lm, cc.x, mask_hi(acc[c + 1], m));
                if (c + 2 < 32) acc[c + 1] = fma(lm, cc.y, mask_hi(acc[c + 2], m));
                else acc[c + 1] = lm * cc.y;          // column k + 32 enters the window
            }
            gacc = fma(lm, dy.y, isk ? 0.0 : gacc);
            // the column: to the fill warp (shared) and to the slab (HBM)
            ltg[s * LTS + rho] = lm;
            __stcs(&F.LT[(size_t)k * LTG + rho], lm);
            if (isk) {
                *reinterpret_cast<double2 *>(&ltg[s * LTS + 32]) = make_double2(w, dy.y);
                F.WP[k] = w;
                F.YP[k] = dy.y;
                gk = (k + 32 < NA) ? g[k + 32] : 0.0;
            }
            rho = (rho - 1) & 31;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.lt_full[ls]);
    }
    return ok;
}

// ---- warp 1: fill row G = Y L^-T (lane = separator row), S -= (G w) G^T, and the separator part of the fused forward
//      substitution  gS -= G (w y) ----
__device__ __noinline__ void factor_fill(IpShared &sh, const Factor F, const double *__restrict__ g, unsigned tick) {
    const int lane = threadIdx.x & 31;
    const int gq = lane >> 2, q = lane & 3;
    const int NA = F.NA;
    const int nunits = (NA + SUB - 1) / SUB;
    double facc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) facc[c] = 0.0;
    double gsacc = 0.0;
    for (int t = 0; t < nunits; ++t) {
        const unsigned ht = tick + t, ls = ht % LT_SLOTS;
        mbar_wait(&sh.lt_full[ls], (ht / LT_SLOTS) & 1u);
        const double *ltg = sh.u.f.lt[ls];
        const int k0 = t * SUB;
        const int nst = min(SUB, NA - k0);
        const bool hasY = (k0 < 32) || (k0 + SUB - 1 >= NA - 32);     // Y = M[sep, chain] is nonzero across the wrap and next to the separator
#pragma unroll 1
        for (int s = 0; s < nst; ++s) {
            const int k = k0 + s;
            const double *col = ltg + s * LTS;
            double yv = 0.0;
            if (hasY) {
                if (k <= lane) yv = F.HB[(size_t)(NA + lane) * HB_PITCH + (k + 32 - lane)];
                else if (k >= NA + lane - 32) yv = F.HB[(size_t)k * HB_PITCH + (NA + lane - k)];
            }
            const double gv = yv - facc[0];
            sh.u.f.gt[s * GTS + lane] = gv;
            __stcs(&F.GT[(size_t)k * GTS + lane], gv);
            const double2 wy = *reinterpret_cast<const double2 *>(&col[32]);
            gsacc = fma(gv, wy.x * wy.y, gsacc);
            // the window slides by one column: facc[c] <- facc[c+1] + g l_c   (slot c: column k + 1 + c)
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                const double2 cc = *reinterpret_cast<const double2 *>(&col[c]);
                facc[c] = fma(gv, cc.x, facc[c + 1]);
                if (c + 2 < 32) facc[c + 1] = fma(gv, cc.y, facc[c + 2]);
                else facc[c + 1] = gv * cc.y;         // column k + 32 enters the window
            }
        }
        if (nst < SUB) {
            for (int s = nst; s < SUB; ++s) sh.u.f.gt[s * GTS + lane] = 0.0;
        }
        __syncwarp();
        // ---- S -= (G w) G^T over the unit's columns: lower 8x8 blocks (I >= J) on the tensor cores ----
#pragma unroll 1
        for (int I = 0; I < 4; ++I) {
            double a[2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int kk = 4 * ks + q;
                const double wq = (kk < nst) ? ltg[kk * LTS + 32] : 0.0;
                a[ks] = sh.u.f.gt[kk * GTS + 8 * I + gq] * wq;
            }
#pragma unroll 1
            for (int J = 0; J <= I; ++J) {
                const int bi = (I * (I + 1)) / 2 + J;
                double c2[2];
                c2[0] = sh.s.sfrag[(2 * bi) * 32 + lane];
                c2[1] = sh.s.sfrag[(2 * bi + 1) * 32 + lane];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) dmma(c2, a[ks], sh.u.f.gt[(4 * ks + q) * GTS + 8 * J + gq]);
                sh.s.sfrag[(2 * bi) * 32 + lane] = c2[0];
                sh.s.sfrag[(2 * bi + 1) * 32 + lane] = c2[1];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.lt_empty[ls]);
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
    for (int i = threadIdx.x; i < NA; i += IP_THREADS) HBw[(size_t)i * HB_PITCH + HB_PITCH - 1] = HBw[(size_t)i * HB_PITCH] + F.DD[i];
    for (int e = threadIdx.x; e < 20 * 32; e += IP_THREADS) sh.s.sfrag[e] = 0.0;
    fence_proxy_async();        // the pivots above (generic proxy) are read by the bulk copies (async proxy)
    __syncthreads();
    if (warp == 0) {
        if (!factor_chain(sh, F, g, tick)) sh.flag = 1;
    } else {
        factor_fill(sh, F, g, tick);
    }
    // ---- separator: S = M[sep, sep] + D_S - G D^-1 G^T, LDL^T in place (warp 1) ----
    double sf[20];
    if (warp == 1) {
#pragma unroll
        for (int e = 0; e < 20; ++e) sf[e] = sh.s.sfrag[e * 32 + lane];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 1024; e += IP_THREADS) {
        const int r = e >> 5, c = e & 31;
        const int lo = min(r, c), dist = abs(r - c);
        double v = F.HB[(size_t)(NA + lo) * HB_PITCH + dist];
        if (r == c) v += F.DD[NA + r];
        sh.s.Ss[r * 33 + c] = v;
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
        bool ok = true;
        double *row = sh.s.Ss + lane * 33;
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
            const double v = row[k];
            double *cb = sh.part[k & 1];
            cb[lane] = v;
            __syncwarp();
            const double d = cb[k];
            if (!(d > 0.0)) ok = false;
            const double w = fast_rcp(d);
            const double l = v * w;
#pragma unroll 4
            for (int c = k + 1; c < 32; ++c) row[c] = fma(-l, cb[c], row[c]);
            if (lane > k) row[k] = l;
            if (lane == k) sh.wS[k] = w;
        }
        if (!ok) sh.flag = 1;
    }
    __syncthreads();
    return sh.flag == 0;
}

// ---- sweep ring: one warp consumes 8-column units streamed HBM -> shared by bulk copies it issues itself ----
struct Ring {
    IpShared &sh;
    unsigned phase;
    uint64_t pol;
    __device__ Ring(IpShared &s) : sh(s), phase(s.ring_phase), pol(l2_evict_first_policy()) {}
    __device__ __forceinline__ void issue(int unit, const double *src, unsigned bytes) {
        const int sl = unit % RING_UNITS;
        mbar_expect_tx(&sh.ring_full[sl], bytes);
        tma_load_1d(sh.u.ring[sl], src, bytes, &sh.ring_full[sl], pol);
    }
    __device__ __forceinline__ const double *wait(int unit) {
        const int sl = unit % RING_UNITS;
        mbar_wait(&sh.ring_full[sl], (phase >> sl) & 1u);
        phase ^= 1u << sl;
        return sh.u.ring[sl];
    }
    __device__ __forceinline__ void close() { sh.ring_phase = phase; }
};

// forward sweep (warp 0): y = L^-1 g, right-looking; lane = row (circular)
__device__ __noinline__ void sweep_forward(IpShared &sh, const Factor F, const double *__restrict__ g) {
    const int lane = threadIdx.x & 31;
    const int NA = F.NA;
    const int nunits = (NA + SUB - 1) / SUB;
    Ring R(sh);
    if (lane == 0)
        for (int u = 0; u < 3 && u < nunits; ++u) R.issue(u, F.LT + (size_t)u * SUB * LTG, LT_UNIT_BYTES);
    double acc = (lane < NA) ? g[lane] : 0.0;
    double gk = (lane + 32 < NA) ? g[lane + 32] : 0.0;
    for (int u = 0; u < nunits; ++u) {
        const double *lt = R.wait(u);
        const int k0 = u * SUB, nst = min(SUB, NA - k0);
        int rho = (lane - k0 - 1) & 31;
        double l[SUB];
#pragma unroll
        for (int s = 0; s < SUB; ++s) l[s] = lt[s * LTG + ((rho - s) & 31)];
#pragma unroll
        for (int s = 0; s < SUB; ++s) {
            if (s < nst) {
                const int k = k0 + s;
                const double yk = __shfl_sync(FULL, acc, k & 31);
                if (lane == (k & 31)) {
                    F.YP[k] = yk;
                    acc = fma(-l[s], yk, gk);
                    gk = (k + 64 < NA) ? g[k + 64] : 0.0;
                } else {
                    acc = fma(-l[s], yk, acc);
                }
            }
        }
        __syncwarp();
        if (lane == 0 && u + 3 < nunits) R.issue(u + 3, F.LT + (size_t)(u + 3) * SUB * LTG, LT_UNIT_BYTES);
    }
    R.close();
}

// separator part of the forward sweep (warp 1):  gs = g_S - G (w y)
__device__ __noinline__ void sweep_sep_rhs(IpShared &sh, const Factor F, const double *__restrict__ g) {
    const int lane = threadIdx.x & 31;
    const int NA = F.NA;
    const int nunits = (NA + SUB - 1) / SUB;
    Ring R(sh);
    if (lane == 0)
        for (int u = 0; u < 4 && u < nunits; ++u) R.issue(u, F.GT + (size_t)u * SUB * GTS, GT_UNIT_BYTES);
    double s0 = 0.0, s1 = 0.0;
    for (int u = 0; u < nunits; ++u) {
        const int k0 = u * SUB, nst = min(SUB, NA - k0);
        const int kk = k0 + (lane & 7);
        const double zl = (kk < NA) ? F.YP[kk] * F.WP[kk] : 0.0;      // lanes 0..7 (replicated): z of the unit's columns
        const double *gt = R.wait(u);
#pragma unroll
        for (int s = 0; s < SUB; s += 2) {
            const double z0 = __shfl_sync(FULL, zl, s), z1 = __shfl_sync(FULL, zl, s + 1);
            if (s < nst) s0 = fma(gt[s * GTS + lane], z0, s0);
            if (s + 1 < nst) s1 = fma(gt[(s + 1) * GTS + lane], z1, s1);
        }
        __syncwarp();
        if (lane == 0 && u + 4 < nunits) R.issue(u + 4, F.GT + (size_t)(u + 4) * SUB * GTS, GT_UNIT_BYTES);
    }
    R.close();
    sh.gs[lane] = g[NA + lane] - (s0 + s1);
}

// separator solve and the right-hand side of the backward sweep (warp 1):
//   x_S = S^-1 gs;   t = (y - G^T x_S) w
__device__ __noinline__ void sweep_sep_solve(IpShared &sh, const Factor F, double *__restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int NA = F.NA;
    const int nunits = (NA + SUB - 1) / SUB;
    Ring R(sh);
    if (lane == 0)
        for (int u = 0; u < 4 && u < nunits; ++u) R.issue(u, F.GT + (size_t)u * SUB * GTS, GT_UNIT_BYTES);
    double a = sh.gs[lane];
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        const double yk = __shfl_sync(FULL, a, k);
        if (lane > k) a = fma(-sh.s.Ss[lane * 33 + k], yk, a);
    }
    a *= sh.wS[lane];
#pragma unroll 4
    for (int k = 31; k >= 0; --k) {
        const double xk = __shfl_sync(FULL, a, k);
        if (lane < k) a = fma(-sh.s.Ss[k * 33 + lane], xk, a);
    }
    x[NA + lane] = a;
    // c_k = sum_r G[r][k] x_S[r]: lane = (column k0 + (lane & 7), quarter lane >> 3 of the separator rows)
    const int kl = lane & 7, qr = lane >> 3;
    double xq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xq[i] = __shfl_sync(FULL, a, 8 * qr + i);
    for (int u = 0; u < nunits; ++u) {
        const int k = u * SUB + kl;
        const double yk = (k < NA) ? F.YP[k] : 0.0, wk = (k < NA) ? F.WP[k] : 0.0;
        const double *gt = R.wait(u) + kl * GTS + 8 * qr;
        double c = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) c = fma(gt[i], xq[i], c);
        c += __shfl_xor_sync(FULL, c, 8);
        c += __shfl_xor_sync(FULL, c, 16);
        if (qr == 0 && k < NA) F.TP[k] = (yk - c) * wk;
        __syncwarp();
        if (lane == 0 && u + 4 < nunits) R.issue(u + 4, F.GT + (size_t)(u + 4) * SUB * GTS, GT_UNIT_BYTES);
    }
    R.close();
}

// backward sweep (warp 0): x = L^-T t, right-looking from the last column; lane = row (circular).  Step k needs row k of
// L, i.e. the entries LT[j][k-1-j] of the 32 columns before k: five resident units + two in flight.
__device__ __noinline__ void sweep_backward(IpShared &sh, const Factor F, double *__restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int NA = F.NA;
    const int nunits = (NA + SUB - 1) / SUB;
    Ring R(sh);
    const int U0 = nunits - 1;
    if (lane == 0)
        for (int u = U0; u > U0 - 6 && u >= 0; --u) R.issue(u, F.LT + (size_t)u * SUB * LTG, LT_UNIT_BYTES);
    for (int u = U0; u > U0 - 4 && u >= 0; --u) R.wait(u);            // (units U0 .. U0-3; U-4 is waited per unit below)
    int row = NA - 1 - ((NA - 1 - lane) & 31);                        // the row == lane (mod 32) among the last 32
    double acc = (row >= 0) ? F.TP[row] : 0.0;
    double tk = (row - 32 >= 0) ? F.TP[row - 32] : 0.0;
    for (int U = U0; U >= 0; --U) {
        if (U - 4 >= 0) R.wait(U - 4);
        const int khi = min(NA - 1, U * SUB + SUB - 1);
#pragma unroll 2
        for (int k = khi; k >= U * SUB; --k) {
            const int kap = k & 31;
            const double xk = __shfl_sync(FULL, acc, kap);
            const int sigma = (kap - 1 - lane) & 31, j = k - 1 - sigma;
            const double l = (j >= 0) ? sh.u.ring[(j >> 3) % RING_UNITS][(j & 7) * LTG + sigma] : 0.0;
            if (lane == kap) {
                x[k] = xk;
                acc = fma(-l, xk, tk);                 // this lane moves on to row k - 32
                tk = (k - 64 >= 0) ? F.TP[k - 64] : 0.0;
            } else {
                acc = fma(-l, xk, acc);
            }
        }
        __syncwarp();
        if (lane == 0 && U - 6 >= 0) R.issue(U - 6, F.LT + (size_t)(U - 6) * SUB * LTG, LT_UNIT_BYTES);   // slot of unit U + 1
    }
    R.close();
}

// x = M^-1 g with the stored factor.  fused: the forward part was done inside factor() (predictor).
__device__ __noinline__ void solve(IpShared &sh, double *slab, const Layout &L, int n, const double *g, double *x, bool fused) {
    const Factor F = make_factor(slab, L, n);
    const int warp = threadIdx.x >> 5;
    fence_proxy_async();        // the ring area was last accessed through the generic proxy (factor hand-off buffers)
    __syncthreads();
    if (!fused) {
        if (warp == 0) sweep_forward(sh, F, g);
        __syncthreads();
        if (warp == 1) sweep_sep_rhs(sh, F, g);
        __syncwarp();
    }
    if (warp == 1) sweep_sep_solve(sh, F, x);
    __syncthreads();
    if (warp == 0) sweep_backward(sh, F, x);
    __syncthreads();
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
        for (int i = 0; i < LT_SLOTS; ++i) { mbar_init(&sh.lt_full[i], 1); mbar_init(&sh.lt_empty[i], 1); }
        for (int i = 0; i < RING_UNITS; ++i) mbar_init(&sh.ring_full[i], 1);
        sh.ring_phase = 0;
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
        const double lam0 = 1e-2 * gmax + 1e-300;
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
    if (reset) {
        unsigned long long z[24] = {0};
        if (cudaMemcpyToSymbol(g_prof, z, sizeof(z)) != cudaSuccess) return -1;
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
        const double *HB = slab + L.o_hb;
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
        const double lam0 = 1e-2 * gmax + 1e-300;
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
