// Pieces shared by the tuned pair kernels (hbt_kernels_v3.cuh): constants of the fast path,
// the literal slow path for pairs the fast path cannot decide, the per-warp survivor queue and
// small PTX helpers.  (The v2 kernels themselves — 4-warp CTAs, one tile pair per CTA — were
// replaced by the persistent single-warp kernels of v3; see DESIGN.md §5 for the history.)
#ifndef HBT_KERNELS_V2_CUH_
#define HBT_KERNELS_V2_CUH_

#include "hbt_kernels_v1.cuh"
#include "hbt_sort.cuh"

#define HBT_V2_WARPS 4
#define HBT_V2_IPL 4  // list-1 particles per lane
#define HBT_V2_TILE_I (32 * HBT_V2_IPL * HBT_V2_WARPS)  // 512
#define HBT_V2_TILE_J 128
#define HBT_V2_SUB (HBT_V2_TILE_I / HBT_V2_TILE_J)  // list-2 sub-tiles per 512 x 512 super-tile
#ifndef HBT_V2_LCAP
#define HBT_V2_LCAP 8                                // per-lane survivor list capacity (v3 kernels; the v4 kernel has its own)
#endif
#define HBT_V2_QCAP (32 + 32 * HBT_V2_LCAP)          // linear warp queue capacity

// constants of the fast path, derived on the host from HbtGrid
struct V2Const {
    double k2lo, k2hi;        // 4*KT_min_sq, 4*KT_max_sq (exact scalings)
    double kt4[HBT_MAX_KT];   // 4*kt_thr_sq[k]
    double W2;                // max(q_lo^2, q_hi^2)
    double g_abs;             // absolute part of the guard band
    double inv_dq;
    double gb_min;            // minimum guard in bin units: 2e-8/delta_q (twice the window inset)
    double nq_d;              // qnpts as a double
    double ub;                // -q_base/delta_q : u = q*inv_dq + ub
    double g0;                // guard, constant part, in bin units: gb_min + g_abs*inv_dq
    double gq;                // guard per unit of (|qx|+|qy|), in bin units: 2^-46 * inv_dq
    double gl;                // same for the q_long error bound (2^-45 * inv_dq)
    double inv_dkphi;         // 1/dK_phi (float-estimate path of the K_phi bin only; never decides an edge)
    float kt_min_f, inv_dkt_f;  // float estimate of the K_T bin (fixed up with the exact thresholds)
    int symmetric;            // |q_lo| == |q_hi| up to 2^-24 relative
    // ---- FP32 decision of mixed-event survivors (v3_mixed_f32): a pair is decided in float only
    // when every quantity is farther from every edge than a bound on the float evaluation error
    float f_inv_dq, f_ub;     // u = q * f_inv_dq + f_ub, bin units
    float f_gs;               // error bound (in units of 2^-24 GeV) -> bin units, with a safety factor 2
    float f_g0;               // constant part of the float guard, bin units: g0 + 8 * 2^-24 * nq
    float f_nkphi;            // n_Kphi as a float (conditioning test of the float K_phi estimate)
    float ktf[HBT_MAX_KT + 1];  // K_T thresholds in k2 space as floats: [0] = k2lo, [k] = kt4[k], [nKT] = k2hi
    int f32_mixed;            // 1: the float path is enabled (HBT_B200_F32MIX=0 disables it)
    // ---- q_inv mode (invariant_radius_flag = 1) on the tuned kernels, see v3_qinv_pair.  With s = -(q_E^2 - q_x^2 -
    // q_y^2 - q_z^2) evaluated as the reference does, its tests on q_inv = sqrt(s) (src :340-346, :597-600) are
    // monotone step functions of the double s; their steps are found once on the host by bisection over doubles on
    // the reference's own expressions (hbt_qinv_thresholds), like the K_T bins.
    double qinv_s_lo;         // q_inv > q_lo  <=>  s >= qinv_s_lo
    double qinv_s_hi;         // q_inv < q_hi  <=>  s <  qinv_s_hi
    const double *qinv_thr;   // device, [nq + 1]: thr[k] = smallest s (>= qinv_s_lo) whose bin index is >= k; thr[0] = qinv_s_lo
    float qinv_w2_f;          // max(q_hi, 0)^2 rounded up: the float prefilter keeps s_f <= qinv_w2_f + margin
    float f_qbase;            // q_base as a float (bin estimate only)
    // replicated q_inv accumulators (the histograms have n_KT x nq bins: reductions from every warp of the device
    // into a few hundred addresses would serialise in the L2); lane l of CTA b adds into replica (32 b + l) mod R,
    // hbt_qinv_fold sums the replicas into the histograms after each launch
    unsigned long long *qrep_u64;  // [R][2][nKT * nq]: same-event count, mixed-event count
    double *qrep_f64;              // [R][2][nKT * nq]: sum q_inv, sum cos
    int qrep_n;                    // R (a power of two, >= 32)
};

__host__ inline V2Const hbt_v2_consts(const HbtGrid &g) {
    V2Const c;
    c.k2lo = 4.0 * g.KT_min_sq;
    c.k2hi = 4.0 * g.KT_max_sq;
    for (int k = 0; k < HBT_MAX_KT; k++) c.kt4[k] = 4.0 * g.kt_thr_sq[k];
    const double a = g.q_lo * g.q_lo, b = g.q_hi * g.q_hi;
    c.W2 = a > b ? a : b;
    const double m = fabs(g.q_lo) > fabs(g.q_hi) ? fabs(g.q_lo) : fabs(g.q_hi);
    c.g_abs = m * 5.7e-14;  // 2^-44
    c.inv_dq = g.inv_dq;
    c.gb_min = 2e-8 * g.inv_dq;
    c.nq_d = static_cast<double>(g.nq);
    c.ub = -g.q_base * g.inv_dq;
    c.g0 = c.gb_min + c.g_abs * g.inv_dq;
    c.gq = 1.5e-14 * g.inv_dq;
    c.gl = 2.9e-14 * g.inv_dq;
    c.inv_dkphi = g.dKphi > 0.0 ? 1.0 / g.dKphi : 0.0;
    c.kt_min_f = static_cast<float>(g.KT_min);
    c.inv_dkt_f = static_cast<float>(1.0 / g.dKT);
    c.symmetric = (g.q_lo < 0.0 && g.q_hi > 0.0 && fabs(a - b) <= 5.9e-8 * c.W2) ? 1 : 0;
    const double u24 = 5.9604644775390625e-8;  // 2^-24, unit roundoff of binary32
    c.f_inv_dq = static_cast<float>(g.inv_dq);
    c.f_ub = static_cast<float>(c.ub);
    c.f_gs = static_cast<float>(2.0 * u24 * g.inv_dq * 1.001);
    c.f_g0 = static_cast<float>((c.g0 + 8.0 * u24 * (c.nq_d + 2.0)) * 1.001);
    c.f_nkphi = static_cast<float>(g.nKphi);
    for (int k = 0; k <= HBT_MAX_KT; k++) c.ktf[k] = 0.f;
    const int nkt = g.nKT < HBT_MAX_KT ? g.nKT : HBT_MAX_KT;
    c.ktf[0] = static_cast<float>(c.k2lo);
    for (int k = 1; k < nkt; k++) c.ktf[k] = static_cast<float>(c.kt4[k]);
    c.ktf[nkt] = static_cast<float>(c.k2hi);
    c.f32_mixed = g.qinv ? 0 : 1;  // (q_inv mode evaluates every survivor in binary64: the q_inv histograms need s)
    c.qinv_s_lo = 0.0; c.qinv_s_hi = 0.0; c.qinv_thr = nullptr;
    {
        const double wq = g.q_hi > 0.0 ? g.q_hi : 0.0;
        c.qinv_w2_f = static_cast<float>(wq * wq * (1.0 + 2.4e-7));
    }
    c.f_qbase = static_cast<float>(g.q_base);
    c.qrep_u64 = nullptr; c.qrep_f64 = nullptr; c.qrep_n = 0;
    return c;
}

// The window need not be symmetric about zero: the prefilter, the culling and the pT range restriction test |q_out|, |q_side| against
// W = max(|q_lo|, |q_hi|) — a superset of any window [q_lo, q_hi] — and the drain decides in bin units
// u = (q - q_base)/delta_q, where the window is [eps, nq - eps] whatever its position.  The float decision of
// mixed-event survivors budgets 8 x 2^-24 x (nq + 2) for the rounding of u and of the offset -q_base/delta_q, so a
// window that lies many bin-counts away from zero (|q_base|/delta_q > 4 (nq + 2)) goes to the literal kernels.
__host__ inline bool hbt_v2_supported(const HbtGrid &g) {
    // K_T bins at least 1e-4 of KT_max wide: the float estimate of the K_T bin is then within one
    // bin (relative error of the estimate ~4e-7); bin index in 32 bits
    const double kt_max = g.KT_min + g.dKT * g.nKT;
    const V2Const c = hbt_v2_consts(g);
    return fabs(c.ub) <= 4.0 * (c.nq_d + 2.0) && g.dq > 1e-6 && g.dKT > 1e-4 * kt_max && g.nbins < (1ll << 31);
}

// global-memory copy of everything the non-inlined device functions need (passing the
// kernel's by-value parameter structs by reference would copy them to each thread's stack)
struct V2Dev {
    HbtGrid g;
    V2Const c;
    HbtAccum acc;
    const unsigned char *closed;  // [2*nslab], always allocated (all zero while no cap was reached)
};

struct V2Counters {
    unsigned nB, nC, nD, nE;
};

// the literal chain for a pair the fast path could not decide; counts and accumulates
template <bool MIXED>
__device__ __noinline__ void v2_slow_pair(const V2Dev *__restrict__ dv, const double *a, const double *b,
                                          double psi_ref, V2Counters &n) {
    const HbtGrid &g = dv->g;
    const HbtAccum &acc = dv->acc;
    PairBin pb;
    const int st = pair_literal(g, a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3], MIXED, psi_ref, pb);
    if (st == PAIR_REJ_KT) return;
    if (st == PAIR_DEFER) { defer_pair(acc, a, b, psi_ref, MIXED ? 1 : 0); return; }
    n.nB++;
    if (st == PAIR_REJ_QO) return;
    n.nC++;
    if (st == PAIR_REJ_QS) return;
    n.nD++;
    if (st == PAIR_REJ_QL) return;
    n.nE++;
    if (st == PAIR_REJ_PHI) return;
    if (dv->closed[pb.slab + (MIXED ? g.nslab : 0)]) return;
    const long long bin = bin_index(g, pb);
    if (MIXED) {
        atomicAdd(&acc.den_count[bin], 1ull);
    } else {
        const double cv = pair_cos(g, a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3], a[4] - b[4], a[5] - b[5],
                                   a[6] - b[6], a[7] - b[7]);
        atomicAdd(&acc.num_count[bin], 1ull);
        atomicAdd(&acc.sum_qo[bin], pb.qo);
        atomicAdd(&acc.sum_qs[bin], pb.qs);
        atomicAdd(&acc.sum_ql[bin], pb.ql);
        atomicAdd(&acc.num_cos[bin], cv);
    }
}

// state of one warp's survivor bookkeeping
struct V2Queue {
    unsigned list_addr;   // shared-memory address of this lane's private list: entry m at list_addr + 64 m
    unsigned cur;         // shared-memory address of the next free slot (advances by 64 bytes)
    int qcount;           // entries in the warp's linear queue (warp-uniform)
};

__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ unsigned lds_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// survivor-queue entries are 16 bits: list-1 slot << 8 | list-2 position (tiles of at most 256 x 256)
__device__ __forceinline__ unsigned lds_u16(unsigned addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u16(unsigned addr, unsigned v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<unsigned short>(v)) : "memory");
}

__device__ __forceinline__ void sts_u32(unsigned addr, unsigned v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// hides a value's origin from the compiler so that it stays in a register instead of being
// rematerialised inside the hot loop
__device__ __forceinline__ unsigned opaque_u32(unsigned v) {
    unsigned r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// cnt += 1 when flag: one predicated add
__device__ __forceinline__ void inc_if(unsigned &cnt, bool flag) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, 0;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(cnt) : "r"(static_cast<int>(flag)));
}

#endif  // HBT_KERNELS_V2_CUH_
