// v2 pair kernels (sm_100a): prefilter + warp-level compaction + guarded fast path.
//
// Why: the literal chain (v1) spends ~150 FP64-pipe instructions per pair and runs 2/3 of
// its lanes idle (ncu: 10.5 of 32 threads active per instruction), because 93 % of the
// pairs leave the chain early at different stages.  v2 splits the work in two:
//
//  1. PREFILTER, every pair, no divergence, 13 FP64-pipe instructions.  Uses only
//     (px, py, pT^2) of the two particles:
//        s = p_i + p_j           k2 = rn(rn(sx^2) + rn(sy^2))  ( = 4 K_perp_sq, bit-exact: the
//                                      reference's 0.5 factors are exact power-of-two scalings)
//        d = pT_i^2 - pT_j^2     ( = q.s = 2 K_perp q_out )
//        x = pxj*pyi - pxi*pyj   ( 2x = q x s = 2 K_perp q_side )
//     K_T cut: exact compare of k2.  q_out / q_side windows: d^2 and 4x^2 against W^2 k2 with a
//     relative band of ~3e-6 (integer compare of the high words); a pair is dropped only when
//     it CERTAINLY fails; anything inside a band goes on.  ~92 % of the pairs end here.
//  2. DRAIN.  Each lane appends its survivors (i,j) to a private shared-memory list (one
//     predicated store, no cross-lane traffic in the hot loop).  When a list is about to
//     fill, the warp compacts all lists into a linear queue (one prefix sum) and processes it
//     32 pairs at a time with all lanes busy: K_T bin by exact thresholds in k2 space (no
//     sqrt/divide), q_out/q_side/q_long from FMA arithmetic and rsqrt (a few ulp from the
//     reference's values), each compared against the window and bin edges with a guard band
//     that bounds the distance to the reference's own rounding.  Inside the guard
//     (probability ~1e-12 per pair) the pair is re-evaluated with the literal chain of
//     hbt_pair.cuh, so every bin decision equals the reference's.  Accepted pairs add
//     cos(q.dx/hbarc) and the q sums with global atomics (histograms are L2 resident).
//
// The accumulated q_out/q_side/q_long are the fast-path values: within ~1e-15 relative of the
// reference's, far inside the 1e-10 tolerance of the sums.
#ifndef HBT_KERNELS_V2_CUH_
#define HBT_KERNELS_V2_CUH_

#include "hbt_kernels_v1.cuh"
#include "hbt_sort.cuh"

#define HBT_V2_WARPS 4
#define HBT_V2_IPL 4  // list-1 particles per lane
#define HBT_V2_TILE_I (32 * HBT_V2_IPL * HBT_V2_WARPS)  // 512
#define HBT_V2_TILE_J 128
#define HBT_V2_SUB (HBT_V2_TILE_I / HBT_V2_TILE_J)  // list-2 sub-tiles per 512 x 512 super-tile
#define HBT_V2_LCAP 8                                // per-lane survivor list capacity
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
    int symmetric;            // |q_lo| == |q_hi| up to 2^-24 relative
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
    c.symmetric = (g.q_lo < 0.0 && g.q_hi > 0.0 && fabs(a - b) <= 5.9e-8 * c.W2) ? 1 : 0;
    return c;
}

// v2 handles the 3-D histograms with a window that is symmetric about zero; everything else
// (q_inv mode, one-sided windows) runs on the literal v1 kernels
__host__ inline bool hbt_v2_supported(const HbtGrid &g) {
    return !g.qinv && hbt_v2_consts(g).symmetric && g.dq > 1e-6;
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

// outcome of a guarded comparison of a fast-path q against the window and the bin grid
enum : int { Q_REJECT = 0, Q_OK = 1, Q_UNSURE = 2 };

// In units of bins, u = (q - q_base)/delta_q, the reference's window is [eps, nq - eps] with
// eps = 1e-8/delta_q (src :363-364) and its bin edges are the integers.  A pair is decided on
// the fast path only when u is farther than gb from every integer, gb >= 2 eps: that single
// test covers the bin edges, both window edges and their '>' / '>=' distinction.  Everything
// within gb of an integer (a ~1e-5 fraction of the survivors) takes the literal chain.
__device__ __forceinline__ int classify_q(const HbtGrid &g, const V2Const &c, double q, double gb, int &idx) {
    const double u = (q - g.q_base) * c.inv_dq;
    if (!(u > -gb && u < c.nq_d + gb)) return Q_REJECT;  // certainly outside (NaN too)
    const double fl = floor(u);
    const double fr = u - fl;
    if (fr < gb || fr > 1.0 - gb) return Q_UNSURE;
    idx = __double2int_rz(fl);
    return Q_OK;
}

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

// one queued survivor through the guarded fast path.  si/sj are the SoA tiles in shared memory
// (component stride TI / TJ doubles).
// ORIENT (same-event list sorted in momentum space): si_o / sj_o hold the particles' positions
// in the reference's gather order; the pair is (earlier, later) in THAT order, so when the
// tile order disagrees the roles are swapped, i.e. q -> -q (K, k2, |q| and cos(q.dx) are even).
template <bool MIXED, int NC, bool ORIENT>
__device__ __forceinline__ void v2_drain_pair(const HbtGrid &g, const V2Const &c, const HbtAccum &acc,
                                              const unsigned char *__restrict__ closed,
                                              const V2Dev *__restrict__ dv, const double *__restrict__ si,
                                              const double *__restrict__ sj, const unsigned *__restrict__ si_o,
                                              const unsigned *__restrict__ sj_o, int il, int jl, double psi_ref,
                                              V2Counters &n) {
    constexpr int TI = HBT_V2_TILE_I, TJ = HBT_V2_TILE_J;
    const bool flip = ORIENT && (si_o[il] > sj_o[jl]);
    const double ax = si[il], ay = si[TI + il], bx = sj[jl], by = sj[TJ + jl];
    const double sx = __dadd_rn(ax, bx), sy = __dadd_rn(ay, by);
    const double k2 = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
    bool unsure = false;  // (the K_T cut was decided exactly by the prefilter)

    const double qx = ax - bx, qy = ay - by;
    const double d = fma(qx, sx, qy * sy);     // 2 K_perp q_out
    const double e = fma(qy, sx, -(qx * sy));  // 2 K_perp q_side
    const double r = rsqrt(k2);                // 1 / (2 K_perp)
    double qo = d * r, qs = e * r;
    if (flip) { qo = -qo; qs = -qs; }
    // |fast - reference| <= ~8 ulp of (|qx|+|qy|); guard = 2^-46 relative + 2^-44 |window|,
    // in bin units and never below gb_min
    const double gt = fmax(c.gb_min, fma(fabs(qx) + fabs(qy), 1.5e-14, c.g_abs) * c.inv_dq);
    int io = 0, is = 0, il_ = 0;
    int stage = 1;  // passed K_T
    double ql = 0.0;
    double az = 0.0, bz = 0.0, aE = 0.0, bE = 0.0;
    const int co = classify_q(g, c, qo, gt, io);
    if (co == Q_UNSURE) unsure = true;
    if (co == Q_OK) {
        stage = 2;
        const int cs = classify_q(g, c, qs, gt, is);
        if (cs == Q_UNSURE) unsure = true;
        if (cs == Q_OK) {
            stage = 3;
            az = si[2 * TI + il]; aE = si[3 * TI + il]; bz = sj[2 * TJ + jl]; bE = sj[3 * TJ + jl];
            const double qz = az - bz;
            if (g.boost) {
                // q_long = gamma (q_z - beta q_E) = (K_E q_z - K_z q_E) / Mt, src :383-390
                const double qE = aE - bE, sz = az + bz, sE = aE + bE;
                const double m2 = (sE - sz) * (sE + sz);  // 4 Mt^2 without cancellation
                const double r2 = rsqrt(m2);
                const double t1 = sE * qz, t2 = sz * qE;
                ql = (t1 - t2) * r2;
                if (flip) ql = -ql;
                const double ch = sE * r2;  // cosh of the pair rapidity: error amplification
                const double gl = fmax(c.gb_min, fma((fabs(t1) + fabs(t2)) * r2 * fma(2.0 * ch, ch, 1.0), 2.9e-14, c.g_abs) * c.inv_dq);
                const int cl = (m2 > 0.0) ? classify_q(g, c, ql, gl, il_) : Q_UNSURE;
                if (cl == Q_UNSURE) unsure = true;
                if (cl == Q_OK) stage = 4;
            } else {
                // q_long = q_z exactly as the reference has it: use its own comparisons
                ql = flip ? -qz : qz;
                if (in_window(ql, g.q_lo, g.q_hi, MIXED)) {
                    il_ = __double2int_rz(__ddiv_rn(__dsub_rn(ql, g.q_base), g.dq));
                    if (il_ < g.nq) stage = 4;
                }
            }
        }
    }
    int iK = 0;  // exact K_T bin: thresholds of int((sqrt(K_perp_sq)-KT_min)/dKT) in k2 space
    if (stage == 4)
        for (int k = 1; k < g.nKT; k++) iK += (k2 >= c.kt4[k]) ? 1 : 0;
    int slab = iK;
    if (!unsure && stage == 4 && g.az) {
        const double Kx = 0.5 * sx, Ky = 0.5 * sy;
        double dphi = __dsub_rn(atan2(Ky, Kx), psi_ref);
        while (dphi < 0.) dphi = __dadd_rn(dphi, g.two_pi);
        while (dphi > g.two_pi) dphi = __dsub_rn(dphi, g.two_pi);
        const double u = __ddiv_rn(dphi, g.dKphi);
        const int iphi = __double2int_rz(u);
        if (fabs(u - rint(u)) < 1e-9) unsure = true;  // the literal path defers it to the host
        else if (iphi < 0 || iphi >= g.nKphi) stage = 5;  // counted through q_long, then dropped
        else slab = iK * g.nKphi + iphi;
    }
    if (unsure) {
        double a8[8], b8[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double va = k < NC ? si[(k < NC ? k : 0) * TI + il] : 0.0;
            const double vb = k < NC ? sj[(k < NC ? k : 0) * TJ + jl] : 0.0;
            a8[k] = flip ? vb : va;
            b8[k] = flip ? va : vb;
        }
        V2Counters tmp = {0, 0, 0, 0};  // keeps n itself out of local memory
        v2_slow_pair<MIXED>(dv, a8, b8, psi_ref, tmp);
        n.nB += tmp.nB; n.nC += tmp.nC; n.nD += tmp.nD; n.nE += tmp.nE;
        return;
    }
    n.nB++;
    if (stage >= 2) n.nC++;
    if (stage >= 3) n.nD++;
    if (stage >= 4) n.nE++;
    if (stage != 4) return;
    if (closed && closed[slab + (MIXED ? g.nslab : 0)]) return;  // needed_number_of_pairs reached earlier
    const long long bin = ((static_cast<long long>(slab) * g.nq + io) * g.nq + is) * g.nq + il_;
    if (MIXED) {
        atomicAdd(&acc.den_count[bin], 1ull);
    } else {
        const double xd = si[(4 % NC) * TI + il] - sj[(4 % NC) * TJ + jl];
        const double yd = si[(5 % NC) * TI + il] - sj[(5 % NC) * TJ + jl];
        const double zd = si[(6 % NC) * TI + il] - sj[(6 % NC) * TJ + jl];
        const double td = si[(7 % NC) * TI + il] - sj[(7 % NC) * TJ + jl];
        const double cv = pair_cos(g, qx, qy, az - bz, aE - bE, xd, yd, zd, td);
        atomicAdd(&acc.num_count[bin], 1ull);
        atomicAdd(&acc.sum_qo[bin], qo);
        atomicAdd(&acc.sum_qs[bin], qs);
        atomicAdd(&acc.sum_ql[bin], ql);
        atomicAdd(&acc.num_cos[bin], cv);
    }
}

// state of one warp's survivor bookkeeping
struct V2Queue {
    unsigned *lane_list;  // this lane's private list: entry m at lane_list[32*m]
    unsigned list_addr;   // its 32-bit shared-memory address
    unsigned cur;         // shared-memory address of the next free slot (advances by 128 bytes)
    unsigned *wq;         // the warp's linear queue
    int qcount;           // entries in wq (warp-uniform)
    unsigned kept;        // survivors queued so far (warp-uniform)
};

__device__ __forceinline__ void sts_u32(unsigned addr, unsigned v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// cnt += 1 when flag: one predicated add
__device__ __forceinline__ void inc_if(unsigned &cnt, bool flag) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, 0;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(cnt) : "r"(static_cast<int>(flag)));
}

// non-inlined twin of v2_drain_pair for the once-per-tile final flush (keeps one inlined copy
// of the drain per kernel variant); reads its constants from the global-memory copy
template <bool MIXED, int NC, bool ORIENT>
__device__ __noinline__ void v2_drain_pair_cold(const V2Dev *__restrict__ dv, const double *__restrict__ si,
                                                const double *__restrict__ sj, const unsigned *__restrict__ si_o,
                                                const unsigned *__restrict__ sj_o, int il, int jl, double psi_ref,
                                                V2Counters &n) {
    V2Counters tmp = {0, 0, 0, 0};
    v2_drain_pair<MIXED, NC, ORIENT>(dv->g, dv->c, dv->acc, dv->closed, dv, si, sj, si_o, sj_o, il, jl, psi_ref, tmp);
    n.nB += tmp.nB; n.nC += tmp.nC; n.nD += tmp.nD; n.nE += tmp.nE;
}

// Compact the per-lane lists into the warp's linear queue (one prefix sum) and process it 32
// survivors at a time; FINAL also processes the last partial batch.
template <bool MIXED, bool FINAL, bool ORIENT>
__device__ __forceinline__ void v2_flush(const HbtGrid &g, const V2Const &c, const HbtAccum &acc,
                                         const unsigned char *__restrict__ closed, const V2Dev *__restrict__ dv, const double *__restrict__ si,
                                         const double *__restrict__ sj, const unsigned *__restrict__ si_o,
                                         const unsigned *__restrict__ sj_o, int lane, double psi_ref, V2Queue &Q,
                                         V2Counters &n) {
    constexpr int NC = MIXED ? 4 : 8;
    const int cnt = static_cast<int>(Q.cur - Q.list_addr) >> 7;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned *dst = Q.wq + Q.qcount + (incl - cnt);
    for (int m = 0; m < cnt; m++) dst[m] = Q.lane_list[32 * m];
    Q.cur = Q.list_addr;
    Q.qcount += total;
    Q.kept += static_cast<unsigned>(total);
    __syncwarp();
    while (Q.qcount >= 32 || (FINAL && Q.qcount > 0)) {
        const int take = min(32, Q.qcount);
        const int base = Q.qcount - take;
        if (lane < take) {
            const unsigned e = Q.wq[base + lane];
            const int il = static_cast<int>(e >> 16), jl = static_cast<int>(e & 0xffffu);
            if (FINAL) {
                V2Counters tmp = {0, 0, 0, 0};
                v2_drain_pair_cold<MIXED, NC, ORIENT>(dv, si, sj, si_o, sj_o, il, jl, psi_ref, tmp);
                n.nB += tmp.nB; n.nC += tmp.nC; n.nD += tmp.nD; n.nE += tmp.nE;
            } else {
                v2_drain_pair<MIXED, NC, ORIENT>(g, c, acc, closed, dv, si, sj, si_o, sj_o, il, jl, psi_ref, n);
            }
        }
        Q.qcount = base;
        __syncwarp();
    }
}

// The hot loop over one list-2 tile.  DIAG: same-event tile that touches the diagonal (only
// j > i pairs count, src :301).  FLOOR: tile pair in which the prefilter's error bound is not
// negligible against the smallest K_T (KT_min ~ 0 or huge momenta): pairs below k2_floor skip
// the window prefilter.
template <bool MIXED, bool DIAG, bool FLOOR, bool STATS>
__device__ __forceinline__ void v2_tile_loop(const HbtGrid &g, const V2Const &c, const HbtAccum &acc,
                                             const unsigned char *__restrict__ closed, const V2Dev *__restrict__ dv, const double *__restrict__ si,
                                             const double *__restrict__ sj, const double *__restrict__ sjt,
                                             const unsigned *__restrict__ si_o, const unsigned *__restrict__ sj_o, int nj,
                                             long long i0, long long j0, int lane, int warp, double k2_floor,
                                             double psi_ref, V2Queue &Q, V2Counters &n, unsigned &cntKT,
                                             unsigned &cntRS) {
    constexpr int TI = HBT_V2_TILE_I, TJ = HBT_V2_TILE_J, IPL = HBT_V2_IPL;
    constexpr bool ORIENT = !MIXED && !STATS;
    double ax[IPL], ay[IPL], at[IPL];
    unsigned ent[IPL];
    long long ig[IPL];
#pragma unroll
    for (int s = 0; s < IPL; s++) {
        const int il = warp * (32 * IPL) + s * 32 + lane;
        ax[s] = si[il];
        ay[s] = si[TI + il];
        at[s] = fma(ax[s], ax[s], ay[s] * ay[s]);
        ent[s] = static_cast<unsigned>(il) << 16;
        ig[s] = i0 + il;
    }
    const double k2lo = c.k2lo, k2hi = c.k2hi, W2 = c.W2;
    const unsigned lim = Q.list_addr + 128u * (HBT_V2_LCAP - IPL);

    for (int j = 0; j < nj; j++) {
        const double bx = sj[j], by = sj[TJ + j], bt = sjt[j];
#pragma unroll
        for (int s = 0; s < IPL; s++) {
            const double sx = __dadd_rn(ax[s], bx), sy = __dadd_rn(ay[s], by);
            const double k2 = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
            bool kt = (k2 >= k2lo) && (k2 <= k2hi);  // exact K_T cut (NaN padding rows fail)
            if (DIAG) kt = kt && (j0 + j > ig[s]);
            const double d = at[s] - bt;
            const double x = fma(bx, ay[s], -(ax[s] * by));
            const double d2 = d * d, x2 = x * x, w = W2 * k2;
            const int hw = __double2hiint(w);
            const int dd = __double2hiint(d2) - hw;               // d^2   vs W^2 k2
            const int dx = __double2hiint(x2) + 0x00200000 - hw;  // 4 x^2 vs W^2 k2
            bool rej_o = dd > 1;                 // q_out certainly outside the window
            bool rej_s = (dd < -1) && (dx > 1);  // q_out certainly inside, q_side certainly outside
            if (FLOOR) {
                const bool tiny = k2 < k2_floor;
                rej_o = rej_o && !tiny;
                rej_s = rej_s && !tiny;
            }
            const bool keep = kt && !(rej_o || rej_s);
            if (STATS) {  // exact K_T-pass and q_out-pass populations (instrumented runs only)
                inc_if(cntKT, kt);
                inc_if(cntRS, kt && rej_s);
            }
            if (keep) {
                sts_u32(Q.cur, ent[s] | static_cast<unsigned>(j));
                Q.cur += 128u;
            }
        }
        if (__any_sync(0xffffffffu, Q.cur > lim)) v2_flush<MIXED, false, ORIENT>(g, c, acc, closed, dv, si, sj, si_o, sj_o, lane, psi_ref, Q, n);
    }
    v2_flush<MIXED, true, ORIENT>(g, c, acc, closed, dv, si, sj, si_o, sj_o, lane, psi_ref, Q, n);
}

#ifndef HBT_V2_MINB_MIXED
#define HBT_V2_MINB_MIXED 4
#endif
// STATS = true : instrumented run — list in the reference's order, every pair goes through the
//                prefilter, the stage populations B, C, D (passed K_T, q_out, q_side) are exact.
// STATS = false: production — the same-event list is Morton-sorted (orig = reference order,
//                bbox = boxes of its 128-particle tiles), tile pairs whose boxes cannot hold an
//                accepted pair are skipped at CTA and at warp level; only the populations that
//                cost nothing (all pairs, passed q_long, accepted) are kept.
template <bool MIXED, bool STATS>
__global__ void __launch_bounds__(32 * HBT_V2_WARPS, MIXED ? HBT_V2_MINB_MIXED : 4)
hbt_pairs_v2(const double *__restrict__ p1, const double *__restrict__ p2, long long n_same,
             const HbtMixSeg *__restrict__ segs, const HbtGrid g, const V2Const c,
             const V2Dev *__restrict__ dv, const HbtAccum acc, const double psi_ref,
             const unsigned long long total_pairs, const unsigned char *__restrict__ closed,
             const unsigned *__restrict__ orig, const HbtBBox *__restrict__ bbox) {
    constexpr bool SORTED = !MIXED && !STATS;
    constexpr int NC = MIXED ? 4 : 8;
    constexpr int TI = HBT_V2_TILE_I, TJ = HBT_V2_TILE_J, NT = 32 * HBT_V2_WARPS;
    extern __shared__ __align__(16) unsigned char dyn[];
    double *si = reinterpret_cast<double *>(dyn);  // [NC][TI]
    double *sj = si + NC * TI;                      // [NC][TJ]
    double *sjt = sj + NC * TJ;                     // [TJ] pT^2 of the list-2 tile
    double *s_max = sjt + TJ;                       // [2*WARPS]
    unsigned *lq = reinterpret_cast<unsigned *>(s_max + 2 * HBT_V2_WARPS);  // [WARPS][LCAP][32]
    unsigned *wq = lq + HBT_V2_WARPS * HBT_V2_LCAP * 32;                    // [WARPS][QCAP]
    unsigned *si_o = wq + HBT_V2_WARPS * HBT_V2_QCAP;                       // [TI] reference-order index
    unsigned *sj_o = si_o + TI;                                             // [TJ]

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    long long i0, j0;
    int ni, nj;
    bool diag = false;
    double rc = 1.0, rs = 0.0;
    if (MIXED) {
        const HbtMixSeg sg = segs[find_segment(segs, static_cast<int>(n_same), blockIdx.x)];
        const int local = static_cast<int>(blockIdx.x - sg.block0);
        const int ti = local / sg.tiles_j, tj = local - ti * sg.tiles_j;
        i0 = sg.i0 + static_cast<long long>(ti) * TI;
        j0 = sg.j0 + static_cast<long long>(tj) * TJ;
        ni = min(TI, sg.ni - ti * TI);
        nj = min(TJ, sg.nj - tj * TJ);
        rc = sg.c; rs = sg.s;
    } else {
        // upper triangle in 512 x 512 super-tiles, each split into HBT_V2_SUB list-2 sub-tiles
        const long long T = (n_same + TI - 1) / TI;
        const long long super = blockIdx.x / HBT_V2_SUB;
        const int sub = static_cast<int>(blockIdx.x - super * HBT_V2_SUB);
        int ti, tj;
        tri_decode(super, T, ti, tj);
        i0 = static_cast<long long>(ti) * TI;
        j0 = static_cast<long long>(tj) * TI + static_cast<long long>(sub) * TJ;
        ni = static_cast<int>(min(static_cast<long long>(TI), n_same - i0));
        nj = static_cast<int>(min(static_cast<long long>(TJ), n_same - j0));
        diag = (ti == tj);
        if (diag && j0 + nj - 1 <= i0) nj = 0;  // sub-tile entirely at or below the diagonal
    }
    if (blockIdx.x == 0 && t == 0) atomicAdd(&acc.stage[MIXED ? 6 : 0], total_pairs);
    if (nj <= 0 || ni <= 0) return;
    bool warp_culled = false;
    if (SORTED) {
        // CTA level: the union of the list-1 sub-tile boxes against the list-2 tile box
        const long long ti0 = i0 / HBT_BBOX_TILE;
        const int nsub = (ni + HBT_BBOX_TILE - 1) / HBT_BBOX_TILE;
        const HbtBBox bj = bbox[j0 / HBT_BBOX_TILE];
        HbtBBox u = bbox[ti0];
        for (int q = 1; q < nsub; q++) {
            const HbtBBox b = bbox[ti0 + q];
            u.xlo = fmin(u.xlo, b.xlo); u.xhi = fmax(u.xhi, b.xhi);
            u.ylo = fmin(u.ylo, b.ylo); u.yhi = fmax(u.yhi, b.yhi);
        }
        if (hbt_boxes_culled(u, bj, c.W2, c.k2lo, c.k2hi)) return;
        // warp level: this warp's 128 list-1 particles
        warp_culled = (warp >= nsub) || hbt_boxes_culled(bbox[ti0 + warp], bj, c.W2, c.k2lo, c.k2hi);
    }

    // ---- stage both tiles (SoA).  Rows beyond the tile end are NaN: they fail the K_T cut.
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double tmax = 0.0, tmaxj = 0.0;
    for (int k = t; k < TI; k += NT) {
        if (k < ni) {
            const double2 *src = reinterpret_cast<const double2 *>(p1 + 8 * (i0 + k));
            const double2 v0 = src[0], v1 = src[1];
            si[k] = v0.x; si[TI + k] = v0.y; si[2 * TI + k] = v1.x; si[3 * TI + k] = v1.y;
            tmax = fmax(tmax, fma(v0.x, v0.x, v0.y * v0.y));
            if (SORTED) si_o[k] = orig[i0 + k];
            if (!MIXED) {
                const double2 v2 = src[2], v3 = src[3];
                si[(4 % NC) * TI + k] = v2.x; si[(5 % NC) * TI + k] = v2.y;
                si[(6 % NC) * TI + k] = v3.x; si[(7 % NC) * TI + k] = v3.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < NC; q++) si[q * TI + k] = nan;
            if (SORTED) si_o[k] = 0u;
        }
    }
    for (int k = t; k < TJ; k += NT) {
        if (k < nj) {
            const double2 *src = reinterpret_cast<const double2 *>(p2 + 8 * (j0 + k));
            const double2 v0 = src[0], v1 = src[1];
            double x = v0.x, y = v0.y;
            if (MIXED) {  // rotation of the partner event, src/HBT_correlation.cpp:522-523
                x = __dsub_rn(__dmul_rn(v0.x, rc), __dmul_rn(v0.y, rs));
                y = __dadd_rn(__dmul_rn(v0.x, rs), __dmul_rn(v0.y, rc));
            }
            sj[k] = x; sj[TJ + k] = y; sj[2 * TJ + k] = v1.x; sj[3 * TJ + k] = v1.y;
            const double pt2 = fma(x, x, y * y);
            sjt[k] = pt2;
            tmaxj = fmax(tmaxj, pt2);
            if (SORTED) sj_o[k] = orig[j0 + k];
            if (!MIXED) {
                const double2 v2 = src[2], v3 = src[3];
                sj[(4 % NC) * TJ + k] = v2.x; sj[(5 % NC) * TJ + k] = v2.y;
                sj[(6 % NC) * TJ + k] = v3.x; sj[(7 % NC) * TJ + k] = v3.y;
            }
        }
    }
    // S = max pT^2 (list 1) + max pT^2 (list 2) bounds the rounding error of d and x
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        tmaxj = fmax(tmaxj, __shfl_xor_sync(0xffffffffu, tmaxj, o));
    }
    if (lane == 0) { s_max[warp] = tmax; s_max[HBT_V2_WARPS + warp] = tmaxj; }
    __syncthreads();  // also publishes the tiles
    double S1 = 0.0, S2 = 0.0;
#pragma unroll
    for (int w = 0; w < HBT_V2_WARPS; w++) { S1 = fmax(S1, s_max[w]); S2 = fmax(S2, s_max[HBT_V2_WARPS + w]); }
    const double S = S1 + S2;
    // Below k2_floor the relative error of d (|err| <= 2^-50 S) against the window edge
    // W*sqrt(k2) could exceed the 2^-21 band of the high-word compare: such pairs are not
    // prefiltered.  (k2_floor = 2^-54 S^2 / W2; far below 4*KT_min_sq unless KT_min ~ 0.)
    const double k2_floor = (5.6e-17 * S) * S / c.W2;
    const bool use_floor = !(k2_floor <= c.k2lo);

    V2Queue Q;
    Q.lane_list = lq + warp * (HBT_V2_LCAP * 32) + lane;
    Q.list_addr = static_cast<unsigned>(__cvta_generic_to_shared(Q.lane_list));
    Q.cur = Q.list_addr;
    Q.wq = wq + warp * HBT_V2_QCAP;
    Q.qcount = 0;
    Q.kept = 0;
    V2Counters n = {0, 0, 0, 0};
    unsigned cntKT = 0, cntRS = 0;
    if (!warp_culled) {
        if (use_floor) {
            if (diag) v2_tile_loop<MIXED, true, true, STATS>(g, c, acc, closed, dv, si, sj, sjt, si_o, sj_o, nj, i0, j0, lane, warp, k2_floor, psi_ref, Q, n, cntKT, cntRS);
            else v2_tile_loop<MIXED, false, true, STATS>(g, c, acc, closed, dv, si, sj, sjt, si_o, sj_o, nj, i0, j0, lane, warp, k2_floor, psi_ref, Q, n, cntKT, cntRS);
        } else {
            if (diag) v2_tile_loop<MIXED, true, false, STATS>(g, c, acc, closed, dv, si, sj, sjt, si_o, sj_o, nj, i0, j0, lane, warp, k2_floor, psi_ref, Q, n, cntKT, cntRS);
            else v2_tile_loop<MIXED, false, false, STATS>(g, c, acc, closed, dv, si, sj, sjt, si_o, sj_o, nj, i0, j0, lane, warp, k2_floor, psi_ref, Q, n, cntKT, cntRS);
        }
    }

    // ---- merge counters: pairs dropped by the prefilter passed K_T (cntKT) minus those queued;
    // those dropped at q_side passed q_out as well (cntRS); queued pairs are counted by the drain
    const unsigned nE = warp_sum(n.nE);
    unsigned long long *stage = acc.stage + (MIXED ? 6 : 0);
    if (STATS) {
        unsigned nB = warp_sum(cntKT + n.nB), nC = warp_sum(cntRS + n.nC);
        const unsigned nD = warp_sum(n.nD);
        nB -= Q.kept;
        if (lane == 0) {
            if (nB) atomicAdd(&stage[1], static_cast<unsigned long long>(nB));
            if (nC) atomicAdd(&stage[2], static_cast<unsigned long long>(nC));
            if (nD) atomicAdd(&stage[3], static_cast<unsigned long long>(nD));
        }
    }
    if (lane == 0 && nE) atomicAdd(&stage[4], static_cast<unsigned long long>(nE));
}

// ---- host-side launch helpers ------------------------------------------------------------
inline size_t hbt_v2_smem_bytes(bool mixed) {
    const int NC = mixed ? 4 : 8;
    return sizeof(double) * (NC * HBT_V2_TILE_I + NC * HBT_V2_TILE_J + HBT_V2_TILE_J + 2 * HBT_V2_WARPS)
           + sizeof(unsigned) * (HBT_V2_WARPS * HBT_V2_LCAP * 32 + HBT_V2_WARPS * HBT_V2_QCAP + HBT_V2_TILE_I + HBT_V2_TILE_J + 8);
}

inline int hbt_v2_configure() {
    cudaError_t e = cudaSuccess;
    const int sm0 = static_cast<int>(hbt_v2_smem_bytes(false)), sm1 = static_cast<int>(hbt_v2_smem_bytes(true));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(hbt_pairs_v2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm0);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(hbt_pairs_v2<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm0);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(hbt_pairs_v2<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(hbt_pairs_v2<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm1);
    return e == cudaSuccess ? HBT_OK : HBT_ERR_CUDA;
}

// d_p: stats = true: the list in the reference's order; stats = false: the Morton-sorted list,
// with orig (reference-order index of each sorted particle) and the tile boxes
inline int hbt_v2_launch_same(cudaStream_t st, const double *d_p, long long n, const HbtGrid &g, const V2Const &c, const V2Dev *d_dv,
                              const HbtAccum &acc, double psi_ref, unsigned long long npairs,
                              const unsigned char *closed, bool stats, const unsigned *orig, const HbtBBox *bbox) {
    const long long T = (n + HBT_V2_TILE_I - 1) / HBT_V2_TILE_I;
    const long long blocks = T * (T + 1) / 2 * HBT_V2_SUB;
    if (blocks > 0x7fffffffLL) return HBT_ERR_INVALID;
    const unsigned nb = static_cast<unsigned>(blocks);
    if (stats)
        hbt_pairs_v2<false, true><<<nb, 32 * HBT_V2_WARPS, hbt_v2_smem_bytes(false), st>>>(
            d_p, d_p, n, nullptr, g, c, d_dv, acc, psi_ref, npairs, closed, nullptr, nullptr);
    else
        hbt_pairs_v2<false, false><<<nb, 32 * HBT_V2_WARPS, hbt_v2_smem_bytes(false), st>>>(
            d_p, d_p, n, nullptr, g, c, d_dv, acc, psi_ref, npairs, closed, orig, bbox);
    return HBT_OK;
}

inline int hbt_v2_launch_mixed(cudaStream_t st, const double *d_p1, const double *d_p2, const HbtMixSeg *d_seg,
                               size_t nseg, long long nblocks, const HbtGrid &g, const V2Const &c, const V2Dev *d_dv, const HbtAccum &acc,
                               double psi_ref, unsigned long long npairs, const unsigned char *closed, bool stats) {
    const unsigned nb = static_cast<unsigned>(nblocks);
    if (stats)
        hbt_pairs_v2<true, true><<<nb, 32 * HBT_V2_WARPS, hbt_v2_smem_bytes(true), st>>>(
            d_p1, d_p2, static_cast<long long>(nseg), d_seg, g, c, d_dv, acc, psi_ref, npairs, closed, nullptr, nullptr);
    else
        hbt_pairs_v2<true, false><<<nb, 32 * HBT_V2_WARPS, hbt_v2_smem_bytes(true), st>>>(
            d_p1, d_p2, static_cast<long long>(nseg), d_seg, g, c, d_dv, acc, psi_ref, npairs, closed, nullptr, nullptr);
    return HBT_OK;
}

#endif  // HBT_KERNELS_V2_CUH_
