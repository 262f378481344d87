// v2 pair kernels (sm_100a): prefilter + warp-level compaction + guarded fast path.
//
// Why: the literal chain (v1) spends ~150 FP64-pipe instructions per pair and runs 2/3 of
// its lanes idle (ncu: 10.5 of 32 threads active per instruction), because 93 % of the
// pairs leave the chain early at different stages.  v2 splits the work in two:
//
//  1. PREFILTER, every pair, no divergence, ~14 FP64-pipe instructions.  Uses only
//     (px, py, pT^2) of the two particles:
//        s = p_i + p_j           k2 = rn(rn(sx^2) + rn(sy^2))  ( = 4 K_perp_sq, bit-exact: the
//                                      reference's 0.5 factors are exact power-of-two scalings)
//        d = pT_i^2 - pT_j^2     ( = q.s = 2 K_perp q_out )
//        x = pxj*pyi - pxi*pyj   ( 2x = q x s = 2 K_perp q_side )
//     K_T cut: exact compare of k2.  q_out / q_side windows: d^2 and x^2 against W^2 k2 with a
//     relative band of ~3e-6 (high-word integer compare); a pair is dropped only when it
//     CERTAINLY fails; anything inside a band goes on.  ~92 % of the pairs end here.
//  2. DRAIN.  Survivors are pushed as (i,j) into a per-warp shared-memory queue; whenever 32
//     are queued the warp processes them with all lanes busy: K_T bin by exact thresholds in
//     k2 space (no sqrt/divide), q_out/q_side/q_long from FMA arithmetic and rsqrt (a few ulp
//     from the reference's values), each compared against the window edges and bin edges with
//     a guard band that bounds the distance to the reference's own rounding.  Inside the guard
//     (probability ~1e-12 per pair) the pair is re-evaluated with the literal chain of
//     hbt_pair.cuh, so every bin decision equals the reference's.  Accepted pairs add
//     cos(q.dx/hbarc) and the q sums with global atomics (histograms are L2 resident).
//
// The accumulated q_out/q_side/q_long are the fast-path values: within ~1e-15 relative of the
// reference's, far inside the 1e-10 tolerance of the sums.
#ifndef HBT_KERNELS_V2_CUH_
#define HBT_KERNELS_V2_CUH_

#include "hbt_kernels_v1.cuh"

#define HBT_V2_WARPS 4
#define HBT_V2_IPL 2  // list-1 particles per lane
#define HBT_V2_TILE_I (32 * HBT_V2_IPL * HBT_V2_WARPS)
#define HBT_V2_TILE_J 256
#define HBT_V2_QCAP 96

// constants of the fast path, derived on the host from HbtGrid
struct V2Const {
    double k2lo, k2hi;        // 4*KT_min_sq, 4*KT_max_sq (exact scalings)
    double kt4[HBT_MAX_KT];   // 4*kt_thr_sq[k]
    double W2;                // max(q_lo^2, q_hi^2)
    double W2q;               // W2 / 4
    double g_abs;             // absolute part of the guard band
    double inv_dq;
    int symmetric;            // |q_lo| == |q_hi| up to 2^-24 relative
};

__host__ inline V2Const hbt_v2_consts(const HbtGrid &g) {
    V2Const c;
    c.k2lo = 4.0 * g.KT_min_sq;
    c.k2hi = 4.0 * g.KT_max_sq;
    for (int k = 0; k < HBT_MAX_KT; k++) c.kt4[k] = 4.0 * g.kt_thr_sq[k];
    const double a = g.q_lo * g.q_lo, b = g.q_hi * g.q_hi;
    c.W2 = a > b ? a : b;
    c.W2q = 0.25 * c.W2;
    const double m = fabs(g.q_lo) > fabs(g.q_hi) ? fabs(g.q_lo) : fabs(g.q_hi);
    c.g_abs = m * 5.7e-14;  // 2^-44
    c.inv_dq = g.inv_dq;
    c.symmetric = (g.q_lo < 0.0 && g.q_hi > 0.0 && fabs(a - b) <= 5.9e-8 * c.W2) ? 1 : 0;
    return c;
}

// v2 handles the 3-D histograms with a window that is symmetric about zero; everything else
// (q_inv mode, one-sided windows) runs on the literal v1 kernels
__host__ inline bool hbt_v2_supported(const HbtGrid &g) {
    return !g.qinv && hbt_v2_consts(g).symmetric;
}

// outcome of a guarded comparison of a fast-path q against the window and the bin grid
enum : int { Q_REJECT = 0, Q_OK = 1, Q_UNSURE = 2 };

__device__ __forceinline__ int classify_q(const HbtGrid &g, const V2Const &c, double q, double guard, int &idx) {
    if (!(q >= g.q_lo - guard && q <= g.q_hi + guard)) return Q_REJECT;  // NaN is rejected
    if (q < g.q_lo + guard || q > g.q_hi - guard) return Q_UNSURE;
    const double u = (q - g.q_base) * c.inv_dq;
    const double fl = floor(u);
    const double fr = u - fl;
    const double gu = guard * c.inv_dq;
    if (fr < gu || fr > 1.0 - gu) return Q_UNSURE;
    idx = __double2int_rz(fl);
    return (idx >= 0 && idx < g.nq) ? Q_OK : Q_UNSURE;
}

// global-memory copy of everything the non-inlined device functions need (passing the
// kernel's by-value parameter structs by reference would copy them to each thread's stack)
struct V2Dev {
    HbtGrid g;
    V2Const c;
    HbtAccum acc;
};

struct V2Counters {
    unsigned nB, nC, nD, nE, nAcc;
};

// the literal chain for a pair the fast path could not decide; counts and accumulates
template <bool MIXED>
__device__ __noinline__ void v2_slow_pair(const V2Dev *__restrict__ dv, const double *a, const double *b,
                                          double psi_ref, V2Counters &n, unsigned *s_slab) {
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
    n.nAcc++;
    atomicAdd(&s_slab[pb.slab], 1u);
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

template <bool MIXED, int NC>
__device__ __noinline__ void v2_drain_pair(const V2Dev *__restrict__ dv, const double (*si)[HBT_V2_TILE_I],
                                           const double (*sj)[HBT_V2_TILE_J], int il, int jl, double psi_ref,
                                           V2Counters &n, unsigned *s_slab) {
    const HbtGrid &g = dv->g;
    const V2Const &c = dv->c;
    const HbtAccum &acc = dv->acc;
    const double ax = si[0][il], ay = si[1][il], bx = sj[0][jl], by = sj[1][jl];
    const double sx = __dadd_rn(ax, bx), sy = __dadd_rn(ay, by);
    const double k2 = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
    if (!(k2 >= c.k2lo && k2 <= c.k2hi)) return;  // only pairs routed here by the tiny-k2 floor
    bool unsure = false;
    int iK = 0;
    for (int k = 1; k < g.nKT; k++) iK += (k2 >= c.kt4[k]) ? 1 : 0;

    const double qx = ax - bx, qy = ay - by;
    const double d = fma(qx, sx, qy * sy);   // 2 K_perp q_out
    const double e = fma(qy, sx, -(qx * sy));  // 2 K_perp q_side
    const double r = rsqrt(k2);              // 1 / (2 K_perp)
    const double qo = d * r, qs = e * r;
    const double gt = fma(fabs(qx) + fabs(qy), 1.5e-14, c.g_abs);  // 2^-46 relative + absolute
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
            az = si[2][il]; aE = si[3][il]; bz = sj[2][jl]; bE = sj[3][jl];
            const double qz = az - bz;
            if (g.boost) {
                // q_long = gamma (q_z - beta q_E) = (K_E q_z - K_z q_E) / Mt, src :383-390
                const double qE = aE - bE, sz = az + bz, sE = aE + bE;
                const double m2 = (sE - sz) * (sE + sz);  // 4 Mt^2 without cancellation
                const double r2 = rsqrt(m2);
                const double t1 = sE * qz, t2 = sz * qE;
                ql = (t1 - t2) * r2;
                const double ch = sE * r2;  // cosh of the pair rapidity: error amplification
                const double gl = fma((fabs(t1) + fabs(t2)) * r2 * fma(2.0 * ch, ch, 1.0), 2.9e-14, c.g_abs);
                const int cl = (m2 > 0.0) ? classify_q(g, c, ql, gl, il_) : Q_UNSURE;
                if (cl == Q_UNSURE) unsure = true;
                if (cl == Q_OK) stage = 4;
            } else {
                // q_long = q_z exactly as the reference has it: use its own comparisons
                ql = qz;
                if (in_window(ql, g.q_lo, g.q_hi, MIXED)) {
                    il_ = __double2int_rz(__ddiv_rn(__dsub_rn(ql, g.q_base), g.dq));
                    if (il_ < g.nq) stage = 4;
                }
            }
        }
    }
    int slab = iK;
    if (!unsure && stage == 4 && g.az) {
        const double Kx = 0.5 * sx, Ky = 0.5 * sy;
        double dphi = __dsub_rn(atan2(Ky, Kx), psi_ref);
        while (dphi < 0.) dphi = __dadd_rn(dphi, g.two_pi);
        while (dphi > g.two_pi) dphi = __dsub_rn(dphi, g.two_pi);
        const double u = __ddiv_rn(dphi, g.dKphi);
        const int iphi = __double2int_rz(u);
        if (fabs(u - rint(u)) < 1e-9) unsure = true;  // literal path defers it to the host
        else if (iphi < 0 || iphi >= g.nKphi) stage = 5;  // counted through q_long, then dropped
        else slab = iK * g.nKphi + iphi;
    }
    if (unsure) {
        double a8[8], b8[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a8[k] = k < NC ? si[k < NC ? k : 0][il] : 0.0;
            b8[k] = k < NC ? sj[k < NC ? k : 0][jl] : 0.0;
        }
        v2_slow_pair<MIXED>(dv, a8, b8, psi_ref, n, s_slab);
        return;
    }
    n.nB++;
    if (stage >= 2) n.nC++;
    if (stage >= 3) n.nD++;
    if (stage >= 4) n.nE++;
    if (stage != 4) return;
    n.nAcc++;
    atomicAdd(&s_slab[slab], 1u);
    const long long bin = ((static_cast<long long>(slab) * g.nq + io) * g.nq + is) * g.nq + il_;
    if (MIXED) {
        atomicAdd(&acc.den_count[bin], 1ull);
    } else {
        const double xd = si[4 % NC][il] - sj[4 % NC][jl], yd = si[5 % NC][il] - sj[5 % NC][jl];
        const double zd = si[6 % NC][il] - sj[6 % NC][jl], td = si[7 % NC][il] - sj[7 % NC][jl];
        const double cv = pair_cos(g, qx, qy, az - bz, aE - bE, xd, yd, zd, td);
        atomicAdd(&acc.num_count[bin], 1ull);
        atomicAdd(&acc.sum_qo[bin], qo);
        atomicAdd(&acc.sum_qs[bin], qs);
        atomicAdd(&acc.sum_ql[bin], ql);
        atomicAdd(&acc.num_cos[bin], cv);
    }
}

template <bool MIXED>
__global__ void __launch_bounds__(32 * HBT_V2_WARPS, 4)
hbt_pairs_v2(const double *__restrict__ p1, const double *__restrict__ p2, long long n_same,
             const HbtMixSeg *__restrict__ segs, const HbtGrid g, const V2Const c,
             const V2Dev *__restrict__ dv, const HbtAccum acc, const double psi_ref,
             const unsigned long long total_pairs) {
    constexpr int NC = MIXED ? 4 : 8;
    constexpr int TI = HBT_V2_TILE_I, TJ = HBT_V2_TILE_J, NT = 32 * HBT_V2_WARPS;
    __shared__ double si[NC][TI];
    __shared__ double sj[NC][TJ];
    __shared__ double sjt[TJ];  // pT^2 of the list-2 tile
    __shared__ unsigned queue[HBT_V2_WARPS][HBT_V2_QCAP];
    __shared__ double s_max[2 * HBT_V2_WARPS];
    __shared__ unsigned s_stage[6];
    extern __shared__ __align__(16) unsigned char dyn[];
    unsigned *s_slab = reinterpret_cast<unsigned *>(dyn);

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
        // square tiling of the upper triangle in units of TJ (= TI) particles
        const long long T = (n_same + TI - 1) / TI;
        int ti, tj;
        tri_decode(blockIdx.x, T, ti, tj);
        i0 = static_cast<long long>(ti) * TI;
        j0 = static_cast<long long>(tj) * TJ;
        ni = static_cast<int>(min(static_cast<long long>(TI), n_same - i0));
        nj = static_cast<int>(min(static_cast<long long>(TJ), n_same - j0));
        diag = (ti == tj);
    }
    if (blockIdx.x == 0 && t == 0) atomicAdd(&acc.stage[MIXED ? 6 : 0], total_pairs);

    for (int k = t; k < g.nslab; k += NT) s_slab[k] = 0;
    if (t < 6) s_stage[t] = 0;

    // ---- stage both tiles (SoA).  Rows beyond the tile end are NaN: they fail the K_T cut.
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double tmax = 0.0;
    for (int k = t; k < TI; k += NT) {
        if (k < ni) {
            const double2 *src = reinterpret_cast<const double2 *>(p1 + 8 * (i0 + k));
            const double2 v0 = src[0], v1 = src[1];
            si[0][k] = v0.x; si[1][k] = v0.y; si[2][k] = v1.x; si[3][k] = v1.y;
            tmax = fmax(tmax, fma(v0.x, v0.x, v0.y * v0.y));
            if (!MIXED) {
                const double2 v2 = src[2], v3 = src[3];
                si[4 % NC][k] = v2.x; si[5 % NC][k] = v2.y; si[6 % NC][k] = v3.x; si[7 % NC][k] = v3.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < NC; q++) si[q][k] = nan;
        }
    }
    double tmaxj = 0.0;
    for (int k = t; k < TJ; k += NT) {
        if (k < nj) {
            const double2 *src = reinterpret_cast<const double2 *>(p2 + 8 * (j0 + k));
            const double2 v0 = src[0], v1 = src[1];
            double x = v0.x, y = v0.y;
            if (MIXED) {  // rotation of the partner event, src/HBT_correlation.cpp:522-523
                x = __dsub_rn(__dmul_rn(v0.x, rc), __dmul_rn(v0.y, rs));
                y = __dadd_rn(__dmul_rn(v0.x, rs), __dmul_rn(v0.y, rc));
            }
            sj[0][k] = x; sj[1][k] = y; sj[2][k] = v1.x; sj[3][k] = v1.y;
            const double pt2 = fma(x, x, y * y);
            sjt[k] = pt2;
            tmaxj = fmax(tmaxj, pt2);
            if (!MIXED) {
                const double2 v2 = src[2], v3 = src[3];
                sj[4 % NC][k] = v2.x; sj[5 % NC][k] = v2.y; sj[6 % NC][k] = v3.x; sj[7 % NC][k] = v3.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < NC; q++) sj[q][k] = nan;
            sjt[k] = nan;
        }
    }
    // S = max pT^2 (list 1) + max pT^2 (list 2) bounds the rounding error of d and x
    double m1 = tmax, m2 = tmaxj;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    if (lane == 0) { s_max[warp] = m1; s_max[HBT_V2_WARPS + warp] = m2; }
    __syncthreads();  // also publishes the tiles
    double S1 = 0.0, S2 = 0.0;
#pragma unroll
    for (int w = 0; w < HBT_V2_WARPS; w++) { S1 = fmax(S1, s_max[w]); S2 = fmax(S2, s_max[HBT_V2_WARPS + w]); }
    const double S = S1 + S2;
    // Below k2_floor the relative error of d (|err| <= 2^-50 S) against the window edge
    // W*sqrt(k2) could exceed the 2^-21 band of the high-word compare: such pairs are not
    // prefiltered.  (k2_floor = 2^-54 S^2 / W2; far below 4*KT_min_sq unless KT_min ~ 0.)
    const double k2_floor = (5.6e-17 * S) * S / c.W2;
    const bool use_floor = k2_floor > c.k2lo;

    // ---- prefilter loop ---------------------------------------------------------------
    const int ia = warp * 64 + lane, ib = ia + 32;
    const double axa = si[0][ia], aya = si[1][ia], axb = si[0][ib], ayb = si[1][ib];
    const double ata = fma(axa, axa, aya * aya), atb = fma(axb, axb, ayb * ayb);
    V2Counters n = {0, 0, 0, 0, 0};  // drain-side counters (may live in local memory)
    unsigned preB = 0, preC = 0;      // prefilter-side counters (registers)
    unsigned *q = queue[warp];
    int qcount = 0;
    const unsigned lt_mask = (1u << lane) - 1u;

    auto prefilter = [&](double ax, double ay, double at, double bx, double by, double bt, bool valid) -> bool {
        const double sx = __dadd_rn(ax, bx), sy = __dadd_rn(ay, by);
        const double k2 = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
        const bool kt = valid && (k2 >= c.k2lo) && (k2 <= c.k2hi);
        const double d = at - bt;
        const double x = fma(bx, ay, -(ax * by));
        const double d2 = d * d, x2 = x * x;
        const double w = c.W2 * k2, wq = c.W2q * k2;
        const int dd = __double2hiint(d2) - __double2hiint(w);
        const int dx = __double2hiint(x2) - __double2hiint(wq);
        const bool fail_o = dd > 1, pass_o = dd < -1, fail_s = dx > 1;
        const bool tiny = use_floor && (k2 < k2_floor);
        const bool rej_o = kt && !tiny && fail_o;
        const bool rej_s = kt && !tiny && pass_o && fail_s;
        preB += (rej_o || rej_s) ? 1u : 0u;
        preC += rej_s ? 1u : 0u;
        return kt && !(rej_o || rej_s);
    };

    auto push = [&](bool keep, int il, int jl) {
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            if (keep) q[qcount + __popc(m & lt_mask)] = (static_cast<unsigned>(il) << 16) | static_cast<unsigned>(jl);
            qcount += __popc(m);
        }
    };

    auto drain = [&](int count) {  // processes the newest `count` (<= 32) entries
        __syncwarp();
        const int base = qcount - count;
        if (lane < count) {
            const unsigned e = q[base + lane];
            v2_drain_pair<MIXED, NC>(dv, si, sj, static_cast<int>(e >> 16), static_cast<int>(e & 0xffffu),
                                     psi_ref, n, s_slab);
        }
        qcount = base;
        __syncwarp();
    };

    for (int j = 0; j < nj; j++) {
        const double bx = sj[0][j], by = sj[1][j], bt = sjt[j];
        const bool va = !diag || (j > ia), vb = !diag || (j > ib);
        const bool ka = prefilter(axa, aya, ata, bx, by, bt, va);
        const bool kb = prefilter(axb, ayb, atb, bx, by, bt, vb);
        push(ka, ia, j);
        push(kb, ib, j);
        while (qcount >= 32) drain(32);
    }
    if (qcount > 0) drain(qcount);

    // ---- merge counters ---------------------------------------------------------------
    const unsigned nB = warp_sum(n.nB + preB), nC = warp_sum(n.nC + preC), nD = warp_sum(n.nD), nE = warp_sum(n.nE),
                   nA = warp_sum(n.nAcc);
    if (lane == 0) {
        atomicAdd(&s_stage[1], nB); atomicAdd(&s_stage[2], nC); atomicAdd(&s_stage[3], nD);
        atomicAdd(&s_stage[4], nE); atomicAdd(&s_stage[5], nA);
    }
    __syncthreads();
    unsigned long long *stage = acc.stage + (MIXED ? 6 : 0);
    if (t >= 1 && t < 6 && s_stage[t]) atomicAdd(&stage[t], static_cast<unsigned long long>(s_stage[t]));
    unsigned long long *npairs = MIXED ? acc.npairs_den : acc.npairs_num;
    for (int k = t; k < g.nslab; k += NT)
        if (s_slab[k]) atomicAdd(&npairs[k], static_cast<unsigned long long>(s_slab[k]));
}

// ---- host-side launch helpers ------------------------------------------------------------

inline int hbt_v2_launch_same(cudaStream_t st, const double *d_p, long long n, const HbtGrid &g, const V2Const &c,
                              const V2Dev *d_dv, const HbtAccum &acc, double psi_ref, unsigned long long npairs) {
    const long long T = (n + HBT_V2_TILE_I - 1) / HBT_V2_TILE_I;
    const long long blocks = T * (T + 1) / 2;
    if (blocks > 0x7fffffffLL) return HBT_ERR_INVALID;
    hbt_pairs_v2<false><<<static_cast<unsigned>(blocks), 32 * HBT_V2_WARPS, ((g.nslab + 3) & ~3) * 4, st>>>(
        d_p, d_p, n, nullptr, g, c, d_dv, acc, psi_ref, npairs);
    return HBT_OK;
}

inline int hbt_v2_launch_mixed(cudaStream_t st, const double *d_p1, const double *d_p2, const HbtMixSeg *d_seg,
                               size_t nseg, long long nblocks, const HbtGrid &g, const V2Const &c, const V2Dev *d_dv,
                               const HbtAccum &acc, double psi_ref, unsigned long long npairs) {
    hbt_pairs_v2<true><<<static_cast<unsigned>(nblocks), 32 * HBT_V2_WARPS, ((g.nslab + 3) & ~3) * 4, st>>>(
        d_p1, d_p2, static_cast<long long>(nseg), d_seg, g, c, d_dv, acc, psi_ref, npairs);
    return HBT_OK;
}

#endif  // HBT_KERNELS_V2_CUH_
