// v3 pair kernels (sm_100a): the v2 pipeline (prefilter -> per-lane survivor lists -> guarded
// fast-path drain, see hbt_kernels_v2.cuh) on a different work decomposition.
//
// One CTA = one warp, persistent: warps pop work UNITS = (128 list-1 particles) x (TJ list-2
// particles) from a global counter until the list is empty.  Why: after the same-event list is
// sorted along a Morton curve, ~80 % of the units can be discarded from their bounding boxes
// (hbt_cull_units builds the list of the others on the device); with several warps per CTA
// sharing a list-2 tile the discarded warps sat idle while the CTA kept its shared memory (ncu:
// 9.9 active warps/SM), and with a static split of the rows the dense-core rows made a long tail
// (6.4 active warps/SM).  Dynamic units of equal size keep every resident warp busy.  No
// __syncthreads anywhere; the survivor queue is flushed once per unit.
//
// Production prefilter in packed FP32 (Blackwell FFMA2/FADD2/FMUL2, two list-1 particles per
// instruction): k2, d = pT_i^2 - pT_j^2 and x = p_j x p_i are evaluated in float from float
// copies of (px, py, pT^2); max(d^2/4, x^2) is compared with (W^2/4)(1 + margin) k2, the margin
// bounding every rounding on the way (inputs, sums, products) from the largest pT^2 of the two
// tiles; the K_T cut is tested with its own margin.  Only pairs that
// CERTAINLY fail are dropped; the drain repeats the K_T cut exactly in FP64, so FP64 still
// decides every cut and every bin edge.  Instrumented runs keep the FP64 prefilter, whose K_T
// test is exact and which counts the stage populations.
#ifndef HBT_KERNELS_V3_CUH_
#define HBT_KERNELS_V3_CUH_

#include <type_traits>
#include <vector>

#include "hbt_kernels_v2.cuh"

#ifndef HBT_V3_SUB_SAME
#define HBT_V3_SUB_SAME 64   // list-1 particles per warp, same-event (2 per lane: more resident warps, finer culling)
#endif
#ifndef HBT_V3_SUB_MIXED
#define HBT_V3_SUB_MIXED 128  // list-1 particles per warp, mixed-event (4 per lane)
#endif
#ifndef HBT_V3_TJ_SAME
#define HBT_V3_TJ_SAME 64    // list-2 tile, same-event (finer culling)
#endif
#ifndef HBT_V3_TJ_MIXED
#define HBT_V3_TJ_MIXED 128  // list-2 tile, mixed-event (no culling: fewer partial drains)
#endif
#ifndef HBT_V3_WARPS_PER_SM
#define HBT_V3_WARPS_PER_SM 18
#endif
#ifndef HBT_V3_WARPS_PER_SM_QINV
#define HBT_V3_WARPS_PER_SM_QINV 16  // q_inv mode keeps two more packed float operands per list-1 pair: 128 registers
#endif
#ifndef HBT_DBG_RED
#define HBT_DBG_RED 0  // control experiments (profiles/r02_controls.txt); never set in the shipped library
#endif
// unit encoding (row << 16 | tile) of the culled list, gridDim.y: fewer than 65536 rows and tiles
#define HBT_V3_MAX_SORTED (65536ll * (HBT_V3_TJ_SAME < HBT_V3_SUB_SAME ? HBT_V3_TJ_SAME : HBT_V3_SUB_SAME) - 64)

// Units of the sorted same-event list that can hold an accepted pair: row a = particles
// [64a, 64a+64), tile t = particles [64t, 64t+64), t >= a (upper triangle incl. the
// diagonal tiles).  work[1] counts them; work[0] is the pop counter of the pair kernel.
// row_end (multi-batch lists, else null): row_end[a] = first tile past the batch that row a belongs to.
__global__ void hbt_cull_units(const HbtBBox *__restrict__ bbox, long long n, double W2, double k2lo, double k2hi,
                               unsigned *__restrict__ units, unsigned *__restrict__ work,
                               const unsigned *__restrict__ row_end = nullptr) {
    constexpr int RB = HBT_V3_SUB_SAME / HBT_BBOX_TILE, TB = HBT_V3_TJ_SAME / HBT_BBOX_TILE;
    const long long nb = (n + HBT_BBOX_TILE - 1) / HBT_BBOX_TILE;
    const int a = blockIdx.y;
    const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    long long ntj = (n + HBT_V3_TJ_SAME - 1) / HBT_V3_TJ_SAME;
    if (row_end) ntj = min(ntj, static_cast<long long>(row_end[a]));
    bool alive = false;
    if (t < ntj && t >= static_cast<long long>(a) * HBT_V3_SUB_SAME / HBT_V3_TJ_SAME) {
        HbtBBox ra = bbox[static_cast<long long>(a) * RB], tb = bbox[t * TB];
        for (int q = 1; q < RB; q++) {
            if (static_cast<long long>(a) * RB + q >= nb) break;
            const HbtBBox b = bbox[static_cast<long long>(a) * RB + q];
            ra.xlo = fmin(ra.xlo, b.xlo); ra.xhi = fmax(ra.xhi, b.xhi); ra.ylo = fmin(ra.ylo, b.ylo); ra.yhi = fmax(ra.yhi, b.yhi);
        }
        for (int q = 1; q < TB; q++) {
            if (t * TB + q >= nb) break;
            const HbtBBox b = bbox[t * TB + q];
            tb.xlo = fmin(tb.xlo, b.xlo); tb.xhi = fmax(tb.xhi, b.xhi); tb.ylo = fmin(tb.ylo, b.ylo); tb.yhi = fmax(tb.yhi, b.yhi);
        }
        alive = !hbt_boxes_culled(ra, tb, W2, k2lo, k2hi);
    }
    const unsigned m = __ballot_sync(0xffffffffu, alive);
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&work[1], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (alive) units[base + __popc(m & ((1u << lane) - 1u))] = (static_cast<unsigned>(a) << 16) | static_cast<unsigned>(t);
    }
}

// Shared memory of one warp, as byte offsets from one base address that the kernel keeps in a
// register (every array is then an immediate offset of an LDS/STS; cvta of separate __shared__
// arrays is recomputed at each use: S2UR + ULEA).
// QINV: invariant_radius_flag = 1 (the 1-D q_inv histograms next to the 3-D ones): no sorted / culled unit list and
// no pT range restriction (both rest on the transverse window alone, which q_inv does not respect), two more float
// arrays (-p_z, -E) for the prefilter's q_inv^2 test.
template <bool MIXED, bool STATS, bool QINV = false>
struct V3Smem {
    static constexpr int NC = MIXED ? 4 : 8;  // doubles kept per particle (mixed-event pairs need no x^mu)
    static constexpr int SUB = MIXED ? HBT_V3_SUB_MIXED : HBT_V3_SUB_SAME, TJ = MIXED ? HBT_V3_TJ_MIXED : HBT_V3_TJ_SAME;
    static constexpr bool SORTED = !MIXED && !STATS && !QINV;
    static constexpr int NF = QINV ? 5 : 3;
    // list-1 sub-tile, SoA, swizzled: lane l keeps particles l, l + 32, ... (s = 0, 1, ...), which sit in the same
    // banks, and a half-warp of a drain round gathers the particles of a few neighbouring lanes: two wavefronts per
    // half-warp for every LDS.64 (ncu: 3.9 per instruction, 1.2 for the list-2 gathers).  Particle il = l + 32 s is kept
    // in slot il ^ (s * SWZ): the s-th particles of neighbouring lanes move SWZ 8-byte banks away from the others.
    // Queue entries carry the slot.
    static constexpr int SWZ = 16 / (SUB / 32);
    static constexpr int SUBP = SUB;                             // stride of the list-1 component arrays
    static constexpr int SI = 0;                                 // double [NC][SUBP]
    // float [NF][TJ] px, py, -pT^2/2 (, -pz, -E) of the list-2 tile (float prefilter).  It sits BEFORE the FP64
    // tile: the prefilter loads particle j+1 while it works on j, and the slot past the last array is then sj[0],
    // which nobody writes during the pair loop
    static constexpr int SJF = SI + 8 * NC * SUBP;
    static constexpr int SJ = SJF + (STATS ? 0 : 4 * NF * TJ);   // double [NC][TJ]   list-2 tile, SoA
    static constexpr int SJT = SJ + 8 * NC * TJ;                 // double [TJ]       pT^2 of the list-2 tile (FP64 prefilter)
    static constexpr int SIO = SJT + (STATS ? 8 * TJ : 0);       // u32    [SUBP]     gather-order index (sorted lists), by slot
    static constexpr int SJO = SIO + (SORTED ? 4 * SUBP : 0);    // u32    [TJ]
    static constexpr int LQ = SJO + (SORTED ? 4 * TJ : 0);       // u16    [LCAP][32] per-lane survivor lists (slot << 8 | position)
    static constexpr int WQ = LQ + 2 * HBT_V2_LCAP * 32;         // u16    [QCAP]     linear warp queue
    static constexpr int BYTES = (WQ + 2 * HBT_V2_QCAP + 15) & ~15;
    static_assert(SUB <= 256 && TJ <= 256, "16-bit queue entries");
};

// ---- guarded fast path for one queued survivor ---------------------------------------------
// In units of bins, u = (q - q_base)/delta_q, the reference's window is [eps, nq - eps] with
// eps = 1e-8/delta_q (src :363-364) and its bin edges are the integers.  The fast path decides
// only when u is farther than gb from every integer, gb >= 2 eps + the bound on
// |fast - reference|: that single test covers the bin edges, both window edges and their
// '>' / '>=' distinction.  Anything closer (a ~1e-5 fraction of the survivors) takes the literal
// chain (v2_slow_pair).  Returns the stage the pair reached: 0 = failed the K_T cut, 1 = passed
// K_T only ... 4 = passed q_long (accepted), or -1 = undecided.
// 1/sqrt(x) for normal x > 0: MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22) and one cubic
// Newton step, e = 1 - x r0^2, r = r0 (1 + e/2 + 3e^2/8): |rel err| < 2^-51, inside the guard.
// 6 instructions against ~20 of rsqrt() with its special-case handling.
__device__ __forceinline__ double v3_rsqrt(double x) {
    double r0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double e = fma(-(x * r0), r0, 1.0);
    return fma(r0 * e, fma(e, 0.375, 0.5), r0);
}

__device__ __forceinline__ float v3_rsqrt_f32(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// cos(x) with a 3-term Cody-Waite reduction by pi/2 and one degree-7 polynomial in r^2 whose
// coefficients (fdlibm's) come from a constant table indexed by the quadrant parity:
// cos r = P0(r^2), sin r = r P1(r^2).  |err| <= 3e-16 for |x| < 1e5 (checked on the host
// against glibc on 2e7 arguments); larger arguments take the library routine.  Feeds a sum
// with a 1e-10 tolerance.  ~35 instructions, no slow-path code in the hot loop.
__constant__ double v3_cos_tab[2][8] = {
    {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
     -1.38888888888741095749e-03, 4.16666666666666019037e-02, -0.5, 1.0},
    {0.0, 1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
     -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01, 1.0}};

// 2/pi and pi/2 in three pieces, as constant-bank operands (an FP64 immediate costs two UMOVs)
__constant__ double v3_cos_red[4] = {6.36619772367581382433e-01, 1.57079632673412561417e+00, 6.07710050650619224932e-11,
                                     2.02226624879595063154e-21};

__device__ __forceinline__ double v3_cos_core(double x);
__device__ __forceinline__ double v3_cos(double x) {
    if (!(fabs(x) < 1.0e5)) return cos(x);
    return v3_cos_core(x);
}
// the branch-free part, |x| < 1e5
__device__ __forceinline__ double v3_cos_core(double x) {
    // n = rint(x 2/pi) and its low bits without FRND / F2I (XU pipe, long latency): x 2/pi + 1.5*2^52 holds n in its
    // low word (|n| < 2^31 for |x| < 1e5); a tie may round the other way than rint() of the rounded product would,
    // which moves r across pi/4 by a rounding error only
    const double magic = __hiloint2double(0x43380000, 0);
    const double tn = fma(x, v3_cos_red[0], magic);
    const double nd = tn - magic;
    double r = fma(-nd, v3_cos_red[1], x);
    r = fma(-nd, v3_cos_red[2], r);
    r = fma(-nd, v3_cos_red[3], r);
    const int nq = __double2loint(tn);
    const double z = r * r;
    const double *t = v3_cos_tab[nq & 1];
    // Estrin's scheme: three dependent FMAs after z instead of seven (the drain is bound by the
    // latency of its dependent FP64 chains, not by their number)
    const double z2 = z * z;
    const double e01 = fma(t[0], z, t[1]), e23 = fma(t[2], z, t[3]), e45 = fma(t[4], z, t[5]), e67 = fma(t[6], z, t[7]);
    const double z4 = z2 * z2;
    const double p = fma(fma(e01, z2, e23), z4, fma(e45, z2, e67));
    const double v = p * ((nq & 1) ? r : 1.0);
    return ((nq + 1) & 2) ? -v : v;
}

// reductions into the (global-memory) histograms: the accumulator pointers arrive as generic
// pointers, for which atomicAdd compiles an address-space test and a shared-memory CAS loop
// next to the RED
__device__ __forceinline__ void red_add_f64(double *p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}
__device__ __forceinline__ void red_inc_u64(unsigned long long *p) {
    asm volatile("red.global.add.u64 [%0], 1;" ::"l"(__cvta_generic_to_global(p)) : "memory");
}

struct V3Bins {
    int io, is, il;
    double qo, qs, ql, qx, qy, qz, qE;
};

template <bool MIXED, bool ORIENT, int TI, int TJ>
__device__ __forceinline__ int v3_fast_bins(const HbtGrid &g, const V2Const &c, unsigned sia, unsigned sja, bool flip,
                                            double &k2, V3Bins &o) {
    const double ax = lds_f64(sia), ay = lds_f64(sia + 8 * TI), bx = lds_f64(sja), by = lds_f64(sja + 8 * TJ);
    const double sx = __dadd_rn(ax, bx), sy = __dadd_rn(ay, by);
    k2 = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
    if (!(k2 >= c.k2lo && k2 <= c.k2hi)) return 0;  // exact K_T cut (the float prefilter only pre-screens it)
    const double qx = ax - bx, qy = ay - by;
    const double d = fma(qx, sx, qy * sy);     // 2 K_perp q_out
    const double e = fma(qy, sx, -(qx * sy));  // 2 K_perp q_side
    const double r = v3_rsqrt(k2);             // 1 / (2 K_perp); k2 >= k2lo > 0 or the NaN falls to the literal chain
    double qo = d * r, qs = e * r;
    if (ORIENT && flip) { qo = -qo; qs = -qs; }
    o.qx = qx; o.qy = qy; o.qo = qo; o.qs = qs;
    // |fast - reference| <= ~8 ulp of (|qx|+|qy|): guard = 2^-46 relative + 2^-44 |window| + 2 eps
    const double gb = fma(fabs(qx) + fabs(qy), c.gq, c.g0);
    const double one_m = 1.0 - gb;
    const unsigned nq = static_cast<unsigned>(g.nq);
    {
        const double u = fma(qo, c.inv_dq, c.ub);
        const int i = __double2int_rd(u);
        if (static_cast<unsigned>(i) >= nq) return (u > -gb && u < c.nq_d + gb) ? -1 : 1;  // outside (NaN too)
        const double fr = u - static_cast<double>(i);
        if (!(fr >= gb && fr <= one_m)) return -1;  // (a NaN lands here too)
        o.io = i;
    }
    {
        const double u = fma(qs, c.inv_dq, c.ub);
        const int i = __double2int_rd(u);
        if (static_cast<unsigned>(i) >= nq) return (u > -gb && u < c.nq_d + gb) ? -1 : 2;
        const double fr = u - static_cast<double>(i);
        if (!(fr >= gb && fr <= one_m)) return -1;
        o.is = i;
    }
    const double az = lds_f64(sia + 16 * TI), aE = lds_f64(sia + 24 * TI), bz = lds_f64(sja + 16 * TJ), bE = lds_f64(sja + 24 * TJ);
    const double qz = az - bz, qE = aE - bE;
    o.qz = qz; o.qE = qE;
    if (g.boost) {
        // q_long = gamma (q_z - beta q_E) = (K_E q_z - K_z q_E) / Mt, src :383-390
        const double sz = az + bz, sE = aE + bE;
        const double m2 = (sE - sz) * (sE + sz);  // 4 Mt^2 without cancellation
        if (!(m2 > 0.0)) return -1;
        const double r2 = v3_rsqrt(m2);
        const double t1 = sE * qz, t2 = sz * qE;
        double ql = (t1 - t2) * r2;
        if (ORIENT && flip) ql = -ql;
        o.ql = ql;
        const double ch = sE * r2;  // cosh of the pair rapidity amplifies the rounding of Mt
        const double gbl = fma((fabs(t1) + fabs(t2)) * r2 * fma(2.0 * ch, ch, 1.0), c.gl, c.g0);
        const double u = fma(ql, c.inv_dq, c.ub);
        const int i = __double2int_rd(u);
        if (static_cast<unsigned>(i) >= nq) return (u > -gbl && u < c.nq_d + gbl) ? -1 : 3;
        const double fr = u - static_cast<double>(i);
        if (!(fr >= gbl && fr <= 1.0 - gbl)) return -1;
        o.il = i;
    } else {
        // q_long = q_z exactly as the reference has it: its own comparisons and index expression
        const double ql = (ORIENT && flip) ? -qz : qz;
        o.ql = ql;
        if (!in_window(ql, g.q_lo, g.q_hi, MIXED)) return 3;
        const int i = __double2int_rz(__ddiv_rn(__dsub_rn(ql, g.q_base), g.dq));
        if (i >= g.nq) return 3;
        o.il = i;
    }
    return 4;
}

__device__ __forceinline__ int v3_kt_bin(const HbtGrid &g, const V2Const &c, double k2);

// Production form of the fast path: the three components are evaluated side by side and decided
// together, so that the q_out/q_side chain (rsqrt of k2) and the q_long chain (rsqrt of 4 Mt^2)
// overlap instead of waiting for each other's early exits (86 % of the queued pairs pass all
// three; the drain is latency-bound at 4-5 warps per scheduler).  Returns 4 (every component
// safely inside a bin), 0 (the K_T cut or some component certainly outside the window) or -1.
__device__ __forceinline__ void v3_classify(const V2Const &c, unsigned nq, double q, double gb, int &i, bool &ok, bool &out) {
    const double u = fma(q, c.inv_dq, c.ub);
    // floor(u) and the distance to the nearest integer without F2I/I2F (XU pipe, long latency):
    // u + 1.5*2^52 holds rint(u) in its low word for |u| < 2^31
    const double magic = __hiloint2double(0x43380000, 0);
    const double t = u + magic;
    const double d = u - (t - magic);  // u - rint(u), in [-0.5, 0.5]
    i = __double2loint(t) - (d < 0.0 ? 1 : 0);
    ok = (static_cast<unsigned>(i) < nq) && (fabs(d) >= gb);  // inside the grid and farther than gb from every bin edge
    out = !(u > -gb && u < c.nq_d + gb);  // certainly outside the window (NaN, huge |u| too: the literal chain rejects them)
}

template <bool MIXED, bool ORIENT, int TI, int TJ>
__device__ __forceinline__ int v3_fast_bins_all(const HbtGrid &g, const V2Const &c, unsigned sia, unsigned sja, bool flip,
                                                double &k2, int &iK, V3Bins &o) {
    const double ax = lds_f64(sia), ay = lds_f64(sia + 8 * TI), bx = lds_f64(sja), by = lds_f64(sja + 8 * TJ);
    const double az = lds_f64(sia + 16 * TI), aE = lds_f64(sia + 24 * TI), bz = lds_f64(sja + 16 * TJ), bE = lds_f64(sja + 24 * TJ);
    const double sx = __dadd_rn(ax, bx), sy = __dadd_rn(ay, by);
    k2 = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
    const bool kt = (k2 >= c.k2lo) && (k2 <= c.k2hi);  // exact K_T cut (the float prefilter only pre-screens it)
    iK = v3_kt_bin(g, c, k2);  // here, not after the decision: its indexed constant loads overlap the chains below
    const double qx = ax - bx, qy = ay - by, qz = az - bz, qE = aE - bE;
    const double d = fma(qx, sx, qy * sy);     // 2 K_perp q_out
    const double e = fma(qy, sx, -(qx * sy));  // 2 K_perp q_side
    const double r = v3_rsqrt(k2);             // 1 / (2 K_perp)
    // sorted lists: the pair may be stored the other way round than the reference takes it; q_out, q_side
    // and q_long change sign with it (a sign folded into the factor: d * (-r) == -(d * r) exactly)
    const double rs = (ORIENT && flip) ? -r : r;
    const double qo = d * rs, qs = e * rs;
    double ql;
    const double gb = fma(fabs(qx) + fabs(qy), c.gq, c.g0);
    const unsigned nq = static_cast<unsigned>(g.nq);
    bool ok_o, out_o, ok_s, out_s, ok_l, out_l;
    if (g.boost) {
        // q_long = gamma (q_z - beta q_E) = (K_E q_z - K_z q_E) / Mt, src :383-390
        const double sz = az + bz, sE = aE + bE;
        const double m2 = (sE - sz) * (sE + sz);  // 4 Mt^2 without cancellation
        const double r2 = v3_rsqrt(m2);
        const double t1 = sE * qz, t2 = sz * qE;
        ql = (t1 - t2) * ((ORIENT && flip) ? -r2 : r2);
        const double ch = sE * r2;  // cosh of the pair rapidity amplifies the rounding of Mt
        const double gbl = fma((fabs(t1) + fabs(t2)) * r2 * fma(2.0 * ch, ch, 1.0), c.gl, c.g0);
        v3_classify(c, nq, ql, gbl, o.il, ok_l, out_l);
        if (!(m2 > 0.0)) { ok_l = false; out_l = false; }  // undecided
    } else {
        // q_long = q_z exactly as the reference has it: its own comparisons and index expression
        ql = (ORIENT && flip) ? -qz : qz;
        o.il = __double2int_rz(__ddiv_rn(__dsub_rn(ql, g.q_base), g.dq));
        ok_l = in_window(ql, g.q_lo, g.q_hi, MIXED) && (o.il < g.nq);
        out_l = !ok_l;
    }
    v3_classify(c, nq, qo, gb, o.io, ok_o, out_o);
    v3_classify(c, nq, qs, gb, o.is, ok_s, out_s);
    o.qx = qx; o.qy = qy; o.qz = qz; o.qE = qE; o.qo = qo; o.qs = qs; o.ql = ql;
    if (!kt || out_o || out_s || out_l) return 0;
    return (ok_o && ok_s && ok_l) ? 4 : -1;
}

// exact K_T bin of k2 = 4 K_perp_sq: float estimate (within one bin: hbt_v2_supported refuses
// grids whose K_T bins are too narrow for that), corrected by one step either way against the
// exact thresholds of int((sqrt(K_perp_sq) - KT_min)/dKT) in k2 space (V2Const::kt4, found by
// bisection on the host)
__device__ __forceinline__ int v3_kt_bin(const HbtGrid &g, const V2Const &c, double k2) {
    // up to four bins (the reference's default): the bin is the number of thresholds at or below k2 (unused ones are
    // +inf) — three compares against constant-bank operands instead of the estimate's F2F / MUFU / F2I and its fix-up
    if (g.nKT <= 4) return (k2 >= c.kt4[1] ? 1 : 0) + (k2 >= c.kt4[2] ? 1 : 0) + (k2 >= c.kt4[3] ? 1 : 0);
    const float k2f = static_cast<float>(k2);
    // (one MUFU: rsqrtf() adds denormal scaling; an estimate only, k2 below 1e-30 lands in bin 0 either way)
    const float kp = 0.5f * k2f * v3_rsqrt_f32(fmaxf(k2f, 1e-30f));
    int iK = static_cast<int>((kp - c.kt_min_f) * c.inv_dkt_f);
    iK = max(0, min(iK, g.nKT - 1));
    if (iK > 0 && k2 < c.kt4[iK]) iK--;
    else if (iK < g.nKT - 1 && k2 >= c.kt4[iK + 1]) iK++;
    return iK;
}

// ---- FP32 decision of a queued mixed-event survivor --------------------------------------------
// A mixed-event pair contributes one count to one bin (src/HBT_correlation.cpp:682): only its K_T
// bin and its three q bins are needed, not the values.  They are evaluated here in binary32 from
// the rounded particle components, together with a bound on |float result - exact value| (u = 2^-24,
// every input rounding and every operation accounted for; X = |ax|+|bx|, Y = |ay|+|by|, S = X+Y,
// Z = |az|+|bz|, E = |aE|+|bE|, r ~ 1/sqrt(k2), r2 ~ 1/sqrt(m2), MUFU.RSQ within 2^-22):
//     |k2_f - k2|       <= u (4 S^2 + 2 k2)
//     |q_out,side_f - q| <= u (6 S^2 r + |q| (6 + 2 S^2 r^2))
//     |q_long_f - q|    <= u (10 W Z + |q| (2 W^2 + 8)),   W = (E + Z) r2
// In bin units (x 1/delta_q) twice that bound plus the guard of the FP64 path and the rounding of
// the float grid constants is the guard band: the pair is decided only if every component is
// farther than its band from every integer (bin edges and both window edges) and k2 is farther
// than its band from the K_T cut and from both edges of its K_T bin.  Otherwise (~1e-3 of the
// survivors) the FP64 path below decides, so every bin index still equals the reference's.
// Returns 1 (accepted: slab and bin set), 0 (certainly outside the window) or -1 (undecided).
// floor(u) and the distance to the nearest integer for |u| < 2^22 (u + 1.5*2^23 holds rint(u) in
// its mantissa); beyond that the band (which grows with |u|) exceeds any distance this returns
__device__ __forceinline__ void v3_classify_f32(float u, float gb, unsigned nq, int &i, bool &ok, bool &out) {
    const float magic = 12582912.f;
    const float t = u + magic;
    const float d = u - (t - magic);  // u - rint(u)
    i = __float_as_int(t) - 0x4B400000 - (d < 0.f ? 1 : 0);
    const bool far = fabsf(d) > gb;   // false for NaN / inf
    const bool in_grid = static_cast<unsigned>(i) < nq;
    ok = far && in_grid;
    out = far && !in_grid;
}

// (the arithmetic on the eight rounded components; hbt_kernels_v4.cuh calls it on tiles that are kept in binary32)
__device__ __forceinline__ int v3_mixed_f32_core(const HbtGrid &g, const V2Const &c, const float ax, const float ay, const float az,
                                                 const float aE, const float bx, const float by, const float bz, const float bE,
                                                 double psi_ref, int &slab, unsigned &bin) {
    const float u8 = 4.76837158203125e-7f;  // 8 * 2^-24
    // transverse plane: k2 = 4 K_perp^2, q_out = d r, q_side = e r
    const float sx = ax + bx, sy = ay + by, qx = ax - bx, qy = ay - by;
    const float S = (fabsf(ax) + fabsf(bx)) + (fabsf(ay) + fabsf(by));
    const float S2 = S * S;
    const float k2 = fmaf(sy, sy, sx * sx);
    const float r = v3_rsqrt_f32(k2);
    // K_T: float estimate of the bin, then k2 against the cut / the bin's own edges with the band
    int iK = static_cast<int>((0.5f * k2 * r - c.kt_min_f) * c.inv_dkt_f);
    iK = max(0, min(iK, g.nKT - 1));
    const float ek = (S2 + k2) * u8;  // 2 x the bound, threshold rounding included
    const bool kt_ok = (k2 - c.ktf[iK] > ek) && (c.ktf[iK + 1] - k2 > ek);
    const float d = fmaf(qx, sx, qy * sy), e = fmaf(qy, sx, -(qx * sy));
    const float qo = d * r, qs = e * r;
    const float A = S2 * r, R = A * r;
    const float c6 = fmaf(2.f, R, 6.f), A6 = 6.f * A;
    const float gbo = fmaf(fmaf(fabsf(qo), c6, A6), c.f_gs, c.f_g0);
    const float gbs = fmaf(fmaf(fabsf(qs), c6, A6), c.f_gs, c.f_g0);
    // longitudinal: q_long = (K_E q_z - K_z q_E) / Mt in the LCMS (src :383-390), or q_z
    const float qz = az - bz;
    const float Z = fabsf(az) + fabsf(bz);
    float ql, gbl;
    if (g.boost) {
        const float sz = az + bz, sE = aE + bE, qE = aE - bE;
        const float m2 = (sE - sz) * (sE + sz);
        const float r2 = v3_rsqrt_f32(m2);  // m2 <= 0: inf / NaN, the band test fails
        const float t1 = sE * qz, t2 = sz * qE;
        ql = (t1 - t2) * r2;
        const float W = ((fabsf(aE) + fabsf(bE)) + Z) * r2;
        gbl = fmaf(fmaf(fabsf(ql), fmaf(2.f * W, W, 8.f), 10.f * W * Z), c.f_gs, c.f_g0);
    } else {
        ql = qz;
        gbl = fmaf(2.f * Z, c.f_gs, c.f_g0);
    }
    const unsigned nq = static_cast<unsigned>(g.nq);
    int io, is, il;
    bool ok_o, out_o, ok_s, out_s, ok_l, out_l;
    v3_classify_f32(fmaf(qo, c.f_inv_dq, c.f_ub), gbo, nq, io, ok_o, out_o);
    v3_classify_f32(fmaf(qs, c.f_inv_dq, c.f_ub), gbs, nq, is, ok_s, out_s);
    v3_classify_f32(fmaf(ql, c.f_inv_dq, c.f_ub), gbl, nq, il, ok_l, out_l);
    // decided means: K_T decided inside, and every component either safely inside a bin or safely outside
    if (!(kt_ok && (ok_o || out_o) && (ok_s || out_s) && (ok_l || out_l))) return -1;
    if (out_o || out_s || out_l) return 0;
    slab = iK;
    if (g.az) {
        // K_phi bin (:392-400) from the float estimate, as in the FP64 path; the angle of the float K
        // is within 2u S r of the exact one: ask for a well-conditioned K and 1e-4 from every edge
        if (!(S * r * c.f_nkphi <= 256.f) || g.nKphi > 256) return -1;
        double de = static_cast<double>(atan2f(sy, sx)) - psi_ref;
        de = de < 0. ? de + g.two_pi : de;
        de = de > g.two_pi ? de - g.two_pi : de;
        const double ue = de * c.inv_dkphi;
        const double fe = ue - floor(ue);
        if (!(fe > 1e-4 && fe < 1.0 - 1e-4 && ue > 0. && ue < static_cast<double>(g.nKphi))) return -1;
        slab = iK * g.nKphi + static_cast<int>(ue);
    }
    bin = ((static_cast<unsigned>(slab) * nq + io) * nq + is) * nq + il;
    return 1;
}

template <int TI, int TJ>
__device__ __forceinline__ int v3_mixed_f32(const HbtGrid &g, const V2Const &c, unsigned sia, unsigned sja, double psi_ref,
                                            int &slab, unsigned &bin) {
    const float ax = static_cast<float>(lds_f64(sia)), ay = static_cast<float>(lds_f64(sia + 8 * TI));
    const float az = static_cast<float>(lds_f64(sia + 16 * TI)), aE = static_cast<float>(lds_f64(sia + 24 * TI));
    const float bx = static_cast<float>(lds_f64(sja)), by = static_cast<float>(lds_f64(sja + 8 * TJ));
    const float bz = static_cast<float>(lds_f64(sja + 16 * TJ)), bE = static_cast<float>(lds_f64(sja + 24 * TJ));
    return v3_mixed_f32_core(g, c, ax, ay, az, aE, bx, by, bz, bE, psi_ref, slab, bin);
}

// ---- q_inv branch of one queued survivor (src :323-356 same event, :585-607 mixed event) -------------------
// Everything that decides is the reference's own binary64 chain: k2 = 4 K_perp_sq exactly as it rounds (the K_T
// cut and, through the host's thresholds, the K_T bin), s = -(q_E^2 - q_x^2 - q_y^2 - q_z^2) with its operation
// order, and the window / bin tests of q_inv = sqrt(s) as comparisons of s with the host's exact thresholds
// (V2Const::qinv_s_lo / qinv_s_hi / qinv_thr).  Accepted pairs go to this lane's replica of the q_inv
// accumulators; sqrt and cos are evaluated only for them.
// k2, iK and the four momentum differences come from the 3-D fast path (v3_fast_bins_all evaluates k2 with the
// reference's operations, the K_T bin from the exact thresholds, and q = p_1 - p_2 as single IEEE subtractions; q_inv
// mode has no sorted lists, so the differences are in the reference's orientation).
template <bool MIXED, int NC, int TI, int TJ>
__device__ __forceinline__ void v3_qinv_pair(const HbtGrid &g, const V2Const &c, const unsigned char *__restrict__ closed,
                                             unsigned sia, unsigned sja, double k2, int iK, double qx, double qy, double qz, double qE) {
    if (!((k2 >= c.k2lo) && (k2 <= c.k2hi))) return;  // :319-321 / :581-583
    const double m2 = __dsub_rn(__dsub_rn(__dsub_rn(__dmul_rn(qE, qE), __dmul_rn(qx, qx)), __dmul_rn(qy, qy)), __dmul_rn(qz, qz));
    const double s = -m2;
    if (!((s >= c.qinv_s_lo) && (s < c.qinv_s_hi))) return;  // q_inv outside (q_lo, q_hi), or NaN (s < 0)
    // bin: float estimate, settled on the exact thresholds
    const int nq = g.nq;
    int iq = static_cast<int>((sqrtf(static_cast<float>(s)) - c.f_qbase) * c.f_inv_dq);
    iq = max(0, min(iq, nq));
    while (iq > 0 && s < c.qinv_thr[iq]) iq--;
    while (iq < nq && s >= c.qinv_thr[iq + 1]) iq++;
    if (iq >= nq) return;  // (:599; the same-event loop would index past its array there)
    if (closed && closed[2 * g.nslab + iK + (MIXED ? g.nKT : 0)]) return;  // 50 x needed_number_of_pairs reached earlier
    const unsigned nb = static_cast<unsigned>(g.nKT * nq);
    const unsigned rep = (blockIdx.x * 32u + (threadIdx.x & 31u)) & static_cast<unsigned>(c.qrep_n - 1);
    const unsigned at = (rep * 2u + (MIXED ? 1u : 0u)) * nb + static_cast<unsigned>(iK * nq + iq);
    red_inc_u64(c.qrep_u64 + at);
    if (!MIXED) {
        const double xd = lds_f64(sia + 8 * TI * (4 % NC)) - lds_f64(sja + 8 * TJ * (4 % NC));
        const double yd = lds_f64(sia + 8 * TI * (5 % NC)) - lds_f64(sja + 8 * TJ * (5 % NC));
        const double zd = lds_f64(sia + 8 * TI * (6 % NC)) - lds_f64(sja + 8 * TJ * (6 % NC));
        const double td = lds_f64(sia + 8 * TI * (7 % NC)) - lds_f64(sja + 8 * TJ * (7 % NC));
        const double cv = v3_cos(g.hbarc_inv * (qE * td - qx * xd - qy * yd - qz * zd));  // :347-350
        double *f = c.qrep_f64 + static_cast<size_t>(rep) * 2u * nb + static_cast<unsigned>(iK * nq + iq);
        red_add_f64(f, __dsqrt_rn(s));  // q_inv as the reference has it
        red_add_f64(f + nb, cv);
    }
}

// Sums the replicas of the q_inv accumulators into the histograms and the per-K_T pair counters, and clears them
// (atomic exchange: another lane's launch may be adding meanwhile).  One block per q_inv bin, one thread per replica.
__global__ void hbt_qinv_fold(const V2Const c, const HbtAccum acc, int nKT, int nq) {
    const unsigned nb = static_cast<unsigned>(nKT * nq);
    const unsigned bin = blockIdx.x;
    unsigned long long cnt = 0, den = 0;
    double sq = 0.0, sc = 0.0;
    for (int r = threadIdx.x; r < c.qrep_n; r += blockDim.x) {
        unsigned long long *u = c.qrep_u64 + static_cast<size_t>(r) * 2u * nb + bin;
        double *f = c.qrep_f64 + static_cast<size_t>(r) * 2u * nb + bin;
        // each of the four taken on its own: another lane's running launch may have counted a pair whose sums have
        // not landed yet; they are picked up by the fold that follows that launch
        if (__ldcg(u)) cnt += atomicExch(u, 0ull);
        if (__ldcg(u + nb)) den += atomicExch(u + nb, 0ull);
        if (__ldcg(f) != 0.0) sq += __longlong_as_double(static_cast<long long>(atomicExch(reinterpret_cast<unsigned long long *>(f), 0ull)));
        if (__ldcg(f + nb) != 0.0) sc += __longlong_as_double(static_cast<long long>(atomicExch(reinterpret_cast<unsigned long long *>(f + nb), 0ull)));
    }
    __shared__ unsigned long long s_u[2][32];
    __shared__ double s_f[2][32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        sc += __shfl_xor_sync(0xffffffffu, sc, o);
    }
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { s_u[0][w] = cnt; s_u[1][w] = den; s_f[0][w] = sq; s_f[1][w] = sc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nw; k++) { cnt += s_u[0][k]; den += s_u[1][k]; sq += s_f[0][k]; sc += s_f[1][k]; }
        const int iK = static_cast<int>(bin) / nq;
        if (cnt) {
            atomicAdd(&acc.qinv_count[bin], cnt);
            atomicAdd(&acc.npairs_num_qinv[iK], cnt);
        }
        if (sq != 0.0) atomicAdd(&acc.qinv_sum[bin], sq);
        if (sc != 0.0) atomicAdd(&acc.qinv_cos[bin], sc);
        if (den) {
            atomicAdd(&acc.qinv_den[bin], den);
            atomicAdd(&acc.npairs_den_qinv[iK], den);
        }
    }
}

// ---- production same-event survivor (sorted lists, no stage counters, no q_inv): v3_fast_bins_all + the
// accumulation of v3_drain_pair in one straight line.  Same arithmetic, same guards; what differs is bookkeeping:
// the three classifications are combined as predicates (no stage code built up by selects), the orientation sign is
// an XOR on the high word of the two reciprocal square roots, and the tile addresses come straight from the two
// halves of the queue entry.
template <int NC, int TI, int TJ, int SI, int SJ, int SIO, int SJO>
__device__ __forceinline__ void v3_same_pair(const HbtGrid &g, const V2Const &c, const HbtAccum &acc,
                                             const unsigned char *__restrict__ closed, const V2Dev *__restrict__ dv,
                                             unsigned sbase, unsigned entry, double psi_ref, V2Counters &n) {
    const unsigned il = entry >> 8, jl = entry & 0xffu;
    const unsigned sia = sbase + SI + 8u * il, sja = sbase + SJ + 8u * jl;
    const bool flip = lds_u32(sbase + SIO + 4u * il) > lds_u32(sbase + SJO + 4u * jl);
    const double ax = lds_f64(sia), ay = lds_f64(sia + 8 * TI), bx = lds_f64(sja), by = lds_f64(sja + 8 * TJ);
    const double az = lds_f64(sia + 16 * TI), aE = lds_f64(sia + 24 * TI), bz = lds_f64(sja + 16 * TJ), bE = lds_f64(sja + 24 * TJ);
    const double sx = __dadd_rn(ax, bx), sy = __dadd_rn(ay, by);
    const double k2 = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
    const bool kt = (k2 >= c.k2lo) && (k2 <= c.k2hi);  // exact K_T cut (the float prefilter only pre-screens it)
    int slab = v3_kt_bin(g, c, k2);
    const double qx = ax - bx, qy = ay - by, qz = az - bz, qE = aE - bE;
    const double d = fma(qx, sx, qy * sy);     // 2 K_perp q_out
    const double e = fma(qy, sx, -(qx * sy));  // 2 K_perp q_side
    const double r = v3_rsqrt(k2);             // 1 / (2 K_perp)
    // the pair may be stored the other way round than the reference takes it: q_out, q_side, q_long change sign with
    // it (a sign folded into the factor: d * (-r) == -(d * r) exactly)
    const int sgn = flip ? static_cast<int>(0x80000000u) : 0;
    const double rs = __hiloint2double(__double2hiint(r) ^ sgn, __double2loint(r));
    const double qo = d * rs, qs = e * rs;
    const double gb = fma(fabs(qx) + fabs(qy), c.gq, c.g0);
    const unsigned nq = static_cast<unsigned>(g.nq);
    double ql;
    int io, is, il_;
    bool ok_o, out_o, ok_s, out_s, ok_l, out_l;
    if (g.boost) {
        // q_long = gamma (q_z - beta q_E) = (K_E q_z - K_z q_E) / Mt, src :383-390
        const double sz = az + bz, sE = aE + bE;
        const double m2 = (sE - sz) * (sE + sz);  // 4 Mt^2 without cancellation
        const double r2 = v3_rsqrt(m2);
        const double t1 = sE * qz, t2 = sz * qE;
        ql = (t1 - t2) * __hiloint2double(__double2hiint(r2) ^ sgn, __double2loint(r2));
        const double ch = sE * r2;  // cosh of the pair rapidity amplifies the rounding of Mt
        const double gbl = fma((fabs(t1) + fabs(t2)) * r2 * fma(2.0 * ch, ch, 1.0), c.gl, c.g0);
        v3_classify(c, nq, ql, gbl, il_, ok_l, out_l);
        if (!(m2 > 0.0)) { ok_l = false; out_l = false; }  // undecided
    } else {
        // q_long = q_z exactly as the reference has it: its own comparisons and index expression
        ql = flip ? -qz : qz;
        il_ = __double2int_rz(__ddiv_rn(__dsub_rn(ql, g.q_base), g.dq));
        ok_l = in_window(ql, g.q_lo, g.q_hi, false) && (il_ < g.nq);
        out_l = !ok_l;
    }
    v3_classify(c, nq, qo, gb, io, ok_o, out_o);
    v3_classify(c, nq, qs, gb, is, ok_s, out_s);
    if (!kt || out_o || out_s || out_l) return;  // the K_T cut or some component certainly outside the window
    bool undecided = !(ok_o && ok_s && ok_l);
    bool dropped = false;  // K_phi out of range: counted through q_long, then dropped (:417-423)
    if (!undecided && g.az) {
        const double Kx = 0.5 * sx, Ky = 0.5 * sy;  // (exact halvings of the reference's sums)
        // K_phi bin (:392-400).  First a float estimate: atan2f (<= 3 ulp) of the float-rounded K gives
        // Delta phi / dK_phi within ~2e-6 n_Kphi/8 of the reference's value; farther than 1e-4 from every
        // integer (bin edges, both ends of the range) the bin is decided.  Otherwise the double-precision
        // expression, and within 1e-9 of an edge the host (glibc atan2).
        double de = static_cast<double>(atan2f(static_cast<float>(Ky), static_cast<float>(Kx))) - psi_ref;
        de = de < 0. ? de + g.two_pi : de;
        de = de > g.two_pi ? de - g.two_pi : de;
        const double ue = de * c.inv_dkphi;
        const double fe = ue - floor(ue);
        if (fe > 1e-4 && fe < 1.0 - 1e-4 && ue > 0. && ue < static_cast<double>(g.nKphi) && g.nKphi <= 256) {
            slab = slab * g.nKphi + static_cast<int>(ue);
        } else {
            double dphi = __dsub_rn(atan2(Ky, Kx), psi_ref);
            while (dphi < 0.) dphi = __dadd_rn(dphi, g.two_pi);
            while (dphi > g.two_pi) dphi = __dsub_rn(dphi, g.two_pi);
            const double u = __ddiv_rn(dphi, g.dKphi);
            const int iphi = __double2int_rz(u);
            if (fabs(u - rint(u)) < 1e-9) undecided = true;  // the literal path hands it to the host
            else if (!(u == u) || iphi < 0 || iphi >= g.nKphi) dropped = true;
            else slab = slab * g.nKphi + iphi;
        }
    }
    if (undecided) {  // literal chain (roles swapped back into the reference's order)
        double a8[8], b8[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double va = lds_f64(sia + 8 * TI * k), vb = lds_f64(sja + 8 * TJ * k);
            a8[k] = flip ? vb : va;
            b8[k] = flip ? va : vb;
        }
        V2Counters tmp = {0, 0, 0, 0};  // keeps n itself out of local memory
        v2_slow_pair<false>(dv, a8, b8, psi_ref, tmp);
        n.nE += tmp.nE;
        return;
    }
    n.nE++;
    if (dropped) return;
    if (closed && closed[slab]) return;  // needed_number_of_pairs reached earlier
    const unsigned bin = ((static_cast<unsigned>(slab) * g.nq + io) * g.nq + is) * g.nq + il_;  // < 2^31 (hbt_create)
    const double xd = lds_f64(sia + 8 * TI * 4) - lds_f64(sja + 8 * TJ * 4);
    const double yd = lds_f64(sia + 8 * TI * 5) - lds_f64(sja + 8 * TJ * 5);
    const double zd = lds_f64(sia + 8 * TI * 6) - lds_f64(sja + 8 * TJ * 6);
    const double td = lds_f64(sia + 8 * TI * 7) - lds_f64(sja + 8 * TJ * 7);
    const double cv = v3_cos(g.hbarc_inv * (qE * td - qx * xd - qy * yd - qz * zd));  // src :431-433
    red_inc_u64(&acc.num_count[bin]);
    red_add_f64(&acc.sum_qo[bin], qo);
    red_add_f64(&acc.sum_qs[bin], qs);
    red_add_f64(&acc.sum_ql[bin], ql);
    red_add_f64(&acc.num_cos[bin], cv);
}

template <bool MIXED, bool STATS, bool QINV = false>
__device__ __forceinline__ void v3_drain_pair(const HbtGrid &g, const V2Const &c, const HbtAccum &acc,
                                              const unsigned char *__restrict__ closed, const V2Dev *__restrict__ dv,
                                              unsigned sbase, unsigned entry, double psi_ref, V2Counters &n) {
    using L = V3Smem<MIXED, STATS, QINV>;
    constexpr int NC = L::NC, TI = L::SUBP, TJ = L::TJ;  // TI: stride of the list-1 component arrays
    constexpr bool ORIENT = L::SORTED;
#if !HBT_DBG_RED && !defined(HBT_V3_NO_LEAN_SAME)
    if constexpr (L::SORTED) {  // production same-event units
        v3_same_pair<NC, TI, TJ, L::SI, L::SJ, L::SIO, L::SJO>(g, c, acc, closed, dv, sbase, entry, psi_ref, n);  // (TI = SUBP)
        return;
    }
#endif
    const unsigned il4 = (entry >> 6) & ~3u, jl4 = (entry & 0xffu) << 2;  // 4 x list-1 / list-2 slot
    const unsigned sia = sbase + L::SI + 2 * il4, sja = sbase + L::SJ + 2 * jl4;
    if (MIXED && !STATS && !QINV && c.f32_mixed) {
        int fslab;
        unsigned fbin;
        const int fs = v3_mixed_f32<TI, TJ>(g, c, sia, sja, psi_ref, fslab, fbin);
        if (fs == 0) return;
        if (fs > 0) {
            n.nE++;
            if (closed && closed[fslab + g.nslab]) return;  // needed_number_of_pairs reached earlier
            red_inc_u64(&acc.den_count[fbin]);
            return;
        }
    }
    const bool flip = ORIENT && (lds_u32(sbase + L::SIO + il4) > lds_u32(sbase + L::SJO + jl4));
    V3Bins b;
    double k2;
    int slab = 0;
    int stage = STATS ? v3_fast_bins<MIXED, ORIENT, TI, TJ>(g, c, sia, sja, flip, k2, b)
                      : v3_fast_bins_all<MIXED, ORIENT, TI, TJ>(g, c, sia, sja, flip, k2, slab, b);
    // q_inv histograms: independent of what the 3-D chain decided (slab = the K_T bin here: K_phi comes later)
    if (QINV) v3_qinv_pair<MIXED, NC, TI, TJ>(g, c, closed, sia, sja, k2, slab, b.qx, b.qy, b.qz, b.qE);
    if (stage == 4) {
        if (STATS) slab = v3_kt_bin(g, c, k2);
        if (g.az) {
            const double Kx = 0.5 * (lds_f64(sia) + lds_f64(sja)), Ky = 0.5 * (lds_f64(sia + 8 * TI) + lds_f64(sja + 8 * TJ));
            // K_phi bin (:392-400).  First a float estimate: atan2f (<= 3 ulp) of the float-rounded K gives
            // Delta phi / dK_phi within ~2e-6 n_Kphi/8 of the reference's value; farther than 1e-4 from every
            // integer (bin edges, both ends of the range) the bin is decided.  Otherwise the double-precision
            // expression, and within 1e-9 of an edge the host (glibc atan2).
            double de = static_cast<double>(atan2f(static_cast<float>(Ky), static_cast<float>(Kx))) - psi_ref;
            de = de < 0. ? de + g.two_pi : de;
            de = de > g.two_pi ? de - g.two_pi : de;
            const double ue = de * c.inv_dkphi;
            const double fe = ue - floor(ue);
            if (!STATS && fe > 1e-4 && fe < 1.0 - 1e-4 && ue > 0. && ue < static_cast<double>(g.nKphi) && g.nKphi <= 256) {
                slab = slab * g.nKphi + static_cast<int>(ue);
            } else {
                double dphi = __dsub_rn(atan2(Ky, Kx), psi_ref);
                while (dphi < 0.) dphi = __dadd_rn(dphi, g.two_pi);
                while (dphi > g.two_pi) dphi = __dsub_rn(dphi, g.two_pi);
                const double u = __ddiv_rn(dphi, g.dKphi);
                const int iphi = __double2int_rz(u);
                if (fabs(u - rint(u)) < 1e-9) stage = -1;  // the literal path hands it to the host
                else if (!(u == u) || iphi < 0 || iphi >= g.nKphi) stage = 5;  // counted through q_long, then dropped (NaN angle too)
                else slab = slab * g.nKphi + iphi;
            }
        }
    }
    if (stage < 0) {  // undecided: literal chain (roles swapped back into the reference's order)
        double a8[8], b8[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double va = k < NC ? lds_f64(sia + 8 * TI * (k < NC ? k : 0)) : 0.0;
            const double vb = k < NC ? lds_f64(sja + 8 * TJ * (k < NC ? k : 0)) : 0.0;
            a8[k] = flip ? vb : va;
            b8[k] = flip ? va : vb;
        }
        V2Counters tmp = {0, 0, 0, 0};  // keeps n itself out of local memory
        v2_slow_pair<MIXED>(dv, a8, b8, psi_ref, tmp);
        n.nB += tmp.nB; n.nC += tmp.nC; n.nD += tmp.nD; n.nE += tmp.nE;
        return;
    }
    if (STATS) {
        if (stage >= 1) n.nB++;
        if (stage >= 2) n.nC++;
        if (stage >= 3) n.nD++;
    }
    if (stage < 4) return;
    n.nE++;
    if (stage != 4) return;
    if (closed && closed[slab + (MIXED ? g.nslab : 0)]) return;  // needed_number_of_pairs reached earlier
    const unsigned bin = ((static_cast<unsigned>(slab) * g.nq + b.io) * g.nq + b.is) * g.nq + b.il;  // < 2^31 (hbt_create)
    if (MIXED) {
#if HBT_DBG_RED == 1
        if (bin == 0x7fffffffu)
#endif
        red_inc_u64(&acc.den_count[bin]);
    } else {
#if HBT_DBG_RED == 6  // control: conflict-free reads of the space-time components (wrong cos, same control flow)
        const unsigned sia2 = sbase + L::SI + 8u * (threadIdx.x & 31), sja2 = sbase + L::SJ + 8u * (threadIdx.x & 31);
#else
        const unsigned sia2 = sia, sja2 = sja;
#endif
        const double xd = lds_f64(sia2 + 8 * TI * (4 % NC)) - lds_f64(sja2 + 8 * TJ * (4 % NC));
        const double yd = lds_f64(sia2 + 8 * TI * (5 % NC)) - lds_f64(sja2 + 8 * TJ * (5 % NC));
        const double zd = lds_f64(sia2 + 8 * TI * (6 % NC)) - lds_f64(sja2 + 8 * TJ * (6 % NC));
        const double td = lds_f64(sia2 + 8 * TI * (7 % NC)) - lds_f64(sja2 + 8 * TJ * (7 % NC));
        const double cv = v3_cos(g.hbarc_inv * (b.qE * td - b.qx * xd - b.qy * yd - b.qz * zd));  // src :431-433
#if HBT_DBG_RED == 1    // control: no reductions; the compiler sinks the cos chain into the dead branch, so this
                        // also removes ~1/4 of the drain's arithmetic (see HBT_DBG_RED == 7)
        if (bin == 0x7fffffffu) {
#elif HBT_DBG_RED == 7  // control: no reductions, the four sums still evaluated (their bits feed a counter)
        n.nE += static_cast<unsigned>(__double2hiint(cv) ^ __double2hiint(b.qo) ^ __double2hiint(b.qs) ^ __double2hiint(b.ql)) >> 31;
        if (bin == 0x7fffffffu) {
#elif HBT_DBG_RED == 2  // control: all reductions into 1024 bins
        const unsigned bin_ = bin;
        {
            const unsigned bin = bin_ & 1023u;
#else
        {
#endif
            red_inc_u64(&acc.num_count[bin]);
#if HBT_DBG_RED == 3    // control: one reduction per accepted pair instead of five
            if (bin == 0x7fffffffu) {
#else
            {
#endif
                red_add_f64(&acc.sum_qo[bin], b.qo);
                red_add_f64(&acc.sum_qs[bin], b.qs);
                red_add_f64(&acc.sum_ql[bin], b.ql);
                red_add_f64(&acc.num_cos[bin], cv);
            }
        }
    }
}

// One work unit (a list-1 sub-tile against one list-2 tile): stage the tiles, prefilter every
// pair, compact the survivors and drain them.  The warp's survivor queue is empty on entry and on
// return.  smem/sbase: this warp's shared memory (layout V3Smem<MIXED, STATS>).
template <bool MIXED, bool STATS, bool QINV = false>
__device__ __forceinline__ void v3_run_unit(unsigned char *smem, const unsigned sbase, const int lane, const unsigned u,
                                            const double *__restrict__ p1, const double *__restrict__ p2, const long long n_same,
                                            const HbtMixSeg *__restrict__ segs, const int *__restrict__ row_item0, const int n_rows,
                                            const unsigned *__restrict__ units, const HbtGrid &g, const V2Const &c,
                                            const V2Dev *__restrict__ dv, const HbtAccum &acc, const double psi_ref,
                                            const unsigned char *__restrict__ closed, const unsigned *__restrict__ orig,
                                            V2Counters &n, unsigned &cntKT, unsigned &cntRS, unsigned &kept) {
    using L = V3Smem<MIXED, STATS, QINV>;
    constexpr int NC = L::NC, SUB = L::SUB, SUBP = L::SUBP, TJ = L::TJ, IPL = SUB / 32;
    constexpr bool SORTED = L::SORTED;
    static_assert(!(QINV && STATS), "instrumented q_inv runs use the literal kernels");
    double *const si = reinterpret_cast<double *>(smem + L::SI);
    double *const sj = reinterpret_cast<double *>(smem + L::SJ);
    double *const sjt = reinterpret_cast<double *>(smem + L::SJT);
    float *const sjf = reinterpret_cast<float *>(smem + L::SJF);
    unsigned *const si_o = reinterpret_cast<unsigned *>(smem + L::SIO);
    unsigned *const sj_o = reinterpret_cast<unsigned *>(smem + L::SJO);
    const unsigned sjf_addr = sbase + L::SJF;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const double k2lo = c.k2lo, k2hi = c.k2hi, W2 = c.W2;
    V2Queue Q;
    Q.list_addr = sbase + L::LQ + 2u * static_cast<unsigned>(lane);
    Q.cur = Q.list_addr;
    Q.qcount = 0;
    const unsigned lim = opaque_u32(Q.list_addr + 64u * (HBT_V2_LCAP - IPL));

    long long i0, jbase, jcount;  // list-2 particles [jbase, jbase + jcount) belong to this row/segment
    int ni, jt;
    double rc = 1.0, rs = 0.0;
    if (MIXED) {
        const HbtMixSeg sg = segs[find_segment(segs, static_cast<int>(n_same), u)];
        const int local = static_cast<int>(u - sg.block0);
        const int ti = local / sg.tiles_j;
        jt = local - ti * sg.tiles_j;
        i0 = sg.i0 + static_cast<long long>(ti) * SUB;
        ni = min(SUB, sg.ni - ti * SUB);
        jbase = sg.j0;
        jcount = sg.nj;
        rc = sg.c; rs = sg.s;
    } else {
        int a;
        if (SORTED) {
            const unsigned e = units[u];
            a = static_cast<int>(e >> 16);
            jt = static_cast<int>(e & 0xffffu);
        } else {  // every unit of the upper triangle: row a owns units [row_item0[a], row_item0[a+1])
            int lo = 0, hi = n_rows - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (row_item0[mid] <= static_cast<int>(u)) lo = mid; else hi = mid - 1;
            }
            a = lo;
            jt = a * SUB / TJ + (static_cast<int>(u) - row_item0[a]);
        }
        i0 = static_cast<long long>(a) * SUB;
        ni = static_cast<int>(min(static_cast<long long>(SUB), n_same - i0));
        jbase = 0;
        jcount = n_same;
    }
    __syncwarp();  // the previous unit's drain has finished reading the tiles

    // ---- list-1 sub-tile: shared memory (SoA, NaN padding) + this lane's 4 particles -------
    double ax[IPL], ay[IPL], at[IPL];
    float azq[QINV ? IPL : 1], aEq[QINV ? IPL : 1];  // q_inv mode: float p_z, E of this lane's particles
    float Zm1 = 0.f, Em1 = 0.f;                        // q_inv mode: largest |p_z|, |E| of the sub-tile
    unsigned ent[IPL];
    long long ig[IPL];
    double S1 = 0.0, L1 = __longlong_as_double(0x7ff0000000000000ll);  // largest / smallest pT^2 of the sub-tile
#pragma unroll
    for (int s = 0; s < IPL; s++) {
        const int il = s * 32 + lane;
        const int sl = il ^ (s * L::SWZ);  // slot in the swizzled arrays
        ent[s] = static_cast<unsigned>(sl) << 8;
        ig[s] = i0 + il;
        if (il < ni) {
            const double2 *src = reinterpret_cast<const double2 *>(p1 + 8 * (i0 + il));
            const double2 v0 = src[0], v1 = src[1];
            si[sl] = v0.x; si[SUBP + sl] = v0.y; si[2 * SUBP + sl] = v1.x; si[3 * SUBP + sl] = v1.y;
            if (!MIXED) {
                const double2 v2 = src[2], v3 = src[3];
                si[(4 % NC) * SUBP + sl] = v2.x; si[(5 % NC) * SUBP + sl] = v2.y;
                si[(6 % NC) * SUBP + sl] = v3.x; si[(7 % NC) * SUBP + sl] = v3.y;
            }
            if (SORTED) si_o[sl] = orig[i0 + il];
            ax[s] = v0.x; ay[s] = v0.y;
            at[s] = fma(v0.x, v0.x, v0.y * v0.y);
            S1 = fmax(S1, at[s]);
            L1 = fmin(L1, at[s]);
            if (QINV) {
                azq[s] = static_cast<float>(v1.x); aEq[s] = static_cast<float>(v1.y);
                Zm1 = fmaxf(Zm1, fabsf(azq[s])); Em1 = fmaxf(Em1, fabsf(aEq[s]));
            }
        } else {
            if (QINV) { azq[s] = 0.f; aEq[s] = 0.f; }  // (NaN px, py already fail the K_T test)
#pragma unroll
            for (int q = 0; q < NC; q++) si[q * SUBP + sl] = nan;
            if (SORTED) si_o[sl] = 0u;
            ax[s] = nan; ay[s] = nan; at[s] = nan;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S1 = fmax(S1, __shfl_xor_sync(0xffffffffu, S1, o));
    // Mixed-event units, production: |q_out| = |pT_i^2 - pT_j^2| / (2 K_perp) >= |pT_i - pT_j| because
    // K_perp <= (pT_i + pT_j)/2, so a list-2 particle whose pT is farther than W from the whole pT range
    // of this sub-tile cannot be accepted with any of its particles.  The host hands the mixed-event
    // loops a copy in which every event is sorted by pT (rotation invariant), so those particles sit
    // at the two ends of the tile: the pair loop runs over [j_first, j_last] only.  (Positions, not
    // counts: nothing is assumed about the order, an unsorted list just skips less.)
    constexpr bool PTRANGE = MIXED && !STATS && !QINV;
    double pt2_lo = 0.0, pt2_hi = 0.0;
    if (PTRANGE) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) L1 = fmin(L1, __shfl_xor_sync(0xffffffffu, L1, o));
        const double Wd = sqrt(W2) * (1.0 + 1e-9);
        const double a = sqrt(L1) - Wd, b = sqrt(S1) + Wd;
        pt2_lo = a > 0.0 ? a * a * (1.0 - 1e-9) : 0.0;
        pt2_hi = b * b * (1.0 + 1e-9);
        if (!(L1 <= S1)) { pt2_lo = 0.0; pt2_hi = __longlong_as_double(0x7ff0000000000000ll); }  // empty / NaN rows: no restriction
    }
    float2 axf[IPL / 2], naxf[IPL / 2], ayf[IPL / 2], atf[IPL / 2];
    float2 azf[QINV ? IPL / 2 : 1], aEf[QINV ? IPL / 2 : 1];
    if (!STATS) {
#pragma unroll
        for (int h = 0; h < IPL / 2; h++) {
            axf[h] = make_float2(static_cast<float>(ax[2 * h]), static_cast<float>(ax[2 * h + 1]));
            naxf[h] = make_float2(-axf[h].x, -axf[h].y);
            ayf[h] = make_float2(static_cast<float>(ay[2 * h]), static_cast<float>(ay[2 * h + 1]));
            atf[h] = make_float2(static_cast<float>(0.5 * at[2 * h]), static_cast<float>(0.5 * at[2 * h + 1]));
            if (QINV) {
                azf[h] = make_float2(azq[2 * h], azq[2 * h + 1]);
                aEf[h] = make_float2(aEq[2 * h], aEq[2 * h + 1]);
            }
        }
    }
    if (QINV) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Zm1 = fmaxf(Zm1, __shfl_xor_sync(0xffffffffu, Zm1, o));
            Em1 = fmaxf(Em1, __shfl_xor_sync(0xffffffffu, Em1, o));
        }
    }

    {
        const long long jl0 = static_cast<long long>(jt) * TJ;
        const int nj = static_cast<int>(min(static_cast<long long>(TJ), jcount - jl0));
        // ---- stage the list-2 tile (the queue is empty here: entries index this unit) --------
        double S2 = 0.0;
        float Zm2 = 0.f, Em2 = 0.f;
        int j_first = TJ, j_last = -1;  // first position with pT^2 >= pt2_lo, last position with pT^2 <= pt2_hi
        for (int k = lane; k < nj; k += 32) {
            const double2 *src = reinterpret_cast<const double2 *>(p2 + 8 * (jbase + jl0 + k));
            const double2 v0 = src[0], v1 = src[1];
            double x = v0.x, y = v0.y;
            if (MIXED) {  // rotation of the partner event, src/HBT_correlation.cpp:522-523
                x = __dsub_rn(__dmul_rn(v0.x, rc), __dmul_rn(v0.y, rs));
                y = __dadd_rn(__dmul_rn(v0.x, rs), __dmul_rn(v0.y, rc));
            }
            sj[k] = x; sj[TJ + k] = y; sj[2 * TJ + k] = v1.x; sj[3 * TJ + k] = v1.y;
            const double pt2 = fma(x, x, y * y);
            if (STATS) sjt[k] = pt2;  // (no such array in production: the offset is shared with sjf)
            S2 = fmax(S2, pt2);
            if (PTRANGE) {
                if (!(pt2 < pt2_lo)) j_first = min(j_first, k);  // (a NaN stays inside the range)
                if (!(pt2 > pt2_hi)) j_last = k;                  // k increases along the loop
            }
            if (!STATS) { sjf[k] = static_cast<float>(x); sjf[TJ + k] = static_cast<float>(y); sjf[2 * TJ + k] = static_cast<float>(-0.5 * pt2); }
            if (QINV) {
                const float zf = static_cast<float>(v1.x), ef = static_cast<float>(v1.y);
                sjf[3 * TJ + k] = -zf; sjf[4 * TJ + k] = -ef;
                Zm2 = fmaxf(Zm2, fabsf(zf)); Em2 = fmaxf(Em2, fabsf(ef));
            }
            if (SORTED) sj_o[k] = orig[jl0 + k];
            if (!MIXED) {
                const double2 v2 = src[2], v3 = src[3];
                sj[(4 % NC) * TJ + k] = v2.x; sj[(5 % NC) * TJ + k] = v2.y;
                sj[(6 % NC) * TJ + k] = v3.x; sj[(7 % NC) * TJ + k] = v3.y;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) S2 = fmax(S2, __shfl_xor_sync(0xffffffffu, S2, o));
        if (QINV) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                Zm2 = fmaxf(Zm2, __shfl_xor_sync(0xffffffffu, Zm2, o));
                Em2 = fmaxf(Em2, __shfl_xor_sync(0xffffffffu, Em2, o));
            }
        }
        int j_begin = 0, j_end = nj;
        if (PTRANGE) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                j_first = min(j_first, __shfl_xor_sync(0xffffffffu, j_first, o));
                j_last = max(j_last, __shfl_xor_sync(0xffffffffu, j_last, o));
            }
            j_begin = min(j_first, nj);
            j_end = max(j_last + 1, j_begin);
        }
        __syncwarp();
        // prefilter error bound against the smallest K_T (see hbt_kernels_v2.cuh)
        const double S = S1 + S2;
        double k2_floor = (5.6e-17 * S) * S / W2;
        bool use_floor = !(k2_floor <= k2lo);
        // ---- float prefilter margins (u = 2^-24; S bounds every squared momentum component):
        //   |k2_f - k2| <= 64 u S,  |d_f - d| <= 8 u S,  |x_f - x| <= 16 u S  (inputs + every op)
        //   at a window edge (d^2 = W^2 k2): relative error of d2_f / w_f
        //        rho <= 64 u S / (W sqrt(k2)) + 64 u S / k2 + 4u ;   the test allows 4 rho
        float klo_f = 0.f, khi_f = 0.f, kfloor_f = 0.f, Wqf = 0.f;
        if (!STATS) {
            const double u = 5.9604644775390625e-8, Ek = 64.0 * u * S;
            const double Wd = sqrt(W2);
            double k2e = k2lo - Ek > 0.0 ? k2lo - Ek : 0.0;
            double rho = k2e > 0.0 ? 64.0 * u * S / (Wd * sqrt(k2e)) + 64.0 * u * S / k2e + 4.0 * u : 1.0;
            use_floor = !(rho <= 0.03);
            if (use_floor) {  // below this k2 the float window test is not trusted at all
                const double a1 = 64.0 * u * S / (0.015 * Wd), a2 = 64.0 * u * S / 0.015;
                k2_floor = fmax(a1 * a1, a2) + Ek;
                rho = 0.03 + 4.0 * u;
            }
            // the float test compares max(d_f^2, x_f^2) with fl(k2_f * Wqf): Wqf = (W^2/4)(1 + 4 rho) rounded
            // up leaves more than the 2 rho the bound asks for (and the roundings of the three products)
            Wqf = __double2float_ru(0.25 * W2 * (1.0 + 4.0 * rho + 16.0 * u));
            klo_f = __double2float_rd(fmax(k2lo - Ek, 0.0));
            khi_f = __double2float_ru(k2hi + Ek);
            kfloor_f = __double2float_ru(k2_floor);
        }
        // q_inv mode: a pair whose q_inv may lie below q_hi is kept whatever its transverse components.  In floats,
        //   s = q_T^2 + q_z^2 - q_E^2,   q_T^2 = 2 (pT_i^2 + pT_j^2) - k2 = 4 e - k2,   e = pT_i^2/2 + pT_j^2/2,
        // tested as  4 e + q_z^2 <= q_E^2 + k2 + B.  Errors (u = 2^-24; Z, E = largest |p_z|, |E| of tile 1 + tile 2):
        // |k2_f - k2| <= 64 u S, |4 e_f - 4 e| <= 8 u S, |q_z,f^2 - q_z^2| <= 5 u Z^2, same for E, and u (|lhs| + |rhs|)
        // <= u (9 S + Z^2 + E^2 + B) for the two final roundings: B = max(q_hi, 0)^2 + 2 u (81 S + 6 Z^2 + 6 E^2) (x2).
        float Bq = 0.f;
        if (QINV) {
            const double u = 5.9604644775390625e-8;
            const double Zs = static_cast<double>(Zm1) + static_cast<double>(Zm2), Es = static_cast<double>(Em1) + static_cast<double>(Em2);
            Bq = __double2float_ru(static_cast<double>(c.qinv_w2_f) * (1.0 + 4.0 * u) + 2.0 * u * (81.0 * S + 6.0 * Zs * Zs + 6.0 * Es * Es));
        }
        const bool diag = !MIXED && (jl0 < i0 + SUB);  // tile reaches back to the diagonal: j > i only

        // the pair loop, specialised on (unit touches the diagonal, error floor active)
        auto tile_loop = [&](auto diag_c, auto floor_c) {
            constexpr bool DIAG = decltype(diag_c)::value, FLOOR = decltype(floor_c)::value;
            int jthr[IPL];  // DIAG: pair (i, j) is taken when j > i, i.e. local j > jthr
#pragma unroll
            for (int s = 0; s < IPL; s++) jthr[s] = DIAG ? static_cast<int>(ig[s] - jl0) : 0;
            const unsigned lane16 = static_cast<unsigned>(lane) << 8;  // (lane in the slot half of a queue entry)
            int j = j_begin;
            // the float copy of list-2 particle j is loaded one trip ahead (LDS latency off the loop's
            // critical path; the last trip reads the first slot of the next array, see V3Smem)
            float pbx = 0.f, pby = 0.f, pnb = 0.f, pnz = 0.f, pnE = 0.f;
            if (!STATS) {
                const unsigned ja0 = sjf_addr + 4u * static_cast<unsigned>(j_begin);
                pbx = lds_f32(ja0); pby = lds_f32(ja0 + 4 * TJ); pnb = lds_f32(ja0 + 8 * TJ);
                if (QINV) { pnz = lds_f32(ja0 + 12 * TJ); pnE = lds_f32(ja0 + 16 * TJ); }
            }
            for (;;) {
                const bool final = (j >= j_end);  // one extra trip: the per-unit final flush shares the call site
                if (!final) {
                    if (!STATS) {
                        const float bxs = pbx, bys = pby, nbh = pnb, nbz = pnz, nbE = pnE;
                        const unsigned ja = sjf_addr + 4u * static_cast<unsigned>(j + 1);
                        pbx = lds_f32(ja); pby = lds_f32(ja + 4 * TJ); pnb = lds_f32(ja + 8 * TJ);
                        if (QINV) { pnz = lds_f32(ja + 12 * TJ); pnE = lds_f32(ja + 16 * TJ); }
                        const float2 nbz2 = make_float2(nbz, nbz), nbE2 = make_float2(nbE, nbE), pbt2 = make_float2(-nbh, -nbh);
                        const float2 four2 = make_float2(4.f, 4.f), Bq2 = make_float2(Bq, Bq);
                        const float2 bx2 = make_float2(bxs, bxs), by2 = make_float2(bys, bys), nbt2 = make_float2(nbh, nbh);
                        const float2 Wq2 = make_float2(Wqf, Wqf);
                        const unsigned ej = lane16 + static_cast<unsigned>(j);
#pragma unroll
                        for (int h = 0; h < IPL / 2; h++) {
                            const float2 sx = __fadd2_rn(axf[h], bx2), sy = __fadd2_rn(ayf[h], by2);
                            const float2 k2 = __ffma2_rn(sy, sy, __fmul2_rn(sx, sx));
                            const float2 d = __fadd2_rn(atf[h], nbt2);  // (pT_i^2 - pT_j^2) / 2 = K_perp q_out
                            const float2 x = __ffma2_rn(bx2, ayf[h], __fmul2_rn(naxf[h], by2));  // K_perp q_side
                            const float2 d2 = __fmul2_rn(d, d), x2 = __fmul2_rn(x, x), w = __fmul2_rn(k2, Wq2);
                            float2 ql = make_float2(0.f, 0.f), qr = make_float2(0.f, 0.f);
                            if (QINV) {
                                const float2 qz = __fadd2_rn(azf[h], nbz2), qE = __fadd2_rn(aEf[h], nbE2);
                                const float2 e2 = __fadd2_rn(atf[h], pbt2);
                                ql = __ffma2_rn(e2, four2, __fmul2_rn(qz, qz));
                                qr = __ffma2_rn(qE, qE, __fadd2_rn(k2, Bq2));
                            }
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                const int s = 2 * h + e;
                                const float k2e = e ? k2.y : k2.x;
                                // K_T cut with margin (NaN rows fail), then max(d^2, x^2) against
                                // (W^2/4)(1 + margin) k2: dropped only when certainly outside the window
                                bool keep = (k2e >= klo_f) && (k2e <= khi_f);
                                if (DIAG) keep = keep && (j > jthr[s]);
                                const float m = fmaxf(e ? d2.y : d2.x, e ? x2.y : x2.x);
                                bool in = m <= (e ? w.y : w.x);
                                if (FLOOR) in = in || (k2e < kfloor_f);
                                if (QINV) in = in || ((e ? ql.y : ql.x) <= (e ? qr.y : qr.x));
                                // the next slot's address goes to a NEW register: advancing the cursor in place
                                // would wait for the STS to release its address operand (WAR, short scoreboard)
                                const unsigned slot = Q.cur;
                                Q.cur = slot + ((keep && in) ? 64u : 0u);
                                // entry: list-1 slot (lane ^ s SWZ) + 32 s, list-2 position
                                if (keep && in) sts_u16(slot, (ej ^ (static_cast<unsigned>(s * L::SWZ) << 8)) + (static_cast<unsigned>(s) << 13));
                            }
                        }
                    } else {
                        const double bx = sj[j], by = sj[TJ + j], bt = sjt[j];
#pragma unroll
                        for (int s = 0; s < IPL; s++) {
                            const double sx = __dadd_rn(ax[s], bx), sy = __dadd_rn(ay[s], by);
                            const double k2 = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
                            bool kt = (k2 >= k2lo) && (k2 <= k2hi);  // exact K_T cut (NaN padding rows fail)
                            if (DIAG) kt = kt && (j > jthr[s]);
                            const double d = at[s] - bt;
                            const double x = fma(bx, ay[s], -(ax[s] * by));
                            const double d2 = d * d, x2 = x * x, w = W2 * k2;
                            const int hw = __double2hiint(w);
                            const int dd = __double2hiint(d2) - hw;               // d^2   vs W^2 k2
                            const int dx = __double2hiint(x2) + 0x00200000 - hw;  // 4 x^2 vs W^2 k2
                            bool rej_o = dd > 1;                 // q_out certainly outside the window
                            // q_out certainly inside, q_side certainly outside.  Only a window symmetric about zero lets
                            // |q_out| < W stand for "passed the q_out cut" in the stage counters: otherwise the drain decides
                            bool rej_s = (dd < -1) && (dx > 1) && c.symmetric;
                            if (FLOOR) {
                                const bool tiny = k2 < k2_floor;
                                rej_o = rej_o && !tiny;
                                rej_s = rej_s && !tiny;
                            }
                            const bool keep = kt && !(rej_o || rej_s);
                            // exact K_T-pass and q_out-pass populations (instrumented runs only)
                            inc_if(cntKT, kt);
                            inc_if(cntRS, kt && rej_s);
                            if (keep) {
                                sts_u16(Q.cur, ent[s] | static_cast<unsigned>(j));
                                Q.cur += 64u;
                            }
                        }
                    }
                    j++;
                    if (!__any_sync(0xffffffffu, Q.cur > lim)) continue;
                }
                {
                    // compact the per-lane lists into the linear queue and drain it 32 at a time
                    const int cnt = static_cast<int>(Q.cur - Q.list_addr) >> 6;
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    const unsigned dst = sbase + L::WQ + 2u * static_cast<unsigned>(Q.qcount + (incl - cnt));
                    for (int m = 0; m < cnt; m++) sts_u16(dst + 2u * m, lds_u16(Q.list_addr + 64u * m));
                    Q.cur = Q.list_addr;
                    Q.qcount += total;
                    kept += static_cast<unsigned>(total);
                    __syncwarp();
                    // drain from the top of the queue, 32 entries a round; the entry of the NEXT round is
                    // loaded before this round's pair is evaluated (one LDS round trip off the chain)
                    // (every queue read is followed by a __syncwarp before the next flush writes the queue)
                    if (Q.qcount >= 32 || (final && Q.qcount > 0)) {
                        unsigned entry = lds_u16(sbase + L::WQ + 2u * static_cast<unsigned>(max(Q.qcount - 32, 0) + lane));
                        do {
                            const int take = min(32, Q.qcount);
                            const int base = Q.qcount - take;
                            const unsigned next_entry = lds_u16(sbase + L::WQ + 2u * static_cast<unsigned>(max(base - 32, 0) + lane));
                            if (lane < take) v3_drain_pair<MIXED, STATS, QINV>(g, c, acc, closed, dv, sbase, entry, psi_ref, n);
                            entry = next_entry;
                            Q.qcount = base;
                            __syncwarp();
                        } while (Q.qcount >= 32 || (final && Q.qcount > 0));
                    }
                }
                if (final) break;
            }
        };
        if (use_floor) {
            if (diag) tile_loop(std::true_type{}, std::true_type{}); else tile_loop(std::false_type{}, std::true_type{});
        } else {
            if (diag) tile_loop(std::true_type{}, std::false_type{}); else tile_loop(std::false_type{}, std::false_type{});
        }
    }
}

template <bool MIXED, bool STATS, bool QINV = false>
__global__ void __launch_bounds__(32, QINV ? HBT_V3_WARPS_PER_SM_QINV : HBT_V3_WARPS_PER_SM)
hbt_pairs_v3(const double *__restrict__ p1, const double *__restrict__ p2, long long n_same,
             const HbtMixSeg *__restrict__ segs, const int *__restrict__ row_item0, int n_rows,
             const unsigned *__restrict__ units, unsigned *__restrict__ work, unsigned n_units,
             const HbtGrid g, const V2Const c, const V2Dev *__restrict__ dv, const HbtAccum acc,
             const double psi_ref, const unsigned long long total_pairs,
             const unsigned char *__restrict__ closed, const unsigned *__restrict__ orig) {
    using L = V3Smem<MIXED, STATS, QINV>;
    __shared__ __align__(16) unsigned char smem[L::BYTES];
    const int lane = threadIdx.x;
    // kept in a register: no per-use S2UR/ULEA
    const unsigned sbase = opaque_u32(static_cast<unsigned>(__cvta_generic_to_shared(smem)));
    if (blockIdx.x == 0 && lane == 0) atomicAdd(&acc.stage[MIXED ? 6 : 0], total_pairs);
    const unsigned total_units = L::SORTED ? work[1] : n_units;
    V2Counters n = {0, 0, 0, 0};
    unsigned cntKT = 0, cntRS = 0, kept = 0;
    unsigned popped = 0;  // lane 0 pops one unit ahead: the atomic's round trip hides behind the current unit
    if (lane == 0) popped = atomicAdd(&work[0], 1u);
    for (;;) {
        const unsigned u = __shfl_sync(0xffffffffu, popped, 0);
        if (u >= total_units) break;
        if (lane == 0) popped = atomicAdd(&work[0], 1u);
        v3_run_unit<MIXED, STATS, QINV>(smem, sbase, lane, u, p1, p2, n_same, segs, row_item0, n_rows, units, g, c, dv, acc, psi_ref,
                                  closed, orig, n, cntKT, cntRS, kept);
    }

    // ---- counters ------------------------------------------------------------------------
    const unsigned nE = warp_sum(n.nE);
    unsigned long long *stage = acc.stage + (MIXED ? 6 : 0);
    if (STATS) {
        unsigned nB = warp_sum(cntKT + n.nB), nC = warp_sum(cntRS + n.nC);
        const unsigned nD = warp_sum(n.nD);
        nB -= kept;
        if (lane == 0) {
            if (nB) atomicAdd(&stage[1], static_cast<unsigned long long>(nB));
            if (nC) atomicAdd(&stage[2], static_cast<unsigned long long>(nC));
            if (nD) atomicAdd(&stage[3], static_cast<unsigned long long>(nD));
        }
    }
    if (lane == 0 && nE) atomicAdd(&stage[4], static_cast<unsigned long long>(nE));
}

// ---- fused launch: the same-event and the mixed-event units of one batch in one kernel --------
// The same-event kernel alone is bound by the spread-address REDs of its accepted pairs (5 per
// pair: the LSU accepts about one RED lane per cycle per SM; ncu: memory pipes 74 % busy, issue
// slots 50 %), the mixed-event kernel alone by instruction issue in its prefilter (1 RED per
// accepted pair).  Interleaving the two unit lists lets every SM work on both kinds at any time,
// so the RED traffic of the same-event pairs is spread over the whole launch and hides behind
// the mixed-event prefilter.  Unit u of the S + M units is a same-event unit when
// floor((u+1) S / (S+M)) > floor(u S / (S+M)) (S = units kept by hbt_cull_units, read from
// work[1]; M = mixed-event units).  Production mode only (no stage counters).
__global__ void __launch_bounds__(32, HBT_V3_WARPS_PER_SM)
hbt_pairs_v3_fused(const double *__restrict__ ps, const long long n_same, const unsigned *__restrict__ units,
                   const unsigned *__restrict__ orig, const double *__restrict__ p1, const double *__restrict__ p2,
                   const long long n_seg, const HbtMixSeg *__restrict__ segs, const unsigned n_mixed_units,
                   unsigned *__restrict__ work, const HbtGrid g, const V2Const c, const V2Dev *__restrict__ dv, const HbtAccum acc,
                   const double psi_ref, const unsigned long long pairs_same, const unsigned long long pairs_mixed,
                   const unsigned char *__restrict__ closed) {
    using LS = V3Smem<false, false>;
    using LM = V3Smem<true, false>;
    __shared__ __align__(16) unsigned char smem[LS::BYTES > LM::BYTES ? LS::BYTES : LM::BYTES];
    const int lane = threadIdx.x;
    const unsigned sbase = opaque_u32(static_cast<unsigned>(__cvta_generic_to_shared(smem)));
    if (blockIdx.x == 0 && lane == 0) {
        atomicAdd(&acc.stage[0], pairs_same);
        atomicAdd(&acc.stage[6], pairs_mixed);
    }
    const unsigned long long S = work[1], T = S + n_mixed_units;
    V2Counters ns = {0, 0, 0, 0}, nm = {0, 0, 0, 0};
    unsigned unused0 = 0, unused1 = 0, unused2 = 0;
    unsigned popped = 0;  // lane 0 pops one unit ahead
    if (lane == 0) popped = atomicAdd(&work[0], 1u);
    for (;;) {
        const unsigned u = __shfl_sync(0xffffffffu, popped, 0);
        if (u >= T) break;
        if (lane == 0) popped = atomicAdd(&work[0], 1u);
        const unsigned long long s0 = u * S / T, s1 = (u + 1ull) * S / T;
        if (s1 > s0)
            v3_run_unit<false, false>(smem, sbase, lane, static_cast<unsigned>(s0), ps, ps, n_same, nullptr, nullptr, 0, units, g, c,
                                      dv, acc, psi_ref, closed, orig, ns, unused0, unused1, unused2);
        else
            v3_run_unit<true, false>(smem, sbase, lane, static_cast<unsigned>(u - s0), p1, p2, n_seg, segs, nullptr, 0, nullptr, g,
                                     c, dv, acc, psi_ref, closed, nullptr, nm, unused0, unused1, unused2);
    }
    const unsigned eS = warp_sum(ns.nE), eM = warp_sum(nm.nE);
    if (lane == 0) {
        if (eS) atomicAdd(&acc.stage[4], static_cast<unsigned long long>(eS));
        if (eM) atomicAdd(&acc.stage[10], static_cast<unsigned long long>(eM));
    }
}

// ---- host side -----------------------------------------------------------------------------
// instrumented same-event runs visit every unit of the upper triangle: row a (128 particles)
// owns the tiles from the one holding its first particle to the end.  Fills the prefix
// row_item0[0..n_rows] and returns the number of units.
inline long long hbt_v3_same_units(long long n, std::vector<int> &row_item0) {
    const long long n_rows = (n + HBT_V3_SUB_SAME - 1) / HBT_V3_SUB_SAME;
    const long long ntj = (n + HBT_V3_TJ_SAME - 1) / HBT_V3_TJ_SAME;
    row_item0.resize(n_rows + 1);
    long long items = 0;
    for (long long a = 0; a < n_rows; a++) {
        row_item0[a] = static_cast<int>(items);
        items += ntj - a * HBT_V3_SUB_SAME / HBT_V3_TJ_SAME;
    }
    row_item0[n_rows] = static_cast<int>(items);
    return items;
}

#endif  // HBT_KERNELS_V3_CUH_
