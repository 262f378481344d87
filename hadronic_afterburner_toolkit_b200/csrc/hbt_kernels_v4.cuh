// v4: the production mixed-event kernel (sm_100a).  Same pipeline as the mixed-event units of
// hbt_kernels_v3.cuh — persistent single-warp CTAs popping 128 x 128 units, packed-FP32 prefilter,
// per-lane survivor lists, warp queue, FP32 decision of the survivors with proven bands
// (v3_mixed_f32_core) — as a kernel of its own, built around what a mixed-event pair needs
// (src/HBT_correlation.cpp:563-689: one count into one bin) and nothing else:
//
//  * the tiles are kept in binary32 only: list 1 as float4 {px, py, pz, E}, list 2 (rotated in
//    binary64 as the reference does, :522-523, then rounded once) as the prefilter's float [3][TJ]
//    and a float4 copy for the drain.  The drain reads two LDS.128 per survivor instead of eight
//    LDS.64 + eight F2F.F32.F64; 8.6 KB of shared memory per warp instead of 11.6;
//  * no binary64 chain in the kernel body.  The ~1e-3 of the survivors the FP32 bands leave open
//    are parked — (list-1 index, list-2 index, segment) in a small per-warp list — and evaluated
//    32 at a time, all lanes busy, by the literal chain out of line (v2_slow_pair: the reference's
//    own operations on the binary64 particles, re-read from global memory and rotated again), so
//    every bin index still equals the reference's.  With the chain out of the way the kernel needs
//    80 registers and 24 warps are resident per SM (v3: 96 registers, 18 warps; the mixed-event
//    units are latency bound and want warps: profiles/r02_controls.txt);
//  * 16-bit queue entries (list-1 slot << 8 | list-2 position), per-lane lists of 16 entries.
//
// Accumulation is a commutative integer count, so parking changes nothing in the result.  Used far
// from the needed-pairs cap only (like every tuned kernel); q_inv mode, instrumented runs and
// HBT_B200_F32MIX=0 stay on v3.  A whole batch launches it NEXT TO the same-event kernel
// (hbt_b200.cu: launch_split_pair): both carry cudaFuncAttributePreferredSharedMemoryCarveout =
// max, without which the two would not be resident on the same SM.  DESIGN.md 5, "Two kernels
// next to each other"; numbers in profiles/r03_*.
#ifndef HBT_KERNELS_V4_CUH_
#define HBT_KERNELS_V4_CUH_

#include "hbt_kernels_v3.cuh"

#ifndef HBT_V4_WARPS_PER_SM
#define HBT_V4_WARPS_PER_SM 24
#endif
#ifndef HBT_V4_LCAP
#define HBT_V4_LCAP 16  // per-lane survivor list capacity: the lists are compacted and drained when one of them holds more than LCAP - 4
#endif
#define HBT_V4_QCAP (32 + 32 * HBT_V4_LCAP)
#ifndef HBT_V4_TI
#define HBT_V4_TI 128  // list-1 particles per unit (4 per lane)
#endif
#define HBT_V4_PARK 64  // parked (undecided) pairs per warp: a drain round adds at most 32, 32 are evaluated as soon as they are there

struct V4Smem {
    static constexpr int TI = HBT_V4_TI, TJ = 128;
    static constexpr int ESH = 7;  // queue entry = list-1 slot << 7 | list-2 position (slots below 512)
    // list-1 sub-tile {px, py, pz, E}, two slots of padding after every 32 records: lane l keeps particles l, l + 32,
    // l + 64, l + 96 (s = 0 .. 3), whose 16-byte records would share their four banks — and a quarter-warp of a drain
    // round reads the records of two or three neighbouring lanes (4-way bank conflicts on every gather).  Particle
    // l + 32 s is kept in slot l + 34 s; queue entries carry the slot.  (Padding, not an XOR swizzle: one base register
    // and immediates in the staging stores.)
    static constexpr int TIP = TI + 2 * (TI / 32 - 1);
    static constexpr int SI = 0;                              // float4 [TIP]
    // float [3][TJ] px, py, -pT^2/2 of the rotated list-2 tile (prefilter; the slot past the last array is read one
    // trip ahead: it is SJ4[0], which nobody writes during the pair loop)
    static constexpr int SJF = SI + 16 * TIP;
    static constexpr int SJ4 = SJF + 12 * TJ;                 // float4 [TJ]    the same particles {px, py, pz, E} (drain)
    static constexpr int LQ = SJ4 + 16 * TJ;                  // u16 [LCAP][32] per-lane survivor lists (slot << 8 | position)
    static constexpr int WQ = LQ + 2 * HBT_V4_LCAP * 32;      // u16 [QCAP]     linear warp queue
    static constexpr int PK = (WQ + 2 * HBT_V4_QCAP + 3) & ~3;  // u32 [3][PARK]  parked pairs: list-1 index, list-2 index, segment
    static constexpr int BYTES = PK + 12 * HBT_V4_PARK;
};

__device__ __forceinline__ float4 lds_f32x4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// Packed binary32 pairs as 64-bit values (Blackwell FADD2 / FMUL2 / FFMA2 take .b64 operands).  The loop-invariant
// list-1 operands are packed ONCE per unit by a volatile mov.b64, which the compiler may not re-create inside the pair
// loop: with float2 operands ptxas sometimes keeps the halves in unrelated registers and re-packs them in every trip.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t p2_pack_once(float lo, float hi) {
    f32x2_t r;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2_t p2_pack(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 p2_unpack(f32x2_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2_t p2_add(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t p2_mul(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t p2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// One parked pair through the literal chain: the two particles in binary64 from global memory, list 2 rotated as the
// staging does (src :522-523).  Returns 1 when the pair was accepted (counted through q_long), else 0.
__device__ __noinline__ unsigned v4_parked_pair(const V2Dev *__restrict__ dv, const double *__restrict__ p1,
                                                const double *__restrict__ p2, const HbtMixSeg *__restrict__ segs, unsigned gi,
                                                unsigned gj, unsigned seg, double psi_ref) {
    const double rc = segs[seg].c, rs = segs[seg].s;
    const double *pa = p1 + 8ull * gi, *pb = p2 + 8ull * gj;
    double a8[8], b8[8];
#pragma unroll
    for (int k = 0; k < 4; k++) { a8[k] = pa[k]; a8[4 + k] = 0.0; b8[4 + k] = 0.0; }
    const double bx = pb[0], by = pb[1];
    b8[0] = __dsub_rn(__dmul_rn(bx, rc), __dmul_rn(by, rs));
    b8[1] = __dadd_rn(__dmul_rn(bx, rs), __dmul_rn(by, rc));
    b8[2] = pb[2];
    b8[3] = pb[3];
    V2Counters tmp = {0, 0, 0, 0};
    v2_slow_pair<true>(dv, a8, b8, psi_ref, tmp);
    return tmp.nE;
}

// evaluates `count` parked pairs (count <= 32) from position `first` of the warp's list
__device__ __forceinline__ unsigned v4_run_parked(const unsigned sbase, const int lane, const int first, const int count,
                                                  const V2Dev *__restrict__ dv, const double *__restrict__ p1,
                                                  const double *__restrict__ p2, const HbtMixSeg *__restrict__ segs,
                                                  const double psi_ref) {
    unsigned acc = 0;
    if (lane < count) {
        const unsigned at = sbase + V4Smem::PK + 4u * static_cast<unsigned>(first + lane);
        acc = v4_parked_pair(dv, p1, p2, segs, lds_u32(at), lds_u32(at + 4 * HBT_V4_PARK), lds_u32(at + 8 * HBT_V4_PARK), psi_ref);
    }
    __syncwarp();
    return acc;
}

__global__ void __launch_bounds__(32, HBT_V4_WARPS_PER_SM)
hbt_pairs_v4_mixed(const double *__restrict__ p1, const double *__restrict__ p2, const int n_seg,
                   const HbtMixSeg *__restrict__ segs, unsigned *__restrict__ work, const unsigned n_units, const HbtGrid g,
                   const V2Const c, const V2Dev *__restrict__ dv, const HbtAccum acc, const double psi_ref,
                   const unsigned long long total_pairs, const unsigned char *__restrict__ closed) {
    using L = V4Smem;
    constexpr int TI = L::TI, TJ = L::TJ, IPL = TI / 32;
    __shared__ __align__(16) unsigned char smem[L::BYTES];
    // (through an identity shuffle: a value ptxas keeps in a register instead of re-reading SR_TID in every trip)
    const int lane = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x), static_cast<int>(threadIdx.x));
    const unsigned sbase = opaque_u32(static_cast<unsigned>(__cvta_generic_to_shared(smem)));  // kept in a register: no per-use S2UR/ULEA
    if (blockIdx.x == 0 && lane == 0) atomicAdd(&acc.stage[6], total_pairs);
    float *const sjf = reinterpret_cast<float *>(smem + L::SJF);
    float4 *const sj4 = reinterpret_cast<float4 *>(smem + L::SJ4);
    const unsigned sjf_addr = sbase + L::SJF;
    const double k2lo = c.k2lo, k2hi = c.k2hi, W2 = c.W2;
    const unsigned list_addr = sbase + L::LQ + 2u * static_cast<unsigned>(lane);
    static_assert(HBT_V4_LCAP > IPL && L::TIP <= 512 && TJ == 128, "queue entries: 9 bits of slot, 7 bits of position");
    const unsigned lim = opaque_u32(list_addr + 64u * (HBT_V4_LCAP - IPL));
    unsigned nE = 0;
    int parked = 0;  // pairs in the warp's parked list (warp-uniform)
    int seg_hint = 0;
    unsigned popped = 0;  // lane 0 pops one unit ahead: the atomic's round trip hides behind the current unit
    if (lane == 0) popped = atomicAdd(work, 1u);
    for (;;) {
        const unsigned u = __shfl_sync(0xffffffffu, popped, 0);
        if (u >= n_units) break;
        if (lane == 0) popped = atomicAdd(work, 1u);
        // ---- the unit's segment: units are popped in increasing order, so the search starts where the last one ended
        int si;
        {
            int lo = seg_hint, step = 32;
            while (lo + step < n_seg && segs[lo + step].block0 <= static_cast<long long>(u)) { lo += step; step <<= 1; }
            int hi = min(lo + step, n_seg) - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (segs[mid].block0 <= static_cast<long long>(u)) lo = mid; else hi = mid - 1;
            }
            si = lo;
            seg_hint = lo;
        }
        const HbtMixSeg sg = segs[si];
        const int local = static_cast<int>(u - sg.block0);
        const int ti = local / sg.tiles_j;
        const int jt = local - ti * sg.tiles_j;
        const long long i0 = sg.i0 + static_cast<long long>(ti) * TI;
        const int ni = min(TI, sg.ni - ti * TI);
        const long long j0 = sg.j0 + static_cast<long long>(jt) * TJ;
        const int nj = min(TJ, sg.nj - jt * TJ);
        const double rc = sg.c, rs = sg.s;
        __syncwarp();  // the previous unit's drain has finished reading the tiles

        // ---- list-1 sub-tile: float4 records (NaN padding) + this lane's 4 particles as packed floats
        f32x2_t axf[IPL / 2], ayf[IPL / 2], atf[IPL / 2];
        double S1 = 0.0, L1 = __longlong_as_double(0x7ff0000000000000ll);  // largest / smallest pT^2 of the sub-tile
        {
            float fx[IPL], fy[IPL], ft[IPL];
#pragma unroll
            for (int s = 0; s < IPL; s++) {
                const int il = s * 32 + lane;
                const unsigned ra = sbase + L::SI + 16u * static_cast<unsigned>(il + 2 * s);  // slot
                if (il < ni) {
                    const double2 *src = reinterpret_cast<const double2 *>(p1 + 8 * (i0 + il));
                    const double2 v0 = src[0], v1 = src[1];
                    const double t = fma(v0.x, v0.x, v0.y * v0.y);
                    S1 = fmax(S1, t);
                    L1 = fmin(L1, t);
                    fx[s] = static_cast<float>(v0.x); fy[s] = static_cast<float>(v0.y); ft[s] = static_cast<float>(0.5 * t);
                    // (scalar stores: one STS.128 would pin the four values to a register quad next to the packed operands)
                    sts_f32(ra, fx[s]); sts_f32(ra + 4, fy[s]);
                    sts_f32(ra + 8, static_cast<float>(v1.x)); sts_f32(ra + 12, static_cast<float>(v1.y));
                } else {
                    const float nf = __int_as_float(0x7fc00000);  // NaN rows fail the K_T test
                    fx[s] = nf; fy[s] = nf; ft[s] = nf;
                    sts_f32(ra, nf); sts_f32(ra + 4, nf); sts_f32(ra + 8, nf); sts_f32(ra + 12, nf);
                }
            }
#pragma unroll
            for (int h = 0; h < IPL / 2; h++) {
                // (through an identity shuffle: values ptxas cannot re-create from the registers the staging stores pinned
                // — it would re-pack the pairs from there in every trip of the pair loop, 20 moves per trip)
                axf[h] = p2_pack_once(__shfl_sync(0xffffffffu, fx[2 * h], lane), __shfl_sync(0xffffffffu, fx[2 * h + 1], lane));
                ayf[h] = p2_pack_once(__shfl_sync(0xffffffffu, fy[2 * h], lane), __shfl_sync(0xffffffffu, fy[2 * h + 1], lane));
                atf[h] = p2_pack_once(__shfl_sync(0xffffffffu, ft[2 * h], lane), __shfl_sync(0xffffffffu, ft[2 * h + 1], lane));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            S1 = fmax(S1, __shfl_xor_sync(0xffffffffu, S1, o));
            L1 = fmin(L1, __shfl_xor_sync(0xffffffffu, L1, o));
        }
        // |q_out| = |pT_i^2 - pT_j^2| / (2 K_perp) >= |pT_i - pT_j| because K_perp <= (pT_i + pT_j)/2: a list-2 particle
        // whose pT is farther than W from the whole pT range of this sub-tile cannot be accepted with any of its
        // particles.  The host hands over a copy in which every event is sorted by pT (rotation invariant), so those
        // particles sit at the two ends of the tile: the pair loop runs over [j_first, j_last] only.  (Positions, not
        // counts: nothing is assumed about the order, an unsorted list just skips less.)
        double pt2_lo, pt2_hi;
        {
            const double Wd = sqrt(W2) * (1.0 + 1e-9);
            const double a = sqrt(L1) - Wd, b = sqrt(S1) + Wd;
            pt2_lo = a > 0.0 ? a * a * (1.0 - 1e-9) : 0.0;
            pt2_hi = b * b * (1.0 + 1e-9);
            if (!(L1 <= S1)) { pt2_lo = 0.0; pt2_hi = __longlong_as_double(0x7ff0000000000000ll); }  // empty / NaN rows: no restriction
        }

        // ---- stage the list-2 tile (the queue is empty here: entries index this unit)
        double S2 = 0.0;
        int j_first = TJ, j_last = -1;  // first position with pT^2 >= pt2_lo, last position with pT^2 <= pt2_hi
        for (int k = lane; k < nj; k += 32) {
            const double2 *src = reinterpret_cast<const double2 *>(p2 + 8 * (j0 + k));
            const double2 v0 = src[0], v1 = src[1];
            // rotation of the partner event, src/HBT_correlation.cpp:522-523
            const double x = __dsub_rn(__dmul_rn(v0.x, rc), __dmul_rn(v0.y, rs));
            const double y = __dadd_rn(__dmul_rn(v0.x, rs), __dmul_rn(v0.y, rc));
            const double pt2 = fma(x, x, y * y);
            S2 = fmax(S2, pt2);
            if (!(pt2 < pt2_lo)) j_first = min(j_first, k);  // (a NaN stays inside the range)
            if (!(pt2 > pt2_hi)) j_last = k;                  // k increases along the loop
            const float xf = static_cast<float>(x), yf = static_cast<float>(y);
            sjf[k] = xf; sjf[TJ + k] = yf; sjf[2 * TJ + k] = static_cast<float>(-0.5 * pt2);
            sj4[k] = make_float4(xf, yf, static_cast<float>(v1.x), static_cast<float>(v1.y));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            S2 = fmax(S2, __shfl_xor_sync(0xffffffffu, S2, o));
            j_first = min(j_first, __shfl_xor_sync(0xffffffffu, j_first, o));
            j_last = max(j_last, __shfl_xor_sync(0xffffffffu, j_last, o));
        }
        const int j_begin = min(j_first, nj);
        const int j_end = max(j_last + 1, j_begin);
        __syncwarp();

        // ---- float prefilter margins (hbt_kernels_v3.cuh; u = 2^-24, S bounds every squared momentum component):
        //   |k2_f - k2| <= 64 u S,  |d_f - d| <= 8 u S,  |x_f - x| <= 16 u S  (inputs + every op)
        //   at a window edge (d^2 = W^2 k2): relative error of d2_f / w_f
        //        rho <= 64 u S / (W sqrt(k2)) + 64 u S / k2 + 4u ;   the test allows 4 rho
        float klo_f, khi_f, kfloor_f, Wqf;
        bool use_floor;
        {
            const double S = S1 + S2;
            const double uu = 5.9604644775390625e-8, Ek = 64.0 * uu * S;
            const double Wd = sqrt(W2);
            const double k2e = k2lo - Ek > 0.0 ? k2lo - Ek : 0.0;
            double rho = k2e > 0.0 ? 64.0 * uu * S / (Wd * sqrt(k2e)) + 64.0 * uu * S / k2e + 4.0 * uu : 1.0;
            double k2_floor = 0.0;
            use_floor = !(rho <= 0.03);
            if (use_floor) {  // below this k2 the float window test is not trusted at all
                const double a1 = 64.0 * uu * S / (0.015 * Wd), a2 = 64.0 * uu * S / 0.015;
                k2_floor = fmax(a1 * a1, a2) + Ek;
                rho = 0.03 + 4.0 * uu;
            }
            // the float test compares max(d_f^2, x_f^2) with fl(k2_f * Wqf): Wqf = (W^2/4)(1 + 4 rho) rounded up leaves
            // more than the 2 rho the bound asks for (and the roundings of the three products)
            Wqf = __double2float_ru(0.25 * W2 * (1.0 + 4.0 * rho + 16.0 * uu));
            klo_f = __double2float_rd(fmax(k2lo - Ek, 0.0));
            khi_f = __double2float_ru(k2hi + Ek);
            kfloor_f = __double2float_ru(k2_floor);
        }

        // ---- the pair loop, specialised on (error floor active)
        auto tile_loop = [&](auto floor_c) {
            constexpr bool FLOOR = decltype(floor_c)::value;
            const unsigned lane16 = opaque_u32(static_cast<unsigned>(lane) << L::ESH);  // (lane in the slot part of a queue entry)
            unsigned cur = list_addr;
            int qcount = 0;
            int j = j_begin;
            // the float copy of list-2 particle j is loaded one trip ahead
            const unsigned ja0 = sjf_addr + 4u * static_cast<unsigned>(j_begin);
            float pbx = lds_f32(ja0), pby = lds_f32(ja0 + 4 * TJ), pnb = lds_f32(ja0 + 8 * TJ);
            for (;;) {
                const bool final = (j >= j_end);  // one extra trip: the per-unit final flush shares the call site
                if (!final) {
                    const float bxs = pbx, bys = pby, nbh = pnb;
                    const unsigned ja = sjf_addr + 4u * static_cast<unsigned>(j + 1);
                    pbx = lds_f32(ja); pby = lds_f32(ja + 4 * TJ); pnb = lds_f32(ja + 8 * TJ);
                    const f32x2_t bx2 = p2_pack(bxs, bxs), by2 = p2_pack(bys, bys), nby2 = p2_pack(-bys, -bys);
                    const f32x2_t nbt2 = p2_pack(nbh, nbh), Wq2 = p2_pack(Wqf, Wqf);
                    // (one add per trip, not one multiply-add under every survivor's predicate: volatile keeps it here)
                    unsigned ej;
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(ej) : "r"(lane16), "r"(static_cast<unsigned>(j)));
#pragma unroll
                    for (int h = 0; h < IPL / 2; h++) {
                        const f32x2_t sx = p2_add(axf[h], bx2), sy = p2_add(ayf[h], by2);
                        const f32x2_t k2p = p2_fma(sy, sy, p2_mul(sx, sx));
                        const f32x2_t d = p2_add(atf[h], nbt2);  // (pT_i^2 - pT_j^2) / 2 = K_perp q_out
                        const f32x2_t x = p2_fma(bx2, ayf[h], p2_mul(axf[h], nby2));  // K_perp q_side
                        const float2 k2 = p2_unpack(k2p), d2 = p2_unpack(p2_mul(d, d)), x2 = p2_unpack(p2_mul(x, x)), w = p2_unpack(p2_mul(k2p, Wq2));
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int s = 2 * h + e;
                            const float k2e = e ? k2.y : k2.x;
                            // K_T cut with margin (NaN rows fail), then max(d^2, x^2) against (W^2/4)(1 + margin) k2:
                            // dropped only when certainly outside the window
                            const bool keep = (k2e >= klo_f) && (k2e <= khi_f);
                            const float m = fmaxf(e ? d2.y : d2.x, e ? x2.y : x2.x);
                            bool in = m <= (e ? w.y : w.x);
                            if (FLOOR) in = in || (k2e < kfloor_f);
                            // the next slot's address goes to a NEW register: advancing the cursor in place would wait
                            // for the STS to release its address operand (WAR, short scoreboard)
                            const unsigned slot = cur;
                            cur = slot + ((keep && in) ? 64u : 0u);
                            if (keep && in) sts_u16(slot, ej + static_cast<unsigned>(s) * (34u << L::ESH));  // list-1 slot lane + 34 s, list-2 position
                        }
                    }
                    j++;
                    if (!__any_sync(0xffffffffu, cur > lim)) continue;
                }
                {
                    // compact the per-lane lists into the linear queue and drain it 32 at a time
                    const int cnt = static_cast<int>(cur - list_addr) >> 6;
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    const unsigned dst = sbase + L::WQ + 2u * static_cast<unsigned>(qcount + (incl - cnt));
                    for (int m = 0; m < cnt; m++) sts_u16(dst + 2u * m, lds_u16(list_addr + 64u * m));
                    cur = list_addr;
                    qcount += total;
                    __syncwarp();
                    // drain from the top of the queue, 32 entries a round; the entry of the NEXT round is loaded before
                    // this round's pair is evaluated (every queue read is followed by a __syncwarp before the next
                    // flush writes the queue)
                    if (qcount >= 32 || (final && qcount > 0)) {
                        unsigned entry = lds_u16(sbase + L::WQ + 2u * static_cast<unsigned>(max(qcount - 32, 0) + lane));
                        do {
                            const int take = min(32, qcount);
                            const int base = qcount - take;
                            const unsigned next_entry = lds_u16(sbase + L::WQ + 2u * static_cast<unsigned>(max(base - 32, 0) + lane));
                            // every lane evaluates (idle lanes on a stale entry of this unit: the tiles are there)
                            // il: slot (< 512 even in a stale entry of an idle lane: inside this warp's shared memory)
                            const unsigned il = entry >> L::ESH, jl = entry & static_cast<unsigned>(TJ - 1);
                            static_assert(16 * 512 <= L::BYTES, "a stale slot must stay inside the warp's shared memory");
                            const float4 a = lds_f32x4(sbase + L::SI + 16u * il), b = lds_f32x4(sbase + L::SJ4 + 16u * jl);
                            int slab;
                            unsigned bin;
                            int fs = v3_mixed_f32_core(g, c, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, psi_ref, slab, bin);
                            if (lane >= take) fs = 0;
                            if (fs > 0) {
                                nE++;
                                if (!(closed && closed[slab + g.nslab])) red_inc_u64(&acc.den_count[bin]);  // (else: needed_number_of_pairs reached earlier)
                            }
                            // undecided: parked, evaluated by the literal chain 32 at a time
                            const unsigned um = __ballot_sync(0xffffffffu, fs < 0);
                            if (um) {
                                if (fs < 0) {
                                    const unsigned at = sbase + L::PK + 4u * static_cast<unsigned>(parked + __popc(um & ((1u << lane) - 1u)));
                                    sts_u32(at, static_cast<unsigned>(i0) + il - 2u * (il / 34u));  // slot -> particle
                                    sts_u32(at + 4 * HBT_V4_PARK, static_cast<unsigned>(j0) + jl);
                                    sts_u32(at + 8 * HBT_V4_PARK, static_cast<unsigned>(si));
                                }
                                parked += __popc(um);
                                __syncwarp();
                                if (parked >= 32) {
                                    parked -= 32;
                                    nE += v4_run_parked(sbase, lane, parked, 32, dv, p1, p2, segs, psi_ref);
                                }
                            }
                            entry = next_entry;
                            qcount = base;
                            __syncwarp();
                        } while (qcount >= 32 || (final && qcount > 0));
                    }
                }
                if (final) break;
            }
        };
        if (use_floor) tile_loop(std::true_type{}); else tile_loop(std::false_type{});
    }
    if (parked > 0) nE += v4_run_parked(sbase, lane, 0, parked, dv, p1, p2, segs, psi_ref);
    nE = warp_sum(nE);
    if (lane == 0 && nE) atomicAdd(&acc.stage[10], static_cast<unsigned long long>(nE));
}

#endif  // HBT_KERNELS_V4_CUH_
