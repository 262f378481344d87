// Device-side pair arithmetic of libhbt_b200 (sm_100a).
//
// LITERAL chain: every operation that decides a cut or a bin edge is one IEEE-754 double
// operation in the reference's order — explicit _rn intrinsics, which nvcc never contracts
// into FMAs — because the reference binary is built without FMA (baseline x86-64, SSE2
// divsd/sqrtsd; /root/reference/CMakeLists.txt:24-28).  Lines cited are
// src/HBT_correlation.cpp.
#ifndef HBT_PAIR_CUH_
#define HBT_PAIR_CUH_

#include "hbt_common.h"

struct PairBin {
    int slab, io, is, il;
    double qo, qs, ql;
};

// status of a pair after the chain
enum : int {
    PAIR_REJ_KT = 0,    // failed the K_T cut              (:319-321 / :581-583)
    PAIR_REJ_QO = 1,    // failed the q_out window          (:363-370 / :612-618)
    PAIR_REJ_QS = 2,    // failed the q_side window         (:373-380 / :621-628)
    PAIR_REJ_QL = 3,    // failed the q_long window         (:392-398 / :640-647)
    PAIR_REJ_PHI = 4,   // K_phi index out of range          (:417-423 / :666-672)
    PAIR_ACCEPT = 5,
    PAIR_DEFER = 6      // K_phi edge too close to call on the device: host decides
};

__device__ __forceinline__ bool in_window(double q, double lo, double hi, bool mixed) {
    // same-event: reject if q < lo || q > hi ; mixed: reject if q < lo || q >= hi.
    // NaN is rejected (the reference would index out of bounds there).
    return mixed ? (q >= lo && q < hi) : (q >= lo && q <= hi);
}

// momentum part of the literal chain.  p = (px,py,pz,E).
__device__ __forceinline__ int pair_literal(const HbtGrid &g, double px1, double py1, double pz1,
                                            double E1, double px2, double py2, double pz2,
                                            double E2, bool mixed, double psi_ref, PairBin &o) {
    const double Kx = __dmul_rn(0.5, __dadd_rn(px1, px2));  // :316-318
    const double Ky = __dmul_rn(0.5, __dadd_rn(py1, py2));
    const double K2 = __dadd_rn(__dmul_rn(Kx, Kx), __dmul_rn(Ky, Ky));
    if (!(K2 >= g.KT_min_sq && K2 <= g.KT_max_sq)) return PAIR_REJ_KT;
    const double Kp = __dsqrt_rn(K2);  // :323-324
    const int iK = __double2int_rz(__ddiv_rn(__dsub_rn(Kp, g.KT_min), g.dKT));

    const double qx = __dsub_rn(px1, px2);  // :327-330
    const double qy = __dsub_rn(py1, py2);
    const double cphi = __ddiv_rn(Kx, Kp);  // :359-360
    const double sphi = __ddiv_rn(Ky, Kp);

    const double qo = __dadd_rn(__dmul_rn(qx, cphi), __dmul_rn(qy, sphi));  // :362
    if (!in_window(qo, g.q_lo, g.q_hi, mixed)) return PAIR_REJ_QO;
    const int io = __double2int_rz(__ddiv_rn(__dsub_rn(qo, g.q_base), g.dq));  // :368-370
    if (io >= g.nq) return PAIR_REJ_QO;

    const double qs = __dsub_rn(__dmul_rn(qy, cphi), __dmul_rn(qx, sphi));  // :372
    if (!in_window(qs, g.q_lo, g.q_hi, mixed)) return PAIR_REJ_QS;
    const int is = __double2int_rz(__ddiv_rn(__dsub_rn(qs, g.q_base), g.dq));
    if (is >= g.nq) return PAIR_REJ_QS;

    const double qz = __dsub_rn(pz1, pz2);
    double ql = qz;
    if (g.boost) {  // :383-390
        const double Kz = __dmul_rn(0.5, __dadd_rn(pz1, pz2));  // :311-313
        const double KE = __dmul_rn(0.5, __dadd_rn(E1, E2));
        const double beta = __ddiv_rn(Kz, KE);
        const double qE = __dsub_rn(E1, E2);
        const double Mt = __dsqrt_rn(__dsub_rn(__dmul_rn(KE, KE), __dmul_rn(Kz, Kz)));
        const double gamma = __ddiv_rn(KE, Mt);
        ql = __dmul_rn(gamma, __dsub_rn(qz, __dmul_rn(beta, qE)));
    }
    if (!in_window(ql, g.q_lo, g.q_hi, mixed)) return PAIR_REJ_QL;
    const int il = __double2int_rz(__ddiv_rn(__dsub_rn(ql, g.q_base), g.dq));
    if (il >= g.nq) return PAIR_REJ_QL;

    int slab = iK;
    if (g.az) {  // :408-423
        // CUDA's atan2 is within 2 ulp of glibc's; the bin is trusted only when the scaled
        // angle is farther than 1e-9 from every integer, otherwise the host decides.
        double dphi = __dsub_rn(atan2(Ky, Kx), psi_ref);
        while (dphi < 0.) dphi = __dadd_rn(dphi, g.two_pi);
        while (dphi > g.two_pi) dphi = __dsub_rn(dphi, g.two_pi);
        const double u = __ddiv_rn(dphi, g.dKphi);
        if (!(u == u)) return PAIR_REJ_PHI;  // NaN angle (a NaN momentum or psi_ref): the reference's int cast gives INT_MIN, :417-423
        if (fabs(u - rint(u)) < 1e-9) return PAIR_DEFER;
        const int iphi = __double2int_rz(u);
        if (iphi < 0 || iphi >= g.nKphi) return PAIR_REJ_PHI;
        slab = iK * g.nKphi + iphi;
    }
    o.slab = slab;
    o.io = io;
    o.is = is;
    o.il = il;
    o.qo = qo;
    o.qs = qs;
    o.ql = ql;
    return PAIR_ACCEPT;
}

// cos(q.dx / hbarC), :431-433.  Feeds a sum with a 1e-10 tolerance, so FMA contraction and
// CUDA's cos (<= 2 ulp) are fine here.
__device__ __forceinline__ double pair_cos(const HbtGrid &g, double qx, double qy, double qz,
                                           double qE, double xd, double yd, double zd, double td) {
    return cos(g.hbarc_inv * (qE * td - qx * xd - qy * yd - qz * zd));
}

__device__ __forceinline__ long long bin_index(const HbtGrid &g, const PairBin &b) {
    return ((static_cast<long long>(b.slab) * g.nq + b.io) * g.nq + b.is) * g.nq + b.il;
}

__device__ __forceinline__ void defer_pair(const HbtAccum &acc, const double *a, const double *b,
                                           double psi_ref, int mixed, long long row = -1, long long pos = -1) {
    const unsigned slot = atomicAdd(&acc.deferred_count[0], 1u);
    if (slot >= acc.deferred_capacity) {
        acc.deferred_count[1] = 1u;  // overflow: reported as HBT_ERR_OVERFLOW at sync
        return;
    }
    HbtDeferred &d = acc.deferred[slot];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        d.a[k] = a[k];
        d.b[k] = b[k];
    }
    d.psi_ref = psi_ref;
    d.row = row;
    d.pos = pos;
    d.mixed = mixed;
    d.pad = 0;
}

#endif  // HBT_PAIR_CUH_
