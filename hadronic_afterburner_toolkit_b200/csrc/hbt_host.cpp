// Host side of libhbt_b200: the O(N) work the reference does around the pair loops,
// written so that it produces the SAME doubles as the reference (same expressions, glibc
// libm, libstdc++ <random>, compiled without FMA contraction).  No device code here.
//
// Reference lines are cited per function (paths relative to /root/reference).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <random>

#include "hbt_common.h"

namespace {

const double kHbarC = 0.197327053;  // src/parameters.h:4

// K_T bin exactly as the reference computes it from K_perp_sq
// (src/HBT_correlation.cpp:323-324)
inline int kt_index(const HbtGrid &g, double kperp_sq) {
    double kperp = std::sqrt(kperp_sq);
    return static_cast<int>((kperp - g.KT_min) / g.dKT);
}

}  // namespace

extern "C" int hbt_host_derive_grid(const hbt_params *p, HbtGrid *g, char *err, int errlen) {
    std::memset(g, 0, sizeof(*g));
    if (p->qnpts < 2 || p->n_KT < 2 || p->n_KT > HBT_MAX_KT || !(p->q_max > p->q_min)
        || !(p->KT_max > p->KT_min) || p->KT_min < 0.0
        || (p->azimuthal_flag == 1 && p->n_Kphi < 1) || !(p->needed_number_of_pairs >= 0.0)) {
        std::snprintf(err, errlen,
                      "invalid HBT parameters (need qnpts>=2, 2<=n_KT<=%d, q_max>q_min, "
                      "KT_max>KT_min>=0, n_Kphi>=1)", HBT_MAX_KT);
        return HBT_ERR_INVALID;
    }
    g->nq = p->qnpts;
    g->nKT = p->n_KT;
    g->nKphi = p->n_Kphi;
    g->az = p->azimuthal_flag == 1 ? 1 : 0;
    g->qinv = p->invariant_radius_flag == 1 ? 1 : 0;
    g->boost = p->long_comoving_boost == 1 ? 1 : 0;
    g->nslab = g->nKT * (g->az ? g->nKphi : 1);
    g->nbins = static_cast<int64_t>(g->nslab) * g->nq * g->nq * g->nq;
    // src/HBT_correlation.cpp:28, :48-49
    g->dq = (p->q_max - p->q_min) / (p->qnpts - 1);
    g->dKT = (p->KT_max - p->KT_min) / (p->n_KT - 1);
    g->dKphi = 2 * M_PI / p->n_Kphi;
    g->two_pi = 2. * M_PI;
    g->KT_min = p->KT_min;
    // :289-290
    g->KT_min_sq = p->KT_min * p->KT_min;
    g->KT_max_sq = p->KT_max * p->KT_max;
    // :363-364, :368-369
    g->q_lo = p->q_min - g->dq / 2. + 1e-8;
    g->q_hi = p->q_max + g->dq / 2. - 1e-8;
    g->q_base = p->q_min - g->dq / 2.;
    g->inv_dq = 1.0 / g->dq;
    g->hbarc_inv = 1. / kHbarC;  // :253
    // :255-256
    g->rap_hi = std::tanh(p->HBTrap_max);
    g->rap_lo = std::tanh(p->HBTrap_min);
    g->needed = static_cast<unsigned long long>(p->needed_number_of_pairs);  // HBT_correlation.h:40

    // Exact K_T thresholds in K_perp_sq space: kt_index is a non-decreasing step function
    // of its argument (sqrt, subtraction of a constant, division by a positive constant and
    // truncation are all monotone), so bin(x) = #{k >= 1 : x >= thr[k]} with
    // thr[k] = min{x in [KT_min_sq, KT_max_sq] : kt_index(x) >= k}, found by bisection on
    // the ordered bit patterns of non-negative doubles.
    for (int k = 0; k < HBT_MAX_KT; k++) g->kt_thr_sq[k] = INFINITY;
    g->kt_thr_sq[0] = g->KT_min_sq;
    for (int k = 1; k < g->nKT; k++) {
        if (kt_index(*g, g->KT_max_sq) < k) break;  // never reached inside the cut
        uint64_t lo, hi;
        double a = g->KT_min_sq, b = g->KT_max_sq;
        std::memcpy(&lo, &a, 8);
        std::memcpy(&hi, &b, 8);
        if (kt_index(*g, a) >= k) {
            g->kt_thr_sq[k] = a;
            continue;
        }
        while (hi - lo > 1) {  // invariant: index(lo) < k <= index(hi)
            uint64_t mid = lo + (hi - lo) / 2;
            double x;
            std::memcpy(&x, &mid, 8);
            if (kt_index(*g, x) >= k) hi = mid; else lo = mid;
        }
        std::memcpy(&g->kt_thr_sq[k], &hi, 8);
    }
    return HBT_OK;
}

extern "C" int64_t hbt_gather_rapidity(const hbt_params *params, const double *in, int64_t n,
                                       double *out) {
    // src/HBT_correlation.cpp:255-266
    const double cut_hi = std::tanh(params->HBTrap_max);
    const double cut_lo = std::tanh(params->HBTrap_min);
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        const double *q = in + 8 * i;
        const double ratio = q[2] / q[3];
        if (ratio > cut_lo && ratio < cut_hi) {
            std::memcpy(out + 8 * m, q, 64);
            m++;
        }
    }
    return m;
}

extern "C" double hbt_psi_ref(const double *p, int64_t n, int32_t n_order) {
    // src/HBT_correlation.cpp:233-249
    double vn_real = 0.0, vn_imag = 0.0;
    for (int64_t i = 0; i < n; i++) {
        const double phi = std::atan2(p[8 * i + 1], p[8 * i]);
        vn_real += std::cos(n_order * phi);
        vn_imag += std::sin(n_order * phi);
    }
    return std::atan2(vn_imag, vn_real) / n_order;
}

// ---- RNG: the same standard-library objects RandomUtil::Random holds -------------------
struct hbt_rng {
    std::mt19937 gen;
    std::uniform_real_distribution<double> real;  // (0, 1), src/Random.cpp:7-8
    std::uniform_int_distribution<int> integer;   // default range, src/Random.h:19
    explicit hbt_rng(int seed) : gen(seed), real(0.0, 1.0) {}
};

extern "C" int hbt_rng_create(int32_t seed, hbt_rng **out) {
    if (!out) return HBT_ERR_INVALID;
    if (seed == -1) {  // src/Random.cpp:10-12
        std::random_device dev;
        seed = static_cast<int32_t>(dev());
    }
    *out = new (std::nothrow) hbt_rng(seed);
    return *out ? HBT_OK : HBT_ERR_INVALID;
}

extern "C" void hbt_rng_destroy(hbt_rng *rng) { delete rng; }
extern "C" int32_t hbt_rng_int_uniform(hbt_rng *rng) { return rng->integer(rng->gen); }
extern "C" double hbt_rng_uniform(hbt_rng *rng) { return rng->real(rng->gen); }

extern "C" int32_t hbt_rng_mixed_plan(hbt_rng *rng, int32_t nev, int32_t nev_mixed,
                                      int32_t *partner_ids, double *cos_sin, double *angles) {
    if (nev_mixed <= 0) return 0;
    const int nmix = nev_mixed / 2 + 1;  // src/HBT_correlation.cpp:200
    for (int iev = 0; iev < nev; iev++) {
        for (int c = 0; c < nmix; c++) {  // :208-215
            int id = rng->integer(rng->gen) % nev_mixed;
            while (iev == id && nev_mixed != 1) id = rng->integer(rng->gen) % nev_mixed;
            if (partner_ids) partner_ids[static_cast<size_t>(iev) * nmix + c] = id;
        }
        for (int c = 0; c < nmix; c++) {  // :495-497
            const double rot = rng->real(rng->gen) * 2 * M_PI;
            const size_t k = static_cast<size_t>(iev) * nmix + c;
            if (angles) angles[k] = rot;
            if (cos_sin) {
                cos_sin[2 * k] = std::cos(rot);
                cos_sin[2 * k + 1] = std::sin(rot);
            }
        }
    }
    return nmix;
}

// ---- literal evaluation of a single pair (deferred pairs only) -------------------------
// q_inv branch of one pair, literally (src/HBT_correlation.cpp:326-356 / :590-607): 1 when the pair
// enters the q_inv histogram of K_T bin *iK (window and index tests; the 50*needed cap is the
// caller's business), with its bin in *iq.
// q_inv mode on the tuned kernels: the steps of the reference's tests on q_inv = sqrt(s), in s space (s >= 0;
// sqrt, the comparisons, the subtraction of a constant, the division by a positive constant and the truncation are
// all monotone in s).  s_lo / s_hi: q_inv > q_lo <=> s >= s_lo, q_inv < q_hi <=> s < s_hi (src :342-343, :597-598);
// thr[k], k = 0 .. nq: smallest s >= s_lo with int((sqrt(s) - q_base) / delta_q) >= k (:344-345), +inf when no s
// below s_hi reaches bin k; thr[0] = s_lo.  Bisection on the ordered bit patterns of non-negative doubles.
namespace {
template <typename F>
double first_nonneg_true(double hi_start, F pred) {  // smallest s >= 0 with pred(s); pred monotone; +inf if pred(hi_start) is false
    if (pred(0.0)) return 0.0;
    if (!pred(hi_start)) return INFINITY;
    uint64_t lo = 0, hi;
    std::memcpy(&hi, &hi_start, 8);
    while (hi - lo > 1) {  // invariant: !pred(lo), pred(hi)
        const uint64_t mid = lo + (hi - lo) / 2;
        double x;
        std::memcpy(&x, &mid, 8);
        if (pred(x)) hi = mid; else lo = mid;
    }
    double r;
    std::memcpy(&r, &hi, 8);
    return r;
}
}  // namespace

extern "C" void hbt_qinv_thresholds(const HbtGrid *g, double *s_lo, double *s_hi, double *thr) {
    const double q_lo = g->q_lo, q_hi = g->q_hi, q_base = g->q_base, dq = g->dq;
    const double top = 4.0 * (q_hi > 1.0 ? q_hi * q_hi : 1.0) + 4.0;  // sqrt(top) > q_hi
    *s_lo = first_nonneg_true(top, [&](double s) { return std::sqrt(s) > q_lo; });
    *s_hi = first_nonneg_true(top, [&](double s) { return !(std::sqrt(s) < q_hi); });  // +inf never happens: sqrt(top) > q_hi
    for (int k = 0; k <= g->nq; k++) {
        thr[k] = first_nonneg_true(top, [&](double s) { return std::sqrt(s) > q_lo && static_cast<int>((std::sqrt(s) - q_base) / dq) >= k; });
        if (!(thr[k] < *s_hi)) thr[k] = INFINITY;
    }
    thr[0] = *s_lo;
}

extern "C" int hbt_host_pair_qinv(const HbtGrid *g, const double *a, const double *b, int *iK, int *iq) {
    const double Kx = 0.5 * (a[0] + b[0]);
    const double Ky = 0.5 * (a[1] + b[1]);
    const double K2 = Kx * Kx + Ky * Ky;
    if (!(K2 >= g->KT_min_sq && K2 <= g->KT_max_sq)) return 0;
    *iK = static_cast<int>((std::sqrt(K2) - g->KT_min) / g->dKT);
    const double qx = a[0] - b[0], qy = a[1] - b[1], qz = a[2] - b[2], qE = a[3] - b[3];
    const double qinv = std::sqrt(-(qE * qE - qx * qx - qy * qy - qz * qz));
    if (!(qinv > g->q_lo && qinv < g->q_hi)) return 0;
    *iq = static_cast<int>((qinv - g->q_base) / g->dq);
    return *iq < g->nq ? 1 : 0;
}

// The device hands a pair back when its K_phi bin decision sits within 1e-9 of an edge:
// there glibc's atan2 (which the reference uses) and CUDA's atan2 may disagree.  The chain
// below is src/HBT_correlation.cpp:311-458 (same event) / :574-687 (mixed event).
extern "C" int hbt_host_pair_literal(const HbtGrid *g, const double *a, const double *b, int mixed,
                                     double psi_ref, HbtCorrection *c, uint64_t *stage) {
    const double Kz = 0.5 * (a[2] + b[2]);
    const double KE = 0.5 * (a[3] + b[3]);
    const double Kx = 0.5 * (a[0] + b[0]);
    const double Ky = 0.5 * (a[1] + b[1]);
    const double K2 = Kx * Kx + Ky * Ky;
    if (!(K2 >= g->KT_min_sq && K2 <= g->KT_max_sq)) return 0;
    stage[1]++;
    const double Kp = std::sqrt(K2);
    const int iK = static_cast<int>((Kp - g->KT_min) / g->dKT);
    const double qx = a[0] - b[0], qy = a[1] - b[1], qz = a[2] - b[2], qE = a[3] - b[3];
    const double cphi = Kx / Kp, sphi = Ky / Kp;
    const double qo = qx * cphi + qy * sphi;
    if (!(qo >= g->q_lo) || (mixed ? !(qo < g->q_hi) : !(qo <= g->q_hi))) return 0;
    const int io = static_cast<int>((qo - g->q_base) / g->dq);
    if (io >= g->nq) return 0;
    stage[2]++;
    const double qs = qy * cphi - qx * sphi;
    if (!(qs >= g->q_lo) || (mixed ? !(qs < g->q_hi) : !(qs <= g->q_hi))) return 0;
    const int is = static_cast<int>((qs - g->q_base) / g->dq);
    if (is >= g->nq) return 0;
    stage[3]++;
    double ql = qz;
    if (g->boost) {
        const double Mt = std::sqrt(KE * KE - Kz * Kz);
        const double gamma = KE / Mt;
        const double beta = Kz / KE;
        ql = gamma * (qz - beta * qE);
    }
    if (!(ql >= g->q_lo) || (mixed ? !(ql < g->q_hi) : !(ql <= g->q_hi))) return 0;
    const int il = static_cast<int>((ql - g->q_base) / g->dq);
    if (il >= g->nq) return 0;
    stage[4]++;
    int slab = iK;
    if (g->az) {
        double dphi = std::atan2(Ky, Kx) - psi_ref;
        while (dphi < 0.) dphi += 2. * M_PI;
        while (dphi > 2. * M_PI) dphi -= 2. * M_PI;
        const int iphi = static_cast<int>(dphi / g->dKphi);
        if (iphi < 0 || iphi >= g->nKphi) return 0;
        slab = iK * g->nKphi + iphi;
    }
    stage[5]++;
    c->slab = slab;
    c->mixed = mixed;
    c->bin = ((static_cast<int64_t>(slab) * g->nq + io) * g->nq + is) * g->nq + il;
    c->qo = qo;
    c->qs = qs;
    c->ql = ql;
    c->cosv = 0.0;
    if (!mixed) {
        const double td = a[7] - b[7], xd = a[4] - b[4], yd = a[5] - b[5], zd = a[6] - b[6];
        c->cosv = std::cos(g->hbarc_inv * (qE * td - qx * xd - qy * yd - qz * zd));
    }
    return 1;
}
