/* TEST INFRASTRUCTURE — see hbt_oracle.h.  Plain-C restatement of
 * /root/reference/src/HBT_correlation.cpp (pair loops) and src/Random.{h,cpp}.
 * Compiled with -ffp-contract=off on baseline x86-64 so that every +,-,*,/,sqrt is one
 * IEEE-754 double operation in the reference's association order; libm calls
 * (cos, sin, atan2, tanh) are glibc's, as in the reference.
 */
#define _GNU_SOURCE
#include "hbt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define HBARC 0.197327053 /* src/parameters.h:4 */

/* ---------------------------------------------------------------------------------- */
/* std::mt19937 + the libstdc++ 13 distribution mappings used by src/Random.h:15-22.   */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt19937_t;

static void mt_seed(mt19937_t *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static uint32_t mt_next(mt19937_t *g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            g->mt[i] = v;
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

struct oracle_state {
    oracle_params p;
    mt19937_t rng;
    /* derived exactly as in the constructor, src/HBT_correlation.cpp:28,48-49 */
    double delta_q, dKT, dKphi;
    unsigned long long needed;
    int64_t nslab, nbins;
    double *num_count, *num_cos, *sum_qo, *sum_qs, *sum_ql, *den_count;
    uint64_t *npairs_num, *npairs_den;
    double *qinv_count, *qinv_mean, *qinv_num, *qinv_den;
    uint64_t *npairs_num_qinv, *npairs_den_qinv;
    uint64_t stage[12];
    double psi_ref;
    int32_t last_nmix, last_nev;
    int32_t *last_ids;
    double *last_angles;
    double t_same, t_mixed;
};

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* uniform_int_distribution<int>() over [0, INT_MAX] with a 32-bit engine: libstdc++'s
 * multiply-shift downscaling reduces to one draw shifted right by one. */
int32_t oracle_rand_int_uniform(oracle_state *s) { return (int32_t)(mt_next(&s->rng) >> 1); }

/* uniform_real_distribution<double>(0,1) = generate_canonical<double,53>: two draws,
 * (lo + hi*2^32) / 2^64, clamped below 1. */
double oracle_rand_uniform(oracle_state *s) {
    double lo = (double)mt_next(&s->rng);
    double hi = (double)mt_next(&s->rng);
    double sum = lo + hi * 4294967296.0;
    double r = sum / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

oracle_state *oracle_create(const oracle_params *p, int32_t seed) {
    oracle_state *s = (oracle_state *)calloc(1, sizeof(*s));
    s->p = *p;
    mt_seed(&s->rng, (uint32_t)seed);
    s->delta_q = (p->q_max - p->q_min) / (p->qnpts - 1);
    s->dKT = (p->KT_max - p->KT_min) / (p->n_KT - 1);
    s->dKphi = 2 * M_PI / p->n_Kphi;
    s->needed = (unsigned long long)p->needed_number_of_pairs;
    s->nslab = (int64_t)p->n_KT * (p->azimuthal_flag == 1 ? p->n_Kphi : 1);
    s->nbins = s->nslab * p->qnpts * p->qnpts * p->qnpts;
    s->num_count = (double *)calloc(s->nbins, sizeof(double));
    s->num_cos = (double *)calloc(s->nbins, sizeof(double));
    s->sum_qo = (double *)calloc(s->nbins, sizeof(double));
    s->sum_qs = (double *)calloc(s->nbins, sizeof(double));
    s->sum_ql = (double *)calloc(s->nbins, sizeof(double));
    s->den_count = (double *)calloc(s->nbins, sizeof(double));
    s->npairs_num = (uint64_t *)calloc(s->nslab, sizeof(uint64_t));
    s->npairs_den = (uint64_t *)calloc(s->nslab, sizeof(uint64_t));
    int64_t n1 = (int64_t)p->n_KT * p->qnpts;
    s->qinv_count = (double *)calloc(n1, sizeof(double));
    s->qinv_mean = (double *)calloc(n1, sizeof(double));
    s->qinv_num = (double *)calloc(n1, sizeof(double));
    s->qinv_den = (double *)calloc(n1, sizeof(double));
    s->npairs_num_qinv = (uint64_t *)calloc(p->n_KT, sizeof(uint64_t));
    s->npairs_den_qinv = (uint64_t *)calloc(p->n_KT, sizeof(uint64_t));
    return s;
}

void oracle_destroy(oracle_state *s) {
    if (!s) return;
    free(s->num_count); free(s->num_cos); free(s->sum_qo); free(s->sum_qs);
    free(s->sum_ql); free(s->den_count); free(s->npairs_num); free(s->npairs_den);
    free(s->qinv_count); free(s->qinv_mean); free(s->qinv_num); free(s->qinv_den);
    free(s->npairs_num_qinv); free(s->npairs_den_qinv);
    free(s->last_ids); free(s->last_angles);
    free(s);
}

/* src/HBT_correlation.cpp:233-249 */
double oracle_psi_ref(const double *p, int64_t n, int32_t n_order) {
    double re = 0.0, im = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double phi = atan2(p[8 * i + 1], p[8 * i + 0]);
        re += cos(n_order * phi);
        im += sin(n_order * phi);
    }
    return atan2(im, re) / n_order;
}

/* single-particle rapidity cut of the gathers, src/HBT_correlation.cpp:255-266,468-476:
 * copies the accepted particles of events [off[ev], off[ev+1]) to out, optionally rotated
 * about z as at :522-533.  Returns the number kept. */
static int64_t gather(const oracle_state *s, const double *p, int64_t b, int64_t e, int rotate,
                      double c, double sn, double *out) {
    const double cut_hi = tanh(s->p.HBTrap_max);
    const double cut_lo = tanh(s->p.HBTrap_min);
    int64_t n = 0;
    for (int64_t i = b; i < e; i++) {
        const double *q = p + 8 * i;
        double ratio = q[2] / q[3];
        if (ratio > cut_lo && ratio < cut_hi) {
            double *o = out + 8 * n++;
            if (rotate) {
                o[0] = q[0] * c - q[1] * sn;
                o[1] = q[0] * sn + q[1] * c;
                o[4] = q[4] * c - q[5] * sn;
                o[5] = q[4] * sn + q[5] * c;
            } else {
                o[0] = q[0]; o[1] = q[1]; o[4] = q[4]; o[5] = q[5];
            }
            o[2] = q[2]; o[3] = q[3]; o[6] = q[6]; o[7] = q[7];
        }
    }
    return n;
}

/* One pair through the chain K_T cut -> K_T bin -> [q_inv] -> q_out -> q_side -> q_long
 * -> [K_phi] -> cap -> accumulate.  mixed=0: src/HBT_correlation.cpp:311-458 (upper
 * edges tested with '>');  mixed=1: :574-687 (upper edges tested with '>='). */
static void pair(oracle_state *s, const double *a, const double *b, int mixed) {
    const oracle_params *P = &s->p;
    uint64_t *st = s->stage + (mixed ? 6 : 0);
    const int nq = P->qnpts;
    const double dq = s->delta_q;
    const double lo = P->q_min - dq / 2. + 1e-8;
    const double hi = P->q_max + dq / 2. - 1e-8;
    const double base = P->q_min - dq / 2.;
    const double KT_min_sq = P->KT_min * P->KT_min;
    const double KT_max_sq = P->KT_max * P->KT_max;

    st[0]++;
    double K_z = 0.5 * (a[2] + b[2]);
    double K_E = 0.5 * (a[3] + b[3]);
    double beta = K_z / K_E;
    double K_x = 0.5 * (a[0] + b[0]);
    double K_y = 0.5 * (a[1] + b[1]);
    double K_perp_sq = K_x * K_x + K_y * K_y;
    if (K_perp_sq < KT_min_sq || K_perp_sq > KT_max_sq) return;
    st[1]++;
    double K_perp = sqrt(K_perp_sq);
    int iK = (int)((K_perp - P->KT_min) / s->dKT);

    double q_x = a[0] - b[0], q_y = a[1] - b[1], q_z = a[2] - b[2], q_E = a[3] - b[3];
    double t_d = a[7] - b[7], x_d = a[4] - b[4], y_d = a[5] - b[5], z_d = a[6] - b[6];

    if (P->invariant_radius_flag == 1) {
        double q_inv = sqrt(-(q_E * q_E - q_x * q_x - q_y * q_y - q_z * q_z));
        if (!mixed) { /* :339-356 */
            if (s->npairs_num_qinv[iK] < 50 * s->needed) {
                if (q_inv > lo && q_inv < hi) {
                    int iq = (int)((q_inv - base) / dq);
                    s->npairs_num_qinv[iK]++;
                    double c = cos((1. / HBARC) * (q_E * t_d - q_x * x_d - q_y * y_d - q_z * z_d));
                    s->qinv_count[iK * nq + iq]++;
                    s->qinv_mean[iK * nq + iq] += q_inv;
                    s->qinv_num[iK * nq + iq] += c;
                }
            }
        } else { /* :595-607 */
            if (q_inv > lo && q_inv < hi) {
                int iq = (int)((q_inv - base) / dq);
                if (iq < nq && s->npairs_den_qinv[iK] < 50 * s->needed) {
                    s->npairs_den_qinv[iK]++;
                    s->qinv_den[iK * nq + iq] += 1.0;
                }
            }
        }
    }

    double cphi = K_x / K_perp;
    double sphi = K_y / K_perp;

    double q_out = q_x * cphi + q_y * sphi;
    if (q_out < lo || (mixed ? q_out >= hi : q_out > hi)) return;
    int io = (int)((q_out - base) / dq);
    if (io >= nq) return;
    st[2]++;

    double q_side = q_y * cphi - q_x * sphi;
    if (q_side < lo || (mixed ? q_side >= hi : q_side > hi)) return;
    int is = (int)((q_side - base) / dq);
    if (is >= nq) return;
    st[3]++;

    double q_long = q_z;
    if (P->long_comoving_boost == 1) { /* :383-390 */
        double Mt = sqrt(K_E * K_E - K_z * K_z);
        double gamma = K_E / Mt;
        q_long = gamma * (q_z - beta * q_E);
    }
    if (q_long < lo || (mixed ? q_long >= hi : q_long > hi)) return;
    int il = (int)((q_long - base) / dq);
    if (il >= nq) return;
    st[4]++;

    int64_t slab;
    uint64_t *cap = mixed ? s->npairs_den : s->npairs_num;
    if (P->azimuthal_flag == 0) { /* :401-406, :650-655 */
        slab = iK;
    } else { /* :408-428, :657-677 */
        double dphi = atan2(K_y, K_x) - s->psi_ref;
        while (dphi < 0.) dphi += 2. * M_PI;
        while (dphi > 2. * M_PI) dphi -= 2. * M_PI;
        int iphi = (int)(dphi / s->dKphi);
        if (iphi < 0 || iphi >= P->n_Kphi) return;
        slab = (int64_t)iK * P->n_Kphi + iphi;
    }
    if (cap[slab] > s->needed) return;
    cap[slab]++;
    st[5]++;

    int64_t bin = ((slab * nq + io) * nq + is) * nq + il;
    if (mixed) {
        s->den_count[bin] += 1.0;
    } else {
        double c = cos((1. / HBARC) * (q_E * t_d - q_x * x_d - q_y * y_d - q_z * z_d));
        s->num_count[bin]++;
        s->sum_qo[bin] += q_out;
        s->sum_qs[bin] += q_side;
        s->sum_ql[bin] += q_long;
        s->num_cos[bin] += c;
    }
}

void oracle_process_batch(oracle_state *s, const double *same, const int64_t *same_off,
                          int32_t nev, const double *mixed, const int64_t *mixed_off,
                          int32_t nev_mixed, int32_t do_mixed) {
    if (!mixed) { /* aliasing of src/particleSamples.cpp:528-530 */
        mixed = same;
        mixed_off = same_off;
        nev_mixed = nev;
    }
    int64_t ntot = nev > 0 ? same_off[nev] : 0;
    if (s->p.azimuthal_flag == 1) /* :181-183; all filtered particles, no rapidity cut */
        s->psi_ref = oracle_psi_ref(same, ntot, 2);

    /* same-event: all events merged into one list (:257-281), i<j loop (:291-301) */
    double *l1 = (double *)malloc((size_t)(ntot > 0 ? ntot : 1) * 64);
    int64_t n = gather(s, same, 0, ntot, 0, 0., 0., l1);
    double t0 = now_s();
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = i + 1; j < n; j++) pair(s, l1 + 8 * i, l1 + 8 * j, 0);
    s->t_same += now_s() - t0;
    if (!do_mixed) {
        free(l1);
        return;
    }

    /* mixed events, :197-217 */
    int nmix = nev_mixed / 2 + 1;
    s->last_nmix = nmix;
    s->last_nev = nev;
    free(s->last_ids);
    free(s->last_angles);
    s->last_ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nev > 0 ? nev : 1) * nmix);
    s->last_angles = (double *)malloc(sizeof(double) * (size_t)(nev > 0 ? nev : 1) * nmix);
    int64_t cap2 = 0;
    for (int e = 0; e < nev_mixed; e++) {
        int64_t m = mixed_off[e + 1] - mixed_off[e];
        if (m > cap2) cap2 = m;
    }
    double *l2 = (double *)malloc((size_t)(cap2 * nmix > 0 ? cap2 * nmix : 1) * 64);
    for (int iev = 0; iev < nev; iev++) {
        int32_t *ids = s->last_ids + (size_t)iev * nmix;
        for (int c = 0; c < nmix; c++) { /* :208-215 */
            int id = oracle_rand_int_uniform(s) % nev_mixed;
            while (iev == id && nev_mixed != 1) id = oracle_rand_int_uniform(s) % nev_mixed;
            ids[c] = id;
        }
        int64_t n1 = gather(s, same, same_off[iev], same_off[iev + 1], 0, 0., 0., l1);
        int64_t n2 = 0;
        for (int c = 0; c < nmix; c++) { /* :493-552 */
            double ang = oracle_rand_uniform(s) * 2 * M_PI;
            s->last_angles[(size_t)iev * nmix + c] = ang;
            double cs = cos(ang), sn = sin(ang);
            n2 += gather(s, mixed, mixed_off[ids[c]], mixed_off[ids[c] + 1], 1, cs, sn, l2 + 8 * n2);
        }
        t0 = now_s();
        for (int64_t i = 0; i < n1; i++)
            for (int64_t j = 0; j < n2; j++) pair(s, l1 + 8 * i, l2 + 8 * j, 1);
        s->t_mixed += now_s() - t0;
    }
    free(l1);
    free(l2);
}

void oracle_skip_batch(oracle_state *s, int32_t nev, int32_t nev_mixed) {
    if (nev_mixed <= 0) return;
    int nmix = nev_mixed / 2 + 1;
    for (int iev = 0; iev < nev; iev++) {
        for (int c = 0; c < nmix; c++) { /* src/HBT_correlation.cpp:208-215 */
            int id = oracle_rand_int_uniform(s) % nev_mixed;
            while (iev == id && nev_mixed != 1) id = oracle_rand_int_uniform(s) % nev_mixed;
        }
        for (int c = 0; c < nmix; c++) (void)oracle_rand_uniform(s); /* :495 */
    }
}

int64_t oracle_nbins(const oracle_state *s) { return s->nbins; }
const double *oracle_num_count(const oracle_state *s) { return s->num_count; }
const double *oracle_num_cos(const oracle_state *s) { return s->num_cos; }
const double *oracle_sum_qo(const oracle_state *s) { return s->sum_qo; }
const double *oracle_sum_qs(const oracle_state *s) { return s->sum_qs; }
const double *oracle_sum_ql(const oracle_state *s) { return s->sum_ql; }
const double *oracle_den_count(const oracle_state *s) { return s->den_count; }
const uint64_t *oracle_npairs_num(const oracle_state *s) { return s->npairs_num; }
const uint64_t *oracle_npairs_den(const oracle_state *s) { return s->npairs_den; }
const double *oracle_qinv_count(const oracle_state *s) { return s->qinv_count; }
const double *oracle_qinv_mean(const oracle_state *s) { return s->qinv_mean; }
const double *oracle_qinv_num(const oracle_state *s) { return s->qinv_num; }
const double *oracle_qinv_den(const oracle_state *s) { return s->qinv_den; }
const uint64_t *oracle_npairs_num_qinv(const oracle_state *s) { return s->npairs_num_qinv; }
const uint64_t *oracle_npairs_den_qinv(const oracle_state *s) { return s->npairs_den_qinv; }
const uint64_t *oracle_stage_counters(const oracle_state *s) { return s->stage; }
double oracle_last_psi_ref(const oracle_state *s) { return s->psi_ref; }
int32_t oracle_last_nmix(const oracle_state *s) { return s->last_nmix; }
const int32_t *oracle_last_partner_ids(const oracle_state *s) { return s->last_ids; }
const double *oracle_last_angles(const oracle_state *s) { return s->last_angles; }
double oracle_time_same(const oracle_state *s) { return s->t_same; }
double oracle_time_mixed(const oracle_state *s) { return s->t_mixed; }
