// Shared host/device plain-old-data of libhbt_b200 (not part of the public ABI).
#ifndef HBT_COMMON_H_
#define HBT_COMMON_H_

#include <stdint.h>

#include "../../include/hbt_b200.h"

#define HBT_MAX_KT 64  // K_T edges supported by the threshold table

// Grid constants, derived ONCE on the host with the reference's own expressions so that
// host and device compare against bit-identical doubles
// (src/HBT_correlation.cpp:28, :48-49, :255-256, :289-290, :363-364).
struct HbtGrid {
    int32_t nq, nKT, nKphi, az, qinv, boost;
    int32_t nslab;
    int32_t pad0;
    int64_t nbins;
    double KT_min, dKT, KT_min_sq, KT_max_sq;
    double q_base;  // q_min - delta_q/2
    double q_lo;    // q_min - delta_q/2 + 1e-8
    double q_hi;    // q_max + delta_q/2 - 1e-8
    double dq;
    double dKphi, two_pi;
    double hbarc_inv;  // 1/hbarC, src/parameters.h:4
    double rap_lo, rap_hi;  // tanh(HBTrap_min), tanh(HBTrap_max)
    uint64_t needed;        // needed_number_of_pairs as the reference's unsigned long long
    // --- fast-path tables (v2 kernels) ---
    // K_T bin of a pair is the number of k in [1, nKT) with K_perp_sq >= kt_thr_sq[k]:
    // exact thresholds of the monotone step function int((sqrt(x) - KT_min)/dKT), found by
    // bisection over doubles on the host with the reference expression.
    double kt_thr_sq[HBT_MAX_KT];
    double inv_dq;  // 1/delta_q (fast-path index estimate only; never decides an edge)
};

// device-resident accumulators (structure of arrays, see include/hbt_b200.h for layout)
struct HbtAccum {
    unsigned long long *num_count;  // [nbins]
    double *num_cos, *sum_qo, *sum_qs, *sum_ql;  // [nbins]
    unsigned long long *den_count;   // [nbins]
    unsigned long long *npairs_num;  // [nslab]
    unsigned long long *npairs_den;  // [nslab]
    unsigned long long *stage;       // [12]: same {A..E, accepted}, mixed {A..E, accepted}
    // q_inv mode, [nKT*nq] / [nKT]
    unsigned long long *qinv_count, *qinv_den;
    double *qinv_sum, *qinv_cos;
    unsigned long long *npairs_num_qinv, *npairs_den_qinv;
    // pairs deferred to the host's literal evaluation
    struct HbtDeferred *deferred;
    unsigned int *deferred_count;  // [0] = count, [1] = overflow flag
    unsigned int deferred_capacity;
};

struct HbtDeferred {
    double a[8];
    double b[8];  // partner, already rotated for mixed events
    double psi_ref;
    int64_t row, pos;  // position of the pair in the reference's loop order (ordered-cap mode)
    int32_t mixed;
    int32_t pad;
};

// needed_number_of_pairs support (src/HBT_correlation.cpp:402-406, :651-655).  `closed` marks
// slabs whose counter already exceeds the cap (numerator slabs, then denominator slabs).  The
// batch in which a slab crosses the cap is replayed in order: pass 1 counts accepted pairs per
// row (list-1 particle) for the crossing slabs, the host locates the last accepted pair, pass 2
// accumulates only pairs at or before that position.
struct HbtCap {
    // channels: c < nslab = the 3-D histogram of slab c (first needed+1 pairs); nslab + iK = the q_inv
    // histogram of K_T bin iK (first 50*needed pairs, invariant_radius_flag=1)
    const unsigned char *closed;  // [2*nslab + 2*nKT]: slabs (same, mixed), q_inv K_T bins (same, mixed); or null
    const int32_t *xidx;          // pass 1: channel -> index among the crossing channels, or -1
    unsigned int *rowcnt;         // pass 1: [n_crossing][nrows]
    int64_t nrows;
    const int64_t *cut_row;       // pass 2: [n_channels] last accepted row ...
    const int64_t *cut_pos;       // ... and position inside that row
};

// one (event, partner) segment of mixed-event work: rows [i0, i0+ni) of list 1 against rows
// [j0, j0+nj) of list 2 rotated by (c, s).  The segment owns the thread blocks
// [block0, block0 + tiles_i*tiles_j); a block finds its segment by binary search on block0.
struct HbtMixSeg {
    int64_t i0, j0;
    int32_t ni, nj;
    double c, s;
    int64_t block0;
    int64_t pos0;  // position of the segment's first partner particle inside the row (loop order)
    int32_t tiles_j;
    int32_t pad;
};

// host-evaluated stage-counter increments of deferred pairs
struct HbtStageDelta {
    unsigned long long v[12];
};

// a correction produced by the host's literal evaluation of a deferred pair
struct HbtCorrection {
    int64_t bin;
    int32_t slab;
    int32_t mixed;
    double qo, qs, ql, cosv;
};

#ifdef __cplusplus
extern "C" {
#endif
// host side (hbt_host.cpp)
int hbt_host_derive_grid(const hbt_params *p, HbtGrid *g, char *err, int errlen);
// literal evaluation of one pair on the host; returns 1 and fills c when the pair is
// accepted into a 3-D bin, 0 otherwise.  stage[6] is incremented like the device does.
int hbt_host_pair_literal(const HbtGrid *g, const double *a, const double *b, int mixed,
                          double psi_ref, HbtCorrection *c, uint64_t *stage);
int hbt_host_pair_qinv(const HbtGrid *g, const double *a, const double *b, int *iK, int *iq);
void hbt_qinv_thresholds(const HbtGrid *g, double *s_lo, double *s_hi, double *thr /* [nq + 1] */);
#ifdef __cplusplus
}
#endif

#endif  // HBT_COMMON_H_
