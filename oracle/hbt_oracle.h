/* TEST INFRASTRUCTURE — CPU restatement of the reference's HBT pair-correlation path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  The product
 * (hadronic_afterburner_toolkit_b200/, include/hbt_b200.h) never links or calls it.
 *
 * Parity status: PINNED — tests/test_oracle_vs_reference.py runs this restatement
 * against the unmodified reference compiled into oracle/_ref (bit-exact on every
 * accumulator, counts and sums alike), and tests/golden/ holds reference dumps for
 * the unit-test fixtures so the pin also holds where /root/reference is absent.
 *
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference).
 */
#ifndef HBT_ORACLE_H_
#define HBT_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the parameters.dat switches read at src/HBT_correlation.cpp:22-46 */
typedef struct {
    int32_t qnpts;
    int32_t n_KT;
    int32_t n_Kphi;
    int32_t azimuthal_flag;
    int32_t invariant_radius_flag;
    int32_t long_comoving_boost;
    double q_min, q_max;
    double KT_min, KT_max;
    double HBTrap_min, HBTrap_max;
    double needed_number_of_pairs;
} oracle_params;

typedef struct oracle_state oracle_state;

oracle_state *oracle_create(const oracle_params *p, int32_t seed);
void oracle_destroy(oracle_state *s);

/* RandomUtil::Random (src/Random.h:21-22, src/Random.cpp:7-14) on libstdc++ 13 */
int32_t oracle_rand_int_uniform(oracle_state *s);
double oracle_rand_uniform(oracle_state *s);

/* HBT_correlation::calculate_flow_event_plane_angle (src/HBT_correlation.cpp:233-249);
 * p = particles of all events of the batch, 8 doubles each (px,py,pz,E,x,y,z,t) */
double oracle_psi_ref(const double *p, int64_t n, int32_t n_order);

/* HBT_correlation::calculate_HBT_correlation_function (src/HBT_correlation.cpp:177-218).
 * Events are given as one flat particle array plus nev+1 offsets.  mixed == NULL means
 * the mixed-event lists alias the same-event lists (src/particleSamples.cpp:528-530).
 * do_mixed = 0 stops after the same-event loop (no RNG is consumed). */
void oracle_process_batch(oracle_state *s, const double *same, const int64_t *same_off,
                          int32_t nev, const double *mixed, const int64_t *mixed_off,
                          int32_t nev_mixed, int32_t do_mixed);

/* consume the draws of a batch another rank owns (same rejection loop, no pair work) */
void oracle_skip_batch(oracle_state *s, int32_t nev, int32_t nev_mixed);

/* raw accumulators; 3-D histograms are flat [(K*(az?n_Kphi:1)+phi)][o][s][l] */
int64_t oracle_nbins(const oracle_state *s);
const double *oracle_num_count(const oracle_state *s);
const double *oracle_num_cos(const oracle_state *s);
const double *oracle_sum_qo(const oracle_state *s);
const double *oracle_sum_qs(const oracle_state *s);
const double *oracle_sum_ql(const oracle_state *s);
const double *oracle_den_count(const oracle_state *s);
/* per-K (az=0: n_KT) or per-(K,phi) (az=1: n_KT*n_Kphi) accepted-pair counters */
const uint64_t *oracle_npairs_num(const oracle_state *s);
const uint64_t *oracle_npairs_den(const oracle_state *s);
/* q_inv mode: [n_KT][qnpts] count, sum q_inv, sum cos, denominator; [n_KT] counters */
const double *oracle_qinv_count(const oracle_state *s);
const double *oracle_qinv_mean(const oracle_state *s);
const double *oracle_qinv_num(const oracle_state *s);
const double *oracle_qinv_den(const oracle_state *s);
const uint64_t *oracle_npairs_num_qinv(const oracle_state *s);
const uint64_t *oracle_npairs_den_qinv(const oracle_state *s);
/* stage populations {all pairs, passed K_T cut, passed q_out, passed q_side,
 * passed q_long, accepted after the cap / K_phi checks}; [0..5] same, [6..11] mixed */
const uint64_t *oracle_stage_counters(const oracle_state *s);
double oracle_last_psi_ref(const oracle_state *s);
/* mixed-event plan of the LAST processed batch: nev*nmix partner ids and angles */
int32_t oracle_last_nmix(const oracle_state *s);
const int32_t *oracle_last_partner_ids(const oracle_state *s);
const double *oracle_last_angles(const oracle_state *s);
/* seconds spent inside the same-event / mixed-event pair loops so far */
double oracle_time_same(const oracle_state *s);
double oracle_time_mixed(const oracle_state *s);

#ifdef __cplusplus
}
#endif
#endif /* HBT_ORACLE_H_ */
