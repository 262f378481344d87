/* hbt_b200 — C ABI of the B200-native HBT pair-correlation engine (libhbt_b200.so).
 *
 * This is the drop-in boundary for ONE path of chunshen1987/hadronic_afterburner_toolkit:
 * the same-event and mixed-event pair loops of class HBT_correlation.  Everything the
 * reference does around those loops (parameters.dat, the particleSamples reader, the RNG
 * stream, the output files) stays on the host side of this boundary.  Citations are
 * relative to the reference tree.
 *
 *   reference                                                 replaced by
 *   ---------------------------------------------------------  ------------------------------
 *   HBT_correlation ctor, histogram allocation                 hbt_create
 *     (src/HBT_correlation.cpp:16-143)
 *   ~HBT_correlation (:145-175)                                hbt_destroy
 *   gather + rapidity cut (:255-281, :468-489, :499-511)       hbt_gather_rapidity (host helper)
 *   calculate_flow_event_plane_angle (:233-249)                hbt_psi_ref        (host helper)
 *   partner draw + rotation angles (:206-215, :495-497)        hbt_rng_mixed_plan (host helper;
 *                                                              the C++ class uses the reference's
 *                                                              own RandomUtil::Random instead)
 *   combine_and_bin_particle_pairs pair loop (:291-460)        hbt_accumulate_same[_dev]
 *   combine_and_bin_particle_pairs_mixed_events loop (:563-689) hbt_accumulate_mixed[_dev]
 *   one whole batch of calculate_HBT_correlation_function      hbt_accumulate_batch
 *     (:177-218)
 *   reading the accumulators for output_* (:694-855)           hbt_read, hbt_read_qinv
 *
 * Conventions: plain C types only; every function returns 0 (HBT_OK) or a negative
 * error code, the message is available from hbt_last_error; no C++ exceptions cross the
 * ABI.  Calls on one context must come from one host thread at a time, in batch order.
 * Work is asynchronous on the context's CUDA stream; hbt_synchronize / hbt_read wait.
 * There is NO CPU fallback: without a usable CUDA device hbt_create fails.
 *
 * Particle layout everywhere: 8 doubles per particle, px,py,pz,E,x,y,z,t — the 64 hot
 * bytes of particle_info (src/particle_info.h:5-12) in the order the loops read them.
 * Histogram layout: flat [slab][q_out][q_side][q_long], slab = K_T index
 * (azimuthal_flag = 0, n_KT slabs) or K_T index * n_Kphi + K_phi index (azimuthal_flag = 1).
 * As in the reference, n_KT counts K_T EDGES and the slab n_KT-1 exists (it receives pairs
 * with K_T == KT_max exactly) but is never written to a file.
 */
#ifndef HBT_B200_H_
#define HBT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HBT_OK 0
#define HBT_ERR_INVALID (-1)     /* bad argument / parameter */
#define HBT_ERR_CUDA (-2)        /* CUDA runtime or kernel failure */
#define HBT_ERR_NO_DEVICE (-3)   /* no usable CUDA device: there is no CPU path */
#define HBT_ERR_CAP (-4)         /* needed_number_of_pairs was reached in a mode that
                                    does not implement the ordered cap */
#define HBT_ERR_OVERFLOW (-5)    /* a device-side side list overflowed */
#define HBT_ERR_NCCL (-6)
#define HBT_ERR_STATE (-7)

/* The parameters.dat switches of the path (src/HBT_correlation.cpp:22-46), verbatim.
 * All grid constants (delta_q, window edges, dKT, dKphi, KT^2 cuts, tanh rapidity cuts)
 * are derived inside hbt_create with the reference's own expressions (:28, :48-49,
 * :255-256, :289-290, :363-364). */
typedef struct hbt_params {
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
} hbt_params;

typedef struct hbt_ctx hbt_ctx;
typedef struct hbt_rng hbt_rng;

/* ---- lifetime -------------------------------------------------------------------- */
/* replaces HBT_correlation::HBT_correlation (src/HBT_correlation.cpp:16-143):
 * allocates and zeroes every accumulator in the HBM of CUDA device `device`. */
int hbt_create(const hbt_params *params, int32_t device, hbt_ctx **out);
/* replaces HBT_correlation::~HBT_correlation (:145-175) */
void hbt_destroy(hbt_ctx *ctx);
/* message of the last failure on ctx (ctx == NULL: of the last failed hbt_create) */
const char *hbt_last_error(const hbt_ctx *ctx);
int64_t hbt_num_bins(const hbt_ctx *ctx);
int32_t hbt_num_slabs(const hbt_ctx *ctx);
int hbt_reset(hbt_ctx *ctx); /* zero all accumulators (a new analysis, same grid) */

/* ---- host helpers (reference-identical host arithmetic, glibc libm) --------------- */
/* single-particle rapidity cut tanh(HBTrap_min) < pz/E < tanh(HBTrap_max) of the gathers
 * (src/HBT_correlation.cpp:255-266, :468-476, :509-511).  Copies the accepted particles of
 * `in` (n x 8) to `out` in order and returns how many were kept. */
int64_t hbt_gather_rapidity(const hbt_params *params, const double *in, int64_t n, double *out);
/* HBT_correlation::calculate_flow_event_plane_angle (:233-249) over n particles */
double hbt_psi_ref(const double *p, int64_t n, int32_t n_order);
/* The RNG stream of the mixed-event step.  hbt_rng is std::mt19937 with the same
 * distribution objects as RandomUtil::Random (src/Random.h:15-22, src/Random.cpp:7-14). */
int hbt_rng_create(int32_t seed, hbt_rng **out);
void hbt_rng_destroy(hbt_rng *rng);
int32_t hbt_rng_int_uniform(hbt_rng *rng);
double hbt_rng_uniform(hbt_rng *rng);
/* One batch worth of draws in the reference's order (src/HBT_correlation.cpp:202-217 and
 * :493-497): for each of the nev events, nmix = nev_mixed/2+1 partner ids
 * (rand_int_uniform() % nev_mixed, redrawn while == iev unless nev_mixed == 1), then nmix
 * rotation angles rand_uniform()*2*M_PI.  Outputs (any may be NULL): partner_ids
 * [nev*nmix], cos_sin [nev*nmix*2] = glibc cos/sin of the angle, angles [nev*nmix].
 * With all outputs NULL this just fast-forwards the stream past the batch (sharded runs).
 * Returns nmix. */
int32_t hbt_rng_mixed_plan(hbt_rng *rng, int32_t nev, int32_t nev_mixed, int32_t *partner_ids,
                           double *cos_sin, double *angles);

/* ---- fast reader: particle samples -> batches ("oversample groups") ------------------ */
/* Replaces, for read_in_mode = 10 (results/particle_samples.gz, gzipped iSS text), 2
 * (results/particle_list.dat, gzipped UrQMD text), 21 (results/particle_list.bin, UrQMD binary), 1
 * (results/particle_list.dat, UrQMD file-13 style text), 0 (results/OSCAR.DAT, OSCAR1997A), 9
 * (results/particle_list.bin, iSS binary), 7 (gzipped SMASH text), 4 / 3 (UrQMD 3.3p / header-less UrQMD text)
 * 5 (JAM text; all three results/particle_list.dat) and 8 (results/particles_binary.bin, extended SMASH binary)
 * — every read_in_mode of the reference —,
 * the reference's reader and the steps between it and the pair loops:
 * read_in_particle_samples_gzipped / _UrQMD_zipped / _UrQMD_binary / _UrQMD / _OSCAR + gz_readline
 * (src/particleSamples.cpp:1247-1286, :910-974, :976-1059, :838-908, :680-714, :2209-2218), the UrQMD id map
 * (:325-357, :389-400; unknown ids are dropped but counted), boostParticles (:441-470), filter_particles for
 * a single species (:625-678, :1329-1346) and, when rapidity_cut is not NULL, the HBT gather's
 * tanh(HBTrap_min) < pz/E < tanh(HBTrap_max) (src/HBT_correlation.cpp:255-281).  Same grouping rule
 * (events are appended while the particle count of all species is below event_buffer_size), same
 * doubles, same order.  A background thread inflates and parses up to two batches ahead. */
typedef struct hbt_reader hbt_reader;
int hbt_reader_open(const char *path, int32_t read_in_mode, int32_t particle_monval, int64_t event_buffer_size,
                    double rap_shift, const hbt_params *rapidity_cut, hbt_reader **out);
/* Next batch: returns its number of events (0 at the end of the file, < 0 on error); *particles
 * (8 doubles each: px py pz E x y z t) and *offsets (nev+1, in particles) stay valid until the
 * next call; *all_particles = particles of every species read for the batch. */
int32_t hbt_reader_next(hbt_reader *reader, const double **particles, const int64_t **offsets, int64_t *all_particles);
const char *hbt_reader_error(const hbt_reader *reader);
uint64_t hbt_reader_bytes(const hbt_reader *reader); /* inflated bytes consumed so far */
void hbt_reader_close(hbt_reader *reader);

/* ---- the hot path: HOST buffers in (copies are part of the call) ------------------ */
/* Same-event pair loop (src/HBT_correlation.cpp:291-460) over the merged, rapidity-cut
 * particle list `p` (n x 8, reference gather order).  psi_ref is only read when
 * azimuthal_flag = 1. */
int hbt_accumulate_same(hbt_ctx *ctx, const double *p, int64_t n, double psi_ref);
/* Mixed-event pair loops (:563-689) for ALL events of a batch in one submission.
 * p1/off1: rapidity-cut particles of the batch's events, flat, with nev1+1 offsets (list 1
 * of each call of the reference routine, :471-489).  p2/off2: rapidity-cut particles of the
 * mixed-event lists (NULL/0: they alias p1/off1, src/particleSamples.cpp:528-530).
 * partner_ids [nev1*nmix] index events of list 2; cos_sin [nev1*nmix*2] are the
 * host-computed cos and sin of each partner's rotation angle; the rotation (:522-523) is
 * applied on the device with the reference's two-multiply-one-add rounding. */
int hbt_accumulate_mixed(hbt_ctx *ctx, const double *p1, const int64_t *off1, int32_t nev1,
                         const double *p2, const int64_t *off2, int32_t nev2,
                         const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                         double psi_ref);
/* One whole batch = body of calculate_HBT_correlation_function (:177-218): same-event loop
 * over the concatenation of the events of list 1, then the mixed-event loops; list 1 is
 * uploaded once.  do_same / do_mixed select the halves. */
int hbt_accumulate_batch(hbt_ctx *ctx, const double *p1, const int64_t *off1, int32_t nev1,
                         const double *p2, const int64_t *off2, int32_t nev2,
                         const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                         double psi_ref, int32_t do_same, int32_t do_mixed);

/* ---- the hot path: particles already resident in HBM ------------------------------ */
/* Same contracts, but d_p / d_p1 / d_p2 are DEVICE pointers on the context's device
 * (offsets, ids and cos_sin stay host pointers: they are the host-side plan).  The
 * device buffers must stay valid until hbt_synchronize. */
int hbt_accumulate_same_dev(hbt_ctx *ctx, const double *d_p, int64_t n, double psi_ref);
int hbt_accumulate_mixed_dev(hbt_ctx *ctx, const double *d_p1, const int64_t *off1, int32_t nev1,
                             const double *d_p2, const int64_t *off2, int32_t nev2,
                             const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                             double psi_ref);

/* One whole batch with list 2 = list 1 (the events of d_p are each other's mixing partners):
 * body of calculate_HBT_correlation_function (:177-218) on a device-resident list.  Both loops
 * are launched together (HBT_OPT_FUSE): the same-event kernel and the mixed-event kernel share
 * the SMs on two streams. */
int hbt_accumulate_batch_dev(hbt_ctx *ctx, const double *d_p, const int64_t *off, int32_t nev,
                             const int32_t *partner_ids, const double *cos_sin, int32_t nmix, double psi_ref);

/* ---- results ---------------------------------------------------------------------- */
int hbt_synchronize(hbt_ctx *ctx);
/* Copies the accumulators out (any pointer may be NULL).  Sizes: hbt_num_bins for the six
 * histograms, hbt_num_slabs for the two accepted-pair counters — the reference's
 * correl_3d_num_count / correl_3d_num / q_out_mean / q_side_mean / q_long_mean /
 * correl_3d_denorm (or the *_Kphi_diff_* set) and number_of_pairs_numerator_KTdiff /
 * number_of_pairs_denormenator_KTdiff (or the *_KTKphidiff set)
 * (src/HBT_correlation.h:42-61).  Counts are exact integers (the reference keeps them in
 * doubles).  After hbt_allreduce every rank reads the global sums. */
int hbt_read(hbt_ctx *ctx, uint64_t *num_count, double *num_cos, double *sum_qo, double *sum_qs,
             double *sum_ql, uint64_t *den_count, uint64_t *npairs_num, uint64_t *npairs_den);
/* invariant_radius_flag = 1 accumulators, each [n_KT][qnpts] (+ [n_KT] counters):
 * correl_1d_inv_num_count, q_inv_mean, correl_1d_inv_num, correl_1d_inv_denorm,
 * number_of_pairs_{numerator,denormenator}_KTdiff_qinv_ */
int hbt_read_qinv(hbt_ctx *ctx, uint64_t *count, double *sum_qinv, double *sum_cos, uint64_t *den,
                  uint64_t *npairs_num, uint64_t *npairs_den);
/* stage populations {all pairs, passed K_T cut, passed q_out, passed q_side, passed q_long,
 * accepted}: the n_A..n_E of the roofline formula.  [1..3] need
 * HBT_OPT_STAGE_COUNTERS (see hbt_set_option). */
int hbt_get_stage_counters(hbt_ctx *ctx, uint64_t same[6], uint64_t mixed[6]);
/* device time of the pair kernels so far (CUDA events on the launching streams; launches of the two
 * compute lanes that overlap are counted once: the time during which at least one pair launch was
 * running, sort / cull helpers included), and the number of pair-kernel launches */
int hbt_get_timers(hbt_ctx *ctx, double *same_ms, double *mixed_ms, uint64_t *same_launches,
                   uint64_t *mixed_launches);
/* Engine options (set between batches; the call synchronises the context).
 * HBT_OPT_STAGE_COUNTERS = 1: instrumented runs — every pair goes through the prefilter and the
 *   stage populations "passed K_T / q_out / q_side" (hbt_get_stage_counters [1..3]) are exact.
 *   0 (default): production — the same-event list is sorted in momentum space and tile pairs that
 *   cannot hold an accepted pair are skipped, so those three populations are not available
 *   (reported as 0); all pairs [0], passed q_long [4] and accepted [5] stay exact, and so does
 *   every histogram.  Default can be changed with the environment variable HBT_B200_STATS=1.
 * HBT_OPT_KERNEL: 2 = tuned kernels (default), 1 = literal kernels (cross-check).
 * HBT_OPT_FUSE: 1 (default) = a whole batch (hbt_accumulate_batch with both halves) is ONE
 *   launch group: the same-event kernel and the mixed-event kernel (hbt_pairs_v4_mixed) run next
 *   to each other on two streams, each with its own registers and shared memory; 0 = one loop
 *   after the other.  Environment: HBT_B200_FUSE.  hbt_get_timers splits the time of such a
 *   launch group between same_ms and mixed_ms by the pairs of each kind.  (Environment only:
 *   HBT_B200_SPLIT=0 = the single fused kernel of earlier versions, HBT_B200_CORUN=0 = the two
 *   kernels one after the other, HBT_B200_CORUN_SAME / HBT_B200_CORUN_MIXED = resident warps per
 *   SM of each during the co-run, HBT_B200_MIXED4=0 = mixed-event loops on the v3 kernel.)
 * HBT_OPT_LANES: 2 (default) = consecutive production batches go to two compute streams in turn,
 *   each with its own scratch, so the sort / cull helpers and the first units of batch k+1 run
 *   under the tail of batch k (batches commute: every accumulation is an atomic add into the
 *   context's histograms); 1 = one stream.  Instrumented runs, the literal kernels and batches
 *   near the pair cap always use one stream.  Environment: HBT_B200_LANES.  hbt_get_timers counts
 *   the time during which at least one pair launch was running (overlaps once).
 * HBT_OPT_PTSORT: the production mixed-event loops can read a copy of the batch in which every event
 *   is sorted by pT (one radix sort per batch; pT is invariant under the partner rotation).  Since
 *   |q_out| >= |pT_i - pT_j|, a work unit then only visits the stretch of its list-2 tile whose pT
 *   lies within the q_out window of its list-1 particles (~40 % fewer pairs to pre-screen on the
 *   benchmark sample).  1 (default) = sort batches with at least 5e8 mixed-event pairs, 2 = always,
 *   0 = never.  Results do not depend on it.  Environment: HBT_B200_PTSORT.
 * HBT_OPT_COALESCE: 1 (default) = small production batches (fewer than 1.5e9 pairs, list 2 = list 1, no K_phi bins,
 *   far from the pair cap) are not launched at once: they wait in the context (host particles are staged at the call,
 *   device-resident ones must stay valid until hbt_synchronize, as documented) and up to 32 of them go out as ONE
 *   launch — their same-event lists sorted side by side, each padded to whole work units, mixed-event segments
 *   concatenated.  An oversample group of 10 events holds ~0.3 ms of pair work, about what its ~8 helper launches and
 *   the ramp / tail of a persistent kernel cost.  hbt_synchronize, hbt_read*, hbt_timer_stop and any batch that does
 *   not qualify flush what waits.  0 = one launch per batch.  Environment: HBT_B200_COALESCE.  Results do not depend
 *   on it (every accumulation is an atomic add). */
#define HBT_OPT_STAGE_COUNTERS 1
#define HBT_OPT_KERNEL 2
#define HBT_OPT_FUSE 3
#define HBT_OPT_LANES 4
#define HBT_OPT_PTSORT 5
#define HBT_OPT_COALESCE 6
int hbt_set_option(hbt_ctx *ctx, int32_t option, int32_t value);

/* Device-side stopwatch on the context's compute stream (CUDA events): everything the
 * context enqueues between start and stop — copies it waits for, pair kernels, the NCCL
 * all-reduce — is inside.  hbt_timer_stop waits for the stream and returns milliseconds. */
int hbt_timer_start(hbt_ctx *ctx);
int hbt_timer_stop(hbt_ctx *ctx, double *ms);
/* number of CUDA kernels this context has launched so far (pair kernels + bookkeeping) */
int hbt_get_launch_count(hbt_ctx *ctx, uint64_t *n);
/* pairs that the device deferred to the host's literal re-evaluation so far (K_phi bin
 * decisions within 1e-9 of an edge: CUDA atan2 vs glibc atan2) */
int hbt_get_deferred_pairs(hbt_ctx *ctx, uint64_t *n);

/* ---- multi-GPU: one context per GPU, one NCCL all-reduce of the histograms -------- */
/* multi-process (one rank per GPU): rank 0 makes an id, every rank joins */
int hbt_comm_unique_id(char id[128]);
int hbt_comm_init_rank(hbt_ctx *ctx, int32_t nranks, int32_t rank, const char id[128]);
/* single-process: ctxs[0..n) on distinct devices become one communicator */
int hbt_comm_init_all(hbt_ctx **ctxs, int32_t n);
/* sum all accumulators (u64 counts, f64 sums, counters) across the communicator over
 * NVLink; for a single-process group call it once with all contexts via hbt_allreduce_all */
int hbt_allreduce(hbt_ctx *ctx);
int hbt_allreduce_all(hbt_ctx **ctxs, int32_t n);

/* ---- one analysis over several GPUs of ONE process: a group of contexts ------------------------------
 * What the drop-in class does when HBT_B200_DEVICES > 1 (host/HBT_correlation.cpp): oversample groups (batches) go to
 * the GPUs in turn — they are independent of each other (src/Analysis.cpp:821-833) except for ONE coupling, the
 * needed_number_of_pairs cap, which is cumulative over the batches in order (src/HBT_correlation.cpp:402-406,
 * :424-428, :651-655, :673-677).  hbt_group_accumulate_batch keeps that exact: while no channel (K_T[,K_phi] slab, q_inv
 * histogram) can reach its quota with the pairs submitted so far, batches run asynchronously on their GPUs; when a
 * batch could close a channel, all contexts are synchronised, their exact per-channel counters are summed, and the
 * batch runs alone on its GPU starting from the other contexts' counts (hbt_cap_set_foreign), so that its ordered
 * replay stops at the very pair the reference stops at.  The reference's default needed_number_of_pairs = 3e7
 * therefore gives the reference's files on any number of GPUs.
 * devices == NULL: devices 0 .. n_devices-1.  A device may be listed more than once (several contexts on one GPU:
 * used by the tests on single-GPU boxes); the final sum is one NCCL all-reduce when the devices are distinct. */
typedef struct hbt_group hbt_group;
int hbt_group_create(const hbt_params *params, int32_t n_devices, const int32_t *devices, hbt_group **out);
void hbt_group_destroy(hbt_group *group);
const char *hbt_group_last_error(const hbt_group *group); /* group == NULL: of the last failed hbt_group_create */
int32_t hbt_group_size(const hbt_group *group);
hbt_ctx *hbt_group_ctx(hbt_group *group, int32_t i);
/* same arguments as hbt_accumulate_batch; the batch goes to the next context in turn */
int hbt_group_accumulate_batch(hbt_group *group, const double *p1, const int64_t *off1, int32_t nev1,
                               const double *p2, const int64_t *off2, int32_t nev2,
                               const int32_t *partner_ids, const double *cos_sin, int32_t nmix,
                               double psi_ref, int32_t do_same, int32_t do_mixed);
/* sums the accumulators of all contexts; afterwards hbt_read / hbt_read_qinv / hbt_get_stage_counters on
 * hbt_group_ctx(group, 0) return the totals (until the next accumulate) */
int hbt_group_reduce(hbt_group *group);
/* batches that had to run in sequence because the cap could engage */
int hbt_group_ordered_batches(const hbt_group *group, uint64_t *n);
/* building blocks of the above, for hosts that drive their contexts themselves (one process per GPU): the cap
 * channels are the slabs, followed by the n_KT q_inv histograms when invariant_radius_flag = 1.
 * hbt_cap_get_counts synchronises and returns this context's own accepted pairs per channel (numerator /
 * denominator); hbt_cap_set_foreign tells it how many pairs the OTHER contexts hold. */
int32_t hbt_cap_channels(const hbt_ctx *ctx);
int hbt_cap_get_counts(hbt_ctx *ctx, uint64_t *num, uint64_t *den);
int hbt_cap_set_foreign(hbt_ctx *ctx, const uint64_t *num, const uint64_t *den);

/* FP64-pipe roofline denominator: runs a dependent-chain DFMA microbenchmark on `device`
 * for about `ms` milliseconds and returns the sustained rate in TFLOP/s (2 flops per DFMA).
 * Measurement aid for bench.py; not part of the reference's interface. */
int hbt_measure_fp64_peak(int32_t device, double ms, double *tflops);

/* ---- next row: the pair loops of class BalanceFunction ------- */
/* BalanceFunction::combine_and_bin_particle_pairs (src/BalanceFunction.cpp:120-157) and
 * combine_and_bin_mixed_particle_pairs (:159-197): per event, every particle of list a against every
 * particle of list b of the partner event, histogram [Bnpts][20] in (Delta y, Delta phi).  The host keeps
 * what the reference computes once per particle with glibc (phi_p = atan2(py,px), rap_y / rap_eta =
 * asinh(...), src/particleSamples.cpp:441-470), the pT cut (:127,129) and the RNG draws (:167-170); the
 * device does the O(n_a n_b) loop with the reference's own IEEE operations (subtract, divide, floor,
 * int cast), so every bin count is the reference's.  Grid constants as in the constructor (:27-36):
 * Bnphi = 20, dphi = 2 pi/20, Bphi_min = -pi/2, drap = 2|Brap_max|/(Bnpts-1), Brap_min = -|Brap_max| - drap/2. */
typedef struct hbt_bf hbt_bf;
#define HBT_BF_NPHI 20
#define HBT_BF_NHIST 8 /* C_ab, C_abarbbar, C_abbar, C_abarb, then the four mixed-event ones */
int hbt_bf_create(int32_t Bnpts, double Brap_max, int32_t device, hbt_bf **out);
void hbt_bf_destroy(hbt_bf *bf);
const char *hbt_bf_last_error(const hbt_bf *bf);
/* One call of either routine for all nev events of a batch.  a / b: (phi_p, rapidity) pairs of the
 * particles that passed the pT cut, flat, with nev+1 / nev_b+1 offsets (in particles); partner[iev] =
 * event of list b that event iev of list a is paired with (iev itself for the same-event routine, the
 * drawn iev_mixed for the mixed one); rotation[iev] is added to Delta phi as the reference adds its
 * global_random_rotation (pass 0.0 for the same-event routine: x + 0.0 is x).  hist in [0, 8). */
int hbt_bf_accumulate(hbt_bf *bf, int32_t hist, const double *a, const int64_t *off_a, int32_t nev,
                      const double *b, const int64_t *off_b, int32_t nev_b, const int32_t *partner,
                      const double *rotation);
/* counts [8][Bnpts][20] (the reference keeps them in doubles; they are integers) */
int hbt_bf_read(hbt_bf *bf, uint64_t *hist);
int hbt_bf_get_timers(hbt_bf *bf, double *kernel_ms, uint64_t *pairs);

/* number of CUDA devices visible to the process (0 when there is none) */
int32_t hbt_device_count(void);

/* library self-description: "hbt_b200 <version> sm_100a" */
const char *hbt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HBT_B200_H_ */
