// Drop-in replacement of the reference's class HBT_correlation
// (/root/reference/src/HBT_correlation.h:14-104): same include guard, same constructor,
// same public methods, same parameters.dat switches, same output files — the two pair loops
// run on B200 GPUs through the C ABI of libhbt_b200.so (include/hbt_b200.h).
//
// It is compiled against the reference's own headers (ParameterReader, Random,
// particleSamples, pretty_ostream), which stay where they are; the reference TUs that use the
// class (src/Analysis.cpp, unit_tests/HBT_unittest.cc) are compiled through one-line wrapper
// TUs that include THIS header first, so the shared include guard turns the reference's
// header into a no-op.
//
// Not carried over: the public create_a_*D_array / delete_a_*D_array template helpers
// (src/HBT_correlation.h:87-103) — the histograms are flat arrays in HBM, nothing calls them.
#ifndef HBT_correlation_h
#define HBT_correlation_h

#include <memory>
#include <string>
#include <vector>

#include "ParameterReader.h"
#include "Random.h"
#include "arsenal.h"
#include "particleSamples.h"
#include "pretty_ostream.h"

#include "hbt_output.h"

class HBT_correlation {
  private:
    const ParameterReader paraRdr_;
    const std::string path_;
    std::shared_ptr<particleSamples> particle_list;
    std::shared_ptr<RandomUtil::Random> ran_gen_ptr_;
    pretty_ostream messager;

    // the switches of src/HBT_correlation.cpp:22-46
    bool long_comoving_boost;
    int qnpts;
    double q_min, q_max, delta_q;
    int azimuthal_flag_;
    int invariant_radius_flag_;
    double psi_ref;
    int n_KT, n_Kphi;
    double dKT, dKphi;
    double KT_min, KT_max;
    double Krap_min_, Krap_max_;
    std::vector<double> KT_array_, Kphi_array_;
    std::vector<double> q_out, q_side, q_long;
    int number_of_mixed_events_;
    int number_of_oversample_events_;
    unsigned long long int needed_number_of_pairs;

    // one engine context per GPU (HBT_B200_DEVICES), as a group of libhbt_b200: batches are dealt round-robin, the
    // ordered pair cap stays exact across the GPUs, histograms are summed once (NCCL all-reduce over NVLink)
    // before the output is written
    hbt_group *group_;

    // host copies of the accumulators, filled by fetch_results()
    hbt_params params_;
    HbtHostResults res_;
    bool fetched_;

    // scratch of the gathers
    std::vector<double> gather1_, gather2_;
    std::vector<long long> off1_, off2_;

    void check(hbt_ctx *ctx, int rc, const char *what);
    hbt_ctx *first_context();
    long long gather_events(bool mixed_list, const std::vector<int> &events, std::vector<double> &out,
                            std::vector<long long> &offsets);
    void fetch_results();
    HbtOutputWriter writer();

  public:
    HBT_correlation(ParameterReader &paraRdr, std::string path, std::shared_ptr<RandomUtil::Random> ran_gen);
    ~HBT_correlation();
    HBT_correlation(const HBT_correlation &) = delete;
    HBT_correlation &operator=(const HBT_correlation &) = delete;

    double get_psi_ref() { return (psi_ref); };

    void set_particle_list(std::shared_ptr<particleSamples> particle_list_in) { particle_list = particle_list_in; }

    void calculate_flow_event_plane_angle(int n_order);
    void calculate_HBT_correlation_function(std::shared_ptr<particleSamples> particle_list_in);
    void combine_and_bin_particle_pairs(std::vector<int> event_list);
    void combine_and_bin_particle_pairs_mixed_events(int event_id, std::vector<int> mixed_event_list);

    void output_HBTcorrelation();
    void output_correlation_function_inv();
    void output_correlation_function();
    void output_correlation_function_Kphi_differential();
};

#endif
