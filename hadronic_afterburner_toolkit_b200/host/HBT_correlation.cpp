// Host side of the drop-in HBT_correlation class: parameters, gathers, RNG replay, logging and
// the output writers stay on the CPU and produce the same doubles / the same files as the
// reference (/root/reference/src/HBT_correlation.cpp); the two pair loops are submitted to
// libhbt_b200.so.  There is no CPU implementation of the pair loops here: if the library
// reports an error the program stops with a message, like the reference does for its own
// fatal conditions (message + exit(1)).
#include "HBT_correlation.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "../../include/hbt_b200.h"
#include "hbt_output.h"

namespace {

// glibc / iostream print a NaN produced by 0.0/0.0 as "-nan"; nothing to do, ostream<<double
// does the same here because the same expressions are evaluated.

// HBT_B200_DEVICES: "all", a count ("4": devices 0..3) or an explicit list ("0,2,3"; a device may repeat)
std::vector<int32_t> devices_from_env() {
    std::vector<int32_t> devs;
    const char *e = std::getenv("HBT_B200_DEVICES");
    if (!e || !*e) return std::vector<int32_t>(1, 0);
    const std::string v(e);
    if (v == "all") {
        for (int d = 0; d < hbt_device_count(); d++) devs.push_back(d);
    } else if (v.find(',') != std::string::npos) {
        std::stringstream ss(v);
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) devs.push_back(std::atoi(tok.c_str()));
    } else {
        for (int d = 0; d < std::max(1, std::atoi(e)); d++) devs.push_back(d);
    }
    if (devs.empty()) devs.push_back(0);
    return devs;
}

}  // namespace

HBT_correlation::HBT_correlation(ParameterReader &paraRdr, std::string path,
                                 std::shared_ptr<RandomUtil::Random> ran_gen)
    : paraRdr_(paraRdr), path_(path), group_(nullptr), fetched_(false) {
    ran_gen_ptr_ = ran_gen;

    // same keys, same order as src/HBT_correlation.cpp:22-46 (a missing key exits in getVal)
    long_comoving_boost = (paraRdr_.getVal("long_comoving_boost") == 1);
    qnpts = paraRdr_.getVal("qnpts");
    q_min = paraRdr_.getVal("q_min");
    q_max = paraRdr_.getVal("q_max");
    delta_q = (q_max - q_min) / (qnpts - 1);
    for (int i = 0; i < qnpts; i++) {
        const double q = q_min + i * delta_q;
        q_out.push_back(q);
        q_side.push_back(q);
        q_long.push_back(q);
    }
    needed_number_of_pairs = paraRdr_.getVal("needed_number_of_pairs");
    azimuthal_flag_ = paraRdr_.getVal("azimuthal_flag");
    invariant_radius_flag_ = paraRdr_.getVal("invariant_radius_flag");
    n_KT = paraRdr_.getVal("n_KT");
    n_Kphi = paraRdr_.getVal("n_Kphi");
    KT_min = paraRdr_.getVal("KT_min");
    KT_max = paraRdr_.getVal("KT_max");
    Krap_min_ = paraRdr_.getVal("HBTrap_min");
    Krap_max_ = paraRdr_.getVal("HBTrap_max");
    dKT = (KT_max - KT_min) / (n_KT - 1);
    dKphi = 2 * M_PI / n_Kphi;
    for (int i = 0; i < n_KT; i++) KT_array_.push_back(KT_min + i * dKT);
    for (int i = 0; i < n_Kphi; i++) Kphi_array_.push_back(i * dKphi);
    psi_ref = 0.;
    number_of_mixed_events_ = 0;
    number_of_oversample_events_ = 0;

    hbt_params &p = params_;
    p.qnpts = qnpts;
    p.n_KT = n_KT;
    p.n_Kphi = n_Kphi;
    p.azimuthal_flag = azimuthal_flag_;
    p.invariant_radius_flag = invariant_radius_flag_;
    p.long_comoving_boost = long_comoving_boost ? 1 : 0;
    p.q_min = q_min;
    p.q_max = q_max;
    p.KT_min = KT_min;
    p.KT_max = KT_max;
    p.HBTrap_min = Krap_min_;
    p.HBTrap_max = Krap_max_;
    p.needed_number_of_pairs = paraRdr_.getVal("needed_number_of_pairs");

    // One engine context per GPU, as a group: oversample groups go to the GPUs in turn, and the group keeps the
    // ordered needed_number_of_pairs cap (src/HBT_correlation.cpp:402-406) exact across them.
    const std::vector<int32_t> devs = devices_from_env();
    const int rc = hbt_group_create(&p, static_cast<int32_t>(devs.size()), devs.data(), &group_);
    if (rc != HBT_OK) {
        messager << "HBT_correlation: cannot create the GPU engine: " << hbt_group_last_error(nullptr) << hbt_last_error(nullptr);
        messager.flush("error");
        exit(1);
    }
    messager << "HBT pair loops run on " << devs.size() << " GPU context(s) [" << hbt_version() << "]";
    messager.flush("info");
}

HBT_correlation::~HBT_correlation() {
    hbt_group_destroy(group_);
}

void HBT_correlation::check(hbt_ctx *ctx, int rc, const char *what) {
    if (rc == HBT_OK) return;
    messager << "HBT_correlation: " << what << " failed: " << hbt_last_error(ctx);
    messager.flush("error");
    exit(1);
}

// the per-method interface (the reference's unit test drives it) stays on the first context
hbt_ctx *HBT_correlation::first_context() {
    fetched_ = false;
    return hbt_group_ctx(group_, 0);
}

//! Psi_n of the batch, src/HBT_correlation.cpp:233-249 (all filtered particles, no rapidity cut)
void HBT_correlation::calculate_flow_event_plane_angle(int n_order) {
    const int nev = particle_list->get_number_of_events();
    double vn_real = 0.0;
    double vn_imag = 0.0;
    for (int iev = 0; iev < nev; iev++) {
        const int npart = particle_list->get_number_of_particles(iev);
        for (int i = 0; i < npart; i++) {
            const particle_info part = particle_list->get_particle(iev, i);
            const double phi = atan2(part.py, part.px);
            vn_real += cos(n_order * phi);
            vn_imag += sin(n_order * phi);
        }
    }
    psi_ref = atan2(vn_imag, vn_real) / n_order;
}

//! Rapidity-cut copy of the listed events (src/HBT_correlation.cpp:255-281, :468-489,
//! :499-511) as 8 doubles per particle (px,py,pz,E,x,y,z,t) plus per-event offsets.
long long HBT_correlation::gather_events(bool mixed_list, const std::vector<int> &events,
                                         std::vector<double> &out, std::vector<long long> &offsets) {
    const double cut_max = tanh(Krap_max_);
    const double cut_min = tanh(Krap_min_);
    out.clear();
    offsets.assign(1, 0);
    for (int ev : events) {
        const int npart = mixed_list ? particle_list->get_number_of_particles_mixed_event(ev)
                                     : particle_list->get_number_of_particles(ev);
        for (int i = 0; i < npart; i++) {
            const particle_info part = mixed_list ? particle_list->get_particle_from_mixed_event(ev, i)
                                                  : particle_list->get_particle(ev, i);
            const double ratio = part.pz / part.E;
            if (ratio > cut_min && ratio < cut_max) {
                const double rec[8] = {part.px, part.py, part.pz, part.E, part.x, part.y, part.z, part.t};
                out.insert(out.end(), rec, rec + 8);
            }
        }
        offsets.push_back(static_cast<long long>(out.size() / 8));
    }
    return offsets.back();
}

//! One batch: src/HBT_correlation.cpp:177-218, submitted as one library call
void HBT_correlation::calculate_HBT_correlation_function(std::shared_ptr<particleSamples> particle_list_in) {
    set_particle_list(particle_list_in);
    const int nev = particle_list->get_number_of_events();
    if (azimuthal_flag_ == 1) calculate_flow_event_plane_angle(2);

    number_of_oversample_events_ = nev;
    messager.info("Compute pairs from the same event ...");
    std::vector<int> all(nev);
    for (int i = 0; i < nev; i++) all[i] = i;
    const long long n1 = gather_events(false, all, gather1_, off1_);
    const unsigned long long same_pairs = n1 > 0 ? static_cast<unsigned long long>(n1) * (n1 - 1) / 2 : 0;
    messager << "number of pairs: " << same_pairs;
    messager.flush("info");

    const int mixed_nev = particle_list->get_number_of_mixed_events();
    messager << "nev_mixed = " << mixed_nev;
    messager.flush("info");
    number_of_mixed_events_ = static_cast<int>(mixed_nev / 2) + 1;
    messager.info("Compute pairs from the mixed event ...");

    const bool real_mixed = (paraRdr_.getVal("read_in_real_mixed_events") == 1);
    const double *p2 = nullptr;
    const long long *o2 = nullptr;
    if (real_mixed) {
        std::vector<int> allm(mixed_nev);
        for (int i = 0; i < mixed_nev; i++) allm[i] = i;
        gather_events(true, allm, gather2_, off2_);
        p2 = gather2_.data();
        o2 = off2_.data();
    }
    const std::vector<long long> &offm = real_mixed ? off2_ : off1_;

    // the reference's draws, in its order: per event the partner ids (:206-215), then one
    // rotation angle per partner inside the mixed routine (:493-497)
    const int nmix = number_of_mixed_events_;
    std::vector<int> ids(static_cast<size_t>(nev) * nmix);
    std::vector<double> cs(static_cast<size_t>(nev) * nmix * 2);
    const bool do_mixed = nev > 0 && mixed_nev > 0;
    if (do_mixed) {
        for (int iev = 0; iev < nev; iev++) {
            messager << "progess: " << iev << "/" << nev;
            messager.flush("info");
            for (int c = 0; c < nmix; c++) {
                int id = (ran_gen_ptr_->rand_int_uniform() % mixed_nev);
                while (iev == id && mixed_nev != 1) id = (ran_gen_ptr_->rand_int_uniform() % mixed_nev);
                ids[static_cast<size_t>(iev) * nmix + c] = id;
            }
            unsigned long long pairs = 0;
            for (int c = 0; c < nmix; c++) {
                const double random_rotation = ran_gen_ptr_->rand_uniform() * 2 * M_PI;
                const size_t k = static_cast<size_t>(iev) * nmix + c;
                cs[2 * k] = cos(random_rotation);
                cs[2 * k + 1] = sin(random_rotation);
                const int id = ids[k];
                pairs += static_cast<unsigned long long>(off1_[iev + 1] - off1_[iev]) * (offm[id + 1] - offm[id]);
            }
            messager << "number of mixed pairs: " << pairs;
            messager.flush("info");
        }
    }
    if (nev == 0) return;  // the reader's trailing empty batch: nothing to do, no draws
    fetched_ = false;
    if (hbt_group_accumulate_batch(group_, gather1_.data(), reinterpret_cast<const int64_t *>(off1_.data()), nev, p2,
                                   reinterpret_cast<const int64_t *>(o2), real_mixed ? mixed_nev : 0, ids.data(), cs.data(),
                                   do_mixed ? nmix : 0, psi_ref, 1, do_mixed ? 1 : 0) != HBT_OK) {
        messager << "HBT_correlation: hbt_group_accumulate_batch failed: " << hbt_group_last_error(group_);
        messager.flush("error");
        exit(1);
    }
}

//! Same-event pairs of the listed events merged into one list, src/HBT_correlation.cpp:251-462
void HBT_correlation::combine_and_bin_particle_pairs(std::vector<int> event_list) {
    const long long n = gather_events(false, event_list, gather1_, off1_);
    messager << "number of pairs: " << (n > 0 ? static_cast<unsigned long long>(n) * (n - 1) / 2 : 0);
    messager.flush("info");
    hbt_ctx *c = first_context();
    check(c, hbt_accumulate_same(c, gather1_.data(), n, psi_ref), "hbt_accumulate_same");
}

//! One event against the listed partner events, each rotated by a fresh random angle,
//! src/HBT_correlation.cpp:464-692
void HBT_correlation::combine_and_bin_particle_pairs_mixed_events(int event_id1, std::vector<int> mixed_event_list) {
    number_of_mixed_events_ = static_cast<int>(mixed_event_list.size());
    const long long n1 = gather_events(false, std::vector<int>(1, event_id1), gather1_, off1_);
    // list 2 is gathered in partner order, one segment per partner (duplicates allowed)
    gather_events(true, mixed_event_list, gather2_, off2_);
    const int nmix = number_of_mixed_events_;
    std::vector<int> ids(nmix);
    std::vector<double> cs(static_cast<size_t>(nmix) * 2);
    for (int c = 0; c < nmix; c++) {
        const double random_rotation = ran_gen_ptr_->rand_uniform() * 2 * M_PI;
        cs[2 * c] = cos(random_rotation);
        cs[2 * c + 1] = sin(random_rotation);
        ids[c] = c;
    }
    messager << "number of mixed pairs: " << static_cast<unsigned long long>(n1) * off2_.back();
    messager.flush("info");
    hbt_ctx *c = first_context();
    check(c,
          hbt_accumulate_mixed(c, gather1_.data(), reinterpret_cast<const int64_t *>(off1_.data()), 1, gather2_.data(),
                               reinterpret_cast<const int64_t *>(off2_.data()), nmix, ids.data(), cs.data(), nmix,
                               psi_ref),
          "hbt_accumulate_mixed");
}

void HBT_correlation::fetch_results() {
    if (hbt_group_reduce(group_) != HBT_OK) {
        messager << "HBT_correlation: hbt_group_reduce failed: " << hbt_group_last_error(group_);
        messager.flush("error");
        exit(1);
    }
    hbt_ctx *c0 = hbt_group_ctx(group_, 0);
    check(c0, hbt_fetch_results(c0, params_, res_), "hbt_read");
    fetched_ = true;
}

// the writers themselves (format frozen by src/HBT_correlation.cpp:694-855) live in hbt_output.h
void HBT_correlation::output_HBTcorrelation() {
    fetch_results();
    if (invariant_radius_flag_ == 1) output_correlation_function_inv();
    if (azimuthal_flag_ == 0) {
        output_correlation_function();
    } else {
        output_correlation_function_Kphi_differential();
    }
}

HbtOutputWriter HBT_correlation::writer() { return HbtOutputWriter(params_, path_, paraRdr_.getVal("ecoOutput", 0) == 1); }

void HBT_correlation::output_correlation_function_inv() {
    if (!fetched_) fetch_results();
    writer().write_inv(res_);
}

void HBT_correlation::output_correlation_function() {
    if (!fetched_) fetch_results();
    writer().write_KT(res_);
}

void HBT_correlation::output_correlation_function_Kphi_differential() {
    if (!fetched_) fetch_results();
    writer().write_KT_Kphi(res_);
}
