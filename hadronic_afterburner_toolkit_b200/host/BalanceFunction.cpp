// Host side of the drop-in BalanceFunction class: parameters, the per-particle pT cut, the RNG
// draws of the mixed-event routine and the output writer stay on the CPU and evaluate the
// reference's expressions (/root/reference/src/BalanceFunction.cpp); the pair loops are submitted to
// libhbt_b200.so (hbt_bf_accumulate).  No CPU implementation of the loops: a library error stops the
// program with a message, as the reference does for its own fatal conditions.
#include "BalanceFunction.h"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "../../include/hbt_b200.h"

using std::endl;

BalanceFunction::BalanceFunction(const ParameterReader &paraRdr, const std::string path,
                                 std::shared_ptr<RandomUtil::Random> ran_gen)
    : paraRdr_(paraRdr), path_(path), bf_(nullptr) {
    ran_gen_ptr = ran_gen;
    // same keys, same order, same expressions as src/BalanceFunction.cpp:20-38
    particle_monval_a = paraRdr_.getVal("particle_alpha");
    particle_monval_b = paraRdr_.getVal("particle_beta");
    same_species = (particle_monval_a == -particle_monval_b);
    BpT_min = paraRdr_.getVal("BpT_min");
    BpT_max = paraRdr_.getVal("BpT_max");
    Bnpts = paraRdr_.getVal("Bnpts");
    Brap_max = paraRdr_.getVal("Brap_max");
    drap = 2. * std::abs(Brap_max) / (Bnpts - 1);
    Brap_min = -std::abs(Brap_max) - 0.5 * drap;
    Bnphi = 20;
    dphi = 2. * M_PI / Bnphi;
    Bphi_min = -M_PI / 2.;
    rap_type_ = paraRdr_.getVal("rap_type");
    N_b = 0;
    N_bbar = 0;
    if (hbt_bf_create(Bnpts, Brap_max, 0, &bf_) != HBT_OK) {
        messager << "BalanceFunction: cannot create the GPU engine: " << hbt_bf_last_error(bf_);
        messager.flush("error");
        exit(1);
    }
    messager << "BalanceFunction pair loops run on the GPU [" << hbt_version() << "]";
    messager.flush("info");
}

BalanceFunction::~BalanceFunction() { hbt_bf_destroy(bf_); }

//! src/BalanceFunction.cpp:100-118 (not called by the reference either; kept for the interface)
bool BalanceFunction::check_same_particle(const particle_info &lhs, const particle_info &rhs) {
    const double tol = 1e-15;
    return lhs.monval == rhs.monval && std::abs(lhs.E - rhs.E) < tol && std::abs(lhs.px - rhs.px) < tol
           && std::abs(lhs.py - rhs.py) < tol && std::abs(lhs.t - rhs.t) < tol && std::abs(lhs.x - rhs.x) < tol
           && std::abs(lhs.y - rhs.y) < tol;
}

//! the pT cut of :127 / :129 and the rapidity choice of :141-143, once per particle
void BalanceFunction::gather(const plist_t *plist, Flat &out) {
    out.v.clear();
    out.off.assign(1, 0);
    for (auto const &ev : (*plist)) {
        for (auto const &part : (*ev)) {
            if (part.pT < BpT_min || part.pT > BpT_max) continue;
            out.v.push_back(part.phi_p);
            out.v.push_back(rap_type_ == 0 ? part.rap_eta : part.rap_y);
        }
        out.off.push_back(static_cast<long long>(out.v.size() / 2));
    }
}

void BalanceFunction::run(int hist, const Flat &a, const Flat &b, const std::vector<int> &partner,
                          const std::vector<double> &rotation) {
    const int rc = hbt_bf_accumulate(bf_, hist, a.v.data(), reinterpret_cast<const int64_t *>(a.off.data()),
                                     static_cast<int>(a.off.size()) - 1, b.v.data(),
                                     reinterpret_cast<const int64_t *>(b.off.data()), static_cast<int>(b.off.size()) - 1,
                                     partner.data(), rotation.data());
    if (rc != HBT_OK) {
        messager << "BalanceFunction: hbt_bf_accumulate failed: " << hbt_bf_last_error(bf_);
        messager.flush("error");
        exit(1);
    }
}

//! src/BalanceFunction.cpp:61-98: the eight calls in the reference's order
void BalanceFunction::calculate_balance_function(std::shared_ptr<particleSamples> particle_list_in) {
    set_particle_list(particle_list_in);
    auto plist_a = particle_list->get_balance_function_particle_list_a();
    auto plist_b = particle_list->get_balance_function_particle_list_b();
    auto plist_abar = particle_list->get_balance_function_particle_list_abar();
    auto plist_bbar = particle_list->get_balance_function_particle_list_bbar();

    N_b += get_number_of_particles(plist_b);
    N_bbar += get_number_of_particles(plist_bbar);

    Flat a, b, abar, bbar;
    gather(plist_a, a);
    gather(plist_b, b);
    gather(plist_abar, abar);
    gather(plist_bbar, bbar);
    const int nev = static_cast<int>(plist_a->size());
    std::vector<int> ident(nev);
    for (int i = 0; i < nev; i++) ident[i] = i;
    const std::vector<double> zero(nev, 0.0);
    messager.info("calculating C_ab ... ");
    run(0, a, b, ident, zero);
    messager.info("calculating C_abarbbar ... ");
    run(1, abar, bbar, ident, zero);
    messager.info("calculating C_abbar ... ");
    run(2, a, bbar, ident, zero);
    messager.info("calculating C_abarb ... ");
    run(3, abar, b, ident, zero);

    messager.info("calculating correlatoin function using mixed events ... ");
    // the mixed-event lists (they alias the batch unless real mixed events are read)
    Flat b_mixed, bbar_mixed;
    gather(particle_list->get_balance_function_particle_list_b_mixed_event(), b_mixed);
    gather(particle_list->get_balance_function_particle_list_bbar_mixed_event(), bbar_mixed);
    const Flat *lists_a[4] = {&a, &abar, &a, &abar};
    const Flat *lists_b[4] = {&b_mixed, &bbar_mixed, &bbar_mixed, &b_mixed};
    for (int h = 0; h < 4; h++) {
        // the draws of src/BalanceFunction.cpp:167-170, one partner and one rotation per event, in order
        const int nev_mixed = static_cast<int>(lists_b[h]->off.size()) - 1;
        std::vector<int> partner(nev);
        std::vector<double> rotation(nev);
        for (int iev = 0; iev < nev; iev++) {
            partner[iev] = (ran_gen_ptr.lock()->rand_int_uniform() % nev_mixed);
            rotation[iev] = (ran_gen_ptr.lock()->rand_uniform() * 2. * M_PI);
        }
        run(4 + h, *lists_a[h], *lists_b[h], partner, rotation);
    }
}

//! src/BalanceFunction.cpp:199-209
int BalanceFunction::get_number_of_particles(const std::vector<std::vector<particle_info> *> *plist_b) {
    int particle_number = 0;
    for (auto const &ev_i : (*plist_b)) {
        for (auto const &part_b : (*ev_i)) {
            if (part_b.pT < BpT_min || part_b.pT > BpT_max) continue;
            particle_number += 1;
        }
    }
    return (particle_number);
}

//! The three files of src/BalanceFunction.cpp:211-319 (names, headers, columns and number format
//! frozen).  The histograms come back as integers; every sum below is a sum of integers held in
//! doubles, exact in any order, so the tables are built once (opposite-sign = ab + abar-bbar,
//! same-sign = a-bbar + abar-b) and projected.
void BalanceFunction::output_balance_function() {
    std::vector<uint64_t> raw(static_cast<size_t>(HBT_BF_NHIST) * Bnpts * Bnphi);
    if (hbt_bf_read(bf_, raw.data()) != HBT_OK) {
        messager << "BalanceFunction: hbt_bf_read failed: " << hbt_bf_last_error(bf_);
        messager.flush("error");
        exit(1);
    }
    const size_t nbin = static_cast<size_t>(Bnpts) * Bnphi;
    // tab[0] = rho2(OS), tab[1] = rho1^2(OS) (mixed), tab[2] = rho2(SS), tab[3] = rho1^2(SS)
    std::vector<double> tab[4];
    const int first[4] = {0, 4, 2, 6};  // C_ab, C_mixed_ab, C_abbar, C_mixed_abbar; the partner histogram follows each
    for (int k = 0; k < 4; k++) {
        tab[k].resize(nbin);
        for (size_t b = 0; b < nbin; b++)
            tab[k][b] = static_cast<double>(raw[first[k] * nbin + b]) + static_cast<double>(raw[(first[k] + 1) * nbin + b]);
    }
    std::vector<double> along_y[4], along_phi[4];
    double total[4] = {0., 0., 0., 0.};
    for (int k = 0; k < 4; k++) {
        along_y[k].assign(Bnpts, 0.);
        along_phi[k].assign(Bnphi, 0.);
        for (int i = 0; i < Bnpts; i++)
            for (int j = 0; j < Bnphi; j++) {
                along_y[k][i] += tab[k][static_cast<size_t>(i) * Bnphi + j];
                along_phi[k][j] += tab[k][static_cast<size_t>(i) * Bnphi + j];
            }
        for (int i = 0; i < Bnpts; i++) total[k] += along_y[k][i];
    }
    const double N_OS = total[0], N_OS_mixed = total[1], N_SS = total[2], N_SS_mixed = total[3];
    std::vector<double> Delta_y(Bnpts), Delta_phi(Bnphi);
    for (int i = 0; i < Bnpts; i++) Delta_y[i] = Brap_min + (i + 0.5) * drap;
    for (int j = 0; j < Bnphi; j++) Delta_phi[j] = Bphi_min + (j + 0.5) * dphi;
    std::ostringstream stem;
    stem << particle_monval_a << "_" << particle_monval_b;

    // the two projections share their layout (:256-297)
    auto write_projection = [&](const std::string &name, const char *header, const std::vector<double> &axis,
                                const std::vector<double> *proj) {
        std::ofstream out((path_ + "/Balance_function_" + stem.str() + name).c_str(), std::ios::out);
        out << header << endl;
        for (size_t i = 0; i < axis.size(); i++) {
            const double C2_OS = proj[0][i] / proj[1][i] * N_OS_mixed / N_OS;
            const double C2_SS = proj[2][i] / proj[3][i] * N_SS_mixed / N_SS;
            out << std::scientific << std::setw(18) << std::setprecision(8) << axis[i] << "   " << C2_OS - C2_SS << "  " << C2_OS
                << "  " << proj[0][i] << "  " << proj[1][i] << "  " << C2_SS << "  " << proj[2][i] << "  " << proj[3][i] << endl;
        }
    };
    write_projection("_Delta_y.dat", "# DeltaY  Delta_C2  C2(OS)  rho2(OS)  rho1^2(OS)  C2(SS) rho2(SS)  rho1^2(SS)", Delta_y, along_y);
    write_projection("_Delta_phi.dat", "# Delta_phi  Delta_C2  C2(OS)  rho2(OS)  rho1^2(OS)  C2(SS) rho2(SS)  rho1^2(SS)", Delta_phi,
                     along_phi);

    // the 2-D table (:299-318)
    std::ofstream out2d((path_ + "/Correlation_function_" + stem.str() + "_2D.dat").c_str(), std::ios::out);
    out2d << "# DY  Dphi  C2(OS)  rho2(OS)  rho1^2(OS)  C2(SS)  rho2(SS)  rho1^2(SS)" << endl;
    for (int i = 0; i < Bnpts; i++)
        for (int j = 0; j < Bnphi; j++) {
            const size_t b = static_cast<size_t>(i) * Bnphi + j;
            const double C2_OS = (tab[0][b] / (tab[1][b] + 1e-15));
            const double C2_SS = (tab[2][b] / (tab[3][b] + 1e-15));
            out2d << std::scientific << std::setw(18) << std::setprecision(8) << Delta_y[i] << "  " << Delta_phi[j] << "  "
                  << C2_OS * N_OS_mixed / N_OS << "  " << tab[0][b] << "  " << tab[1][b] << "  " << C2_SS * N_SS_mixed / N_SS
                  << "  " << tab[2][b] << "  " << tab[3][b] << endl;
        }
}
