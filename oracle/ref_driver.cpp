// TEST INFRASTRUCTURE — not part of the product path.
//
// Driver around the UNMODIFIED reference library (compiled from /root/reference/src
// by oracle/Makefile into oracle/_ref/).  It feeds the reference's own
// HBT_correlation class (src/HBT_correlation.h:14-104) and dumps the RAW
// accumulators (the .dat text output only carries 9 significant digits, which is
// not enough to check the 1e-10 tolerance).  Private members are reached with
// `#define private public`; nothing of the reference is copied into this file.
//
// Modes
//   ref_driver mem   <params.dat> <batches.bin> <out.bin> [same_only | only=i,j,k]
//       (only=...: process just those batches; the draws of the others are replayed from the
//        shared Random object in the reference's order, so a process that owns a subset of the
//        groups sees exactly the stream positions of the single-process run)
//       batches.bin (format HBTIN001, see tests/hbtio.py) is loaded into a
//       particleSamples object in memory (particle_list / particle_list_mixed_event,
//       src/particleSamples.h:82,90) and every batch is pushed through
//       HBT_correlation::calculate_HBT_correlation_function (src/HBT_correlation.cpp:177).
//   ref_driver files <params.dat> <path> <out.bin> <particles_out.bin>
//       the reference's own reader loop (src/Analysis.cpp:817-835) on the files under
//       <path>; additionally writes the filtered particle lists it saw as HBTIN001 so
//       that the same input can be replayed on a box that has no reference.
//   ref_driver rng   <seed> <n>
//       prints n rand_int_uniform() then n rand_uniform() draws of RandomUtil::Random
//       (src/Random.h:21-22) — pins the host-side RNG replay.
//
// cwd must hold EOS/pdg.dat (src/particle_decay.cpp:24-36); `mem` mode creates a
// scratch directory with a link to oracle/_ref/EOS/pdg.dat and an empty mode-10 file.

#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

// every standard header the reference headers pull in comes first, so that the
// access override below only touches the reference's own classes
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <utility>

#define private public
#include "HBT_correlation.h"
#undef private

namespace {

typedef std::vector<std::vector<particle_info> *> evlist_t;

struct Batch {
    std::vector<std::vector<double>> same, mixed;  // each event: n*8 doubles
};

void die(const char *msg) {
    fprintf(stderr, "ref_driver: %s\n", msg);
    exit(2);
}

void rd(FILE *f, void *p, size_t n) {
    if (fread(p, 1, n, f) != n) die("short read");
}

std::vector<Batch> read_batches(const char *fn) {
    FILE *f = fopen(fn, "rb");
    if (!f) die("cannot open batches file");
    char magic[8];
    rd(f, magic, 8);
    if (memcmp(magic, "HBTIN001", 8) != 0) die("bad magic");
    int32_t nb;
    rd(f, &nb, 4);
    std::vector<Batch> out(nb);
    for (auto &b : out) {
        int32_t nev, nmx;
        rd(f, &nev, 4);
        rd(f, &nmx, 4);
        b.same.resize(nev);
        b.mixed.resize(nmx);
        for (auto *lst : {&b.same, &b.mixed}) {
            for (auto &ev : *lst) {
                int32_t n;
                rd(f, &n, 4);
                ev.resize(static_cast<size_t>(n) * 8);
                if (n) rd(f, ev.data(), ev.size() * sizeof(double));
            }
        }
    }
    fclose(f);
    return out;
}

void fill(evlist_t *dst, const std::vector<std::vector<double>> &src, double mass, int monval) {
    for (auto *v : *dst) delete v;
    dst->clear();
    for (const auto &ev : src) {
        auto *v = new std::vector<particle_info>;
        size_t n = ev.size() / 8;
        v->resize(n);
        for (size_t i = 0; i < n; i++) {
            particle_info &p = (*v)[i];
            memset(&p, 0, sizeof(p));
            p.monval = monval;
            p.mass = mass;
            p.px = ev[8 * i + 0];
            p.py = ev[8 * i + 1];
            p.pz = ev[8 * i + 2];
            p.E = ev[8 * i + 3];
            p.x = ev[8 * i + 4];
            p.y = ev[8 * i + 5];
            p.z = ev[8 * i + 6];
            p.t = ev[8 * i + 7];
        }
        dst->push_back(v);
    }
}

void wr(FILE *f, const void *p, size_t n) {
    if (fwrite(p, 1, n, f) != n) die("short write");
}

// raw accumulator dump, format HBTOUT01 (tests/hbtio.py reads it)
void dump(HBT_correlation &h, const char *fn, const std::vector<double> &psi,
          double t_same, double t_total, uint64_t pairs_same, uint64_t pairs_mixed) {
    FILE *f = fopen(fn, "wb");
    if (!f) die("cannot open output file");
    wr(f, "HBTOUT01", 8);
    int32_t hdr[6] = {h.azimuthal_flag_, h.invariant_radius_flag_, h.n_KT, h.n_Kphi, h.qnpts,
                      static_cast<int32_t>(psi.size())};
    wr(f, hdr, sizeof(hdr));
    wr(f, psi.data(), psi.size() * 8);
    double tt[2] = {t_same, t_total};
    wr(f, tt, 16);
    uint64_t pp[2] = {pairs_same, pairs_mixed};
    wr(f, pp, 16);
    const int nK = h.n_KT, nP = h.n_Kphi, nq = h.qnpts;
    std::vector<uint64_t> c;
    for (int k = 0; k < nK; k++) c.push_back(h.number_of_pairs_numerator_KTdiff[k]);
    for (int k = 0; k < nK; k++) c.push_back(h.number_of_pairs_denormenator_KTdiff[k]);
    for (int k = 0; k < nK; k++) c.push_back(h.number_of_pairs_numerator_KTdiff_qinv_[k]);
    for (int k = 0; k < nK; k++) c.push_back(h.number_of_pairs_denormenator_KTdiff_qinv_[k]);
    wr(f, c.data(), c.size() * 8);
    if (h.azimuthal_flag_ == 1) {
        c.clear();
        for (int k = 0; k < nK; k++)
            for (int p = 0; p < nP; p++) c.push_back(h.number_of_pairs_numerator_KTKphidiff[k][p]);
        for (int k = 0; k < nK; k++)
            for (int p = 0; p < nP; p++)
                c.push_back(h.number_of_pairs_denormenator_KTKphidiff[k][p]);
        wr(f, c.data(), c.size() * 8);
    }
    // six 3-D histograms, flat index ((K[*nP+phi])*nq+o)*nq+s)*nq+l
    if (h.azimuthal_flag_ == 0) {
        double ****arr[6] = {h.correl_3d_num_count, h.correl_3d_num, h.q_out_mean,
                             h.q_side_mean,         h.q_long_mean,   h.correl_3d_denorm};
        for (auto a : arr)
            for (int k = 0; k < nK; k++)
                for (int o = 0; o < nq; o++)
                    for (int s = 0; s < nq; s++) wr(f, a[k][o][s], nq * 8);
    } else {
        double *****arr[6] = {h.correl_3d_Kphi_diff_num_count, h.correl_3d_Kphi_diff_num,
                              h.q_out_diff_mean,               h.q_side_diff_mean,
                              h.q_long_diff_mean,              h.correl_3d_Kphi_diff_denorm};
        for (auto a : arr)
            for (int k = 0; k < nK; k++)
                for (int p = 0; p < nP; p++)
                    for (int o = 0; o < nq; o++)
                        for (int s = 0; s < nq; s++) wr(f, a[k][p][o][s], nq * 8);
    }
    if (h.invariant_radius_flag_ == 1) {
        double **arr[4] = {h.correl_1d_inv_num_count, h.q_inv_mean, h.correl_1d_inv_num,
                           h.correl_1d_inv_denorm};
        for (auto a : arr)
            for (int k = 0; k < nK; k++) wr(f, a[k], nq * 8);
    }
    fclose(f);
}

std::string exe_dir() {
    char buf[4096];
    ssize_t n = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
    if (n <= 0) die("readlink");
    buf[n] = 0;
    std::string s(buf);
    return s.substr(0, s.rfind('/'));
}

uint64_t count_same_pairs(HBT_correlation &h, particleSamples &ps) {
    // pairs as the reference logs them (src/HBT_correlation.cpp:282-284): after the
    // single-particle rapidity cut of :264-266
    const double lo = tanh(h.Krap_min_), hi = tanh(h.Krap_max_);
    uint64_t n = 0;
    for (auto *ev : *ps.particle_list)
        for (auto &p : *ev) {
            double r = p.pz / p.E;
            if (r > lo && r < hi) n++;
        }
    return n * (n - 1) / 2;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) die("usage: ref_driver mem|files|rng ...");
    std::string mode = argv[1];

    if (mode == "rng") {
        if (argc < 4) die("usage: ref_driver rng seed n");
        RandomUtil::Random r(atoi(argv[2]));
        int n = atoi(argv[3]);
        for (int i = 0; i < n; i++) printf("%d\n", r.rand_int_uniform());
        for (int i = 0; i < n; i++) printf("%.17g\n", r.rand_uniform());
        return 0;
    }

    if (argc < 5) die("usage: ref_driver mem|files params in out ...");
    char abs_params[4096], abs_in[4096], abs_out[4096], abs_aux[4096] = "";
    if (!realpath(argv[2], abs_params)) die("params path");
    if (!realpath(argv[3], abs_in)) die("input path");
    {
        // output may not exist yet: resolve its directory
        std::string o(argv[4]);
        FILE *t = fopen(o.c_str(), "wb");
        if (!t) die("cannot create output");
        fclose(t);
        if (!realpath(o.c_str(), abs_out)) die("output path");
    }
    bool same_only = false;
    std::vector<int> only;
    bool use_only = false;
    for (int ia = 5; mode == "mem" && ia < argc; ia++) {
        const std::string arg(argv[ia]);
        if (arg == "same_only") same_only = true;
        if (arg.rfind("only=", 0) == 0) {
            use_only = true;
            const std::string lst = arg.substr(5);
            size_t pos = 0;
            while (pos < lst.size()) {
                size_t c = lst.find(',', pos);
                if (c == std::string::npos) c = lst.size();
                if (c > pos) only.push_back(atoi(lst.substr(pos, c - pos).c_str()));
                pos = c + 1;
            }
        }
    }
    if (mode == "files") {
        if (argc < 6) die("files mode needs particles_out");
        FILE *t = fopen(argv[5], "wb");
        if (!t) die("cannot create particles_out");
        fclose(t);
        if (!realpath(argv[5], abs_aux)) die("particles_out path");
    }

    ParameterReader paraRdr;
    paraRdr.readFromFile(abs_params);
    int seed = paraRdr.getVal("randomSeed");
    std::shared_ptr<RandomUtil::Random> ran(new RandomUtil::Random(seed));

    std::string path;
    char scratch[] = "/tmp/hbt_ref_XXXXXX";
    if (mode == "mem") {
        if (!mkdtemp(scratch)) die("mkdtemp");
        if (chdir(scratch) != 0) die("chdir");
        mkdir("EOS", 0755);
        std::string pdg = exe_dir() + "/EOS/pdg.dat";
        if (symlink(pdg.c_str(), "EOS/pdg.dat") != 0) die("symlink pdg");
        mkdir("results", 0755);
        gzFile g = gzopen("results/particle_samples.gz", "wb");
        gzclose(g);
        paraRdr.setVal("read_in_mode", 10);
        paraRdr.setVal("read_in_real_mixed_events", 0);
        path = "results";
    } else {
        // cwd = directory that holds EOS/ and <path>; <path> given relative to it or absolute
        path = abs_in;
    }

    auto plist = std::make_shared<particleSamples>(paraRdr, path, ran);
    HBT_correlation hbt(paraRdr, path, ran);
    std::vector<double> psi;
    double t_same = 0., t_total = 0.;
    uint64_t pairs_same = 0, pairs_mixed = 0;
    typedef std::chrono::steady_clock clk;

    if (mode == "mem") {
        std::vector<Batch> batches = read_batches(abs_in);
        double mass = paraRdr.getVal("particle_mass", 0.13957);
        int monval = paraRdr.getVal("particle_monval");
        evlist_t *own_mixed = plist->particle_list_mixed_event;
        int ibatch = -1;
        for (auto &b : batches) {
            ibatch++;
            if (use_only && std::find(only.begin(), only.end(), ibatch) == only.end()) {
                // replay this batch's draws (src/HBT_correlation.cpp:200-215 and :495), no pair work
                const int nev = b.same.size();
                const int mixed_nev = b.mixed.empty() ? nev : static_cast<int>(b.mixed.size());
                const int nmix = mixed_nev / 2 + 1;
                for (int iev = 0; iev < nev && mixed_nev > 0 && !same_only; iev++) {
                    for (int c = 0; c < nmix; c++) {
                        int id = ran->rand_int_uniform() % mixed_nev;
                        while (iev == id && mixed_nev != 1) id = ran->rand_int_uniform() % mixed_nev;
                    }
                    for (int c = 0; c < nmix; c++) ran->rand_uniform();
                }
                psi.push_back(0.0);
                continue;
            }
            fill(plist->particle_list, b.same, mass, monval);
            if (b.mixed.empty()) {
                // same aliasing as src/particleSamples.cpp:528-530
                plist->particle_list_mixed_event = plist->particle_list;
            } else {
                fill(own_mixed, b.mixed, mass, monval);
                plist->particle_list_mixed_event = own_mixed;
            }
            pairs_same += count_same_pairs(hbt, *plist);
            auto t0 = clk::now();
            if (same_only) {
                hbt.set_particle_list(plist);
                int nev = plist->get_number_of_events();
                if (hbt.azimuthal_flag_ == 1) hbt.calculate_flow_event_plane_angle(2);
                hbt.number_of_oversample_events_ = nev;
                std::vector<int> lst(nev);
                for (int i = 0; i < nev; i++) lst[i] = i;
                hbt.combine_and_bin_particle_pairs(lst);
                t_same += std::chrono::duration<double>(clk::now() - t0).count();
            } else {
                hbt.calculate_HBT_correlation_function(plist);
            }
            t_total += std::chrono::duration<double>(clk::now() - t0).count();
            psi.push_back(hbt.get_psi_ref());
        }
        plist->particle_list_mixed_event = own_mixed;  // let the dtor free what it owns
        // mixed pair count (src/HBT_correlation.cpp:555-558) is not recomputed here; the
        // python side derives it from the replayed partner ids.
        dump(hbt, abs_out, psi, t_same, t_total, pairs_same, pairs_mixed);
        if (chdir("/") != 0) die("chdir /");
        std::string rm = std::string("rm -rf ") + scratch;
        if (system(rm.c_str()) != 0) fprintf(stderr, "ref_driver: cleanup failed\n");
        return 0;
    }

    // files mode: the loop of src/Analysis.cpp:817-835, plus a record of what was read
    FILE *pf = fopen(abs_aux, "wb");
    std::vector<std::vector<std::vector<double>>> rec_same, rec_mixed;
    bool real_mixed = paraRdr.getVal("read_in_real_mixed_events") == 1;
    while (!plist->end_of_file()) {
        plist->read_in_particle_samples_and_filter();
        plist->read_in_particle_samples_mixed_event_and_filter();
        auto grab = [](evlist_t *l) {
            std::vector<std::vector<double>> out;
            for (auto *ev : *l) {
                std::vector<double> v;
                for (auto &p : *ev)
                    for (double d : {p.px, p.py, p.pz, p.E, p.x, p.y, p.z, p.t}) v.push_back(d);
                out.push_back(v);
            }
            return out;
        };
        rec_same.push_back(grab(plist->particle_list));
        rec_mixed.push_back(real_mixed ? grab(plist->particle_list_mixed_event)
                                       : std::vector<std::vector<double>>());
        pairs_same += count_same_pairs(hbt, *plist);
        auto t0 = clk::now();
        hbt.calculate_HBT_correlation_function(plist);
        t_total += std::chrono::duration<double>(clk::now() - t0).count();
        psi.push_back(hbt.get_psi_ref());
    }
    hbt.output_HBTcorrelation();
    wr(pf, "HBTIN001", 8);
    int32_t nb = rec_same.size();
    wr(pf, &nb, 4);
    for (int b = 0; b < nb; b++) {
        int32_t nev = rec_same[b].size(), nmx = rec_mixed[b].size();
        wr(pf, &nev, 4);
        wr(pf, &nmx, 4);
        for (auto *lst : {&rec_same[b], &rec_mixed[b]})
            for (auto &ev : *lst) {
                int32_t n = ev.size() / 8;
                wr(pf, &n, 4);
                if (n) wr(pf, ev.data(), ev.size() * 8);
            }
    }
    fclose(pf);
    dump(hbt, abs_out, psi, t_same, t_total, pairs_same, pairs_mixed);
    return 0;
}
