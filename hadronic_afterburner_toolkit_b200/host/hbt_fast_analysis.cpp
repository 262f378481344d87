// hbt_fast_analysis.e — the HBT analysis of the reference program
// (Analysis::HBTAnalysis, /root/reference/src/Analysis.cpp:817-835) end to end on libhbt_b200.so,
// for the input formats hbt_reader_* reads (read_in_mode = 10: results/particle_samples.gz, the
// production configs' format; 2: gzipped UrQMD text results/particle_list.dat; 21: UrQMD binary
// results/particle_list.bin; 0: results/OSCAR.DAT; 1: UrQMD file-13 text results/particle_list.dat; 9: iSS
// binary results/particle_list.bin; 7: gzipped SMASH text, 4 / 3: UrQMD 3.3p / header-less UrQMD text, 5: JAM
// text, all results/particle_list.dat; 8: extended SMASH binary results/particles_binary.bin):
//
//   reader thread   hbt_reader_*      inflate + parse + species filter, two batches ahead
//   host            psi_2, rapidity cut, the reference's RNG draws (partner events, rotation angles)
//   GPU             hbt_accumulate_batch: one fused pair kernel per batch, asynchronous
//   host            hbt_read + the reference's output files (hbt_output.h)
//
// Same command line as the reference (run in a directory with parameters.dat and results/;
// name=value arguments override the file), same results/HBT_correlation_function_*.dat.
// read_in_real_mixed_events = 1 reads the partner events from the second file of the mode
// (results/particle_samples_mixed_event.gz, particle_list_mixed_event.dat / .bin), one batch per
// batch of the main file, as Analysis::HBTAnalysis does.
// It does not link any reference code.  Switches it cannot honour (other read_in_mode values,
// resonance feed-down, species groups) are refused with a message: use the drop-in binary
// (hadronic_afterburner_tools_b200.e) for those.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/hbt_b200.h"
#include "hbt_output.h"

namespace {

// "name = value  # comment" lines, then name=value arguments (ParameterReader::readFromFile / readFromArguments)
struct Params {
    std::map<std::string, double> v;
    static std::string trim(const std::string &s) {
        const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? "" : s.substr(a, b - a + 1);
    }
    void set(const std::string &line) {
        std::string l = line.substr(0, line.find('#'));
        const size_t eq = l.find('=');
        if (eq == std::string::npos) return;
        std::string name = trim(l.substr(0, eq)), val = trim(l.substr(eq + 1));
        if (name.empty() || val.empty()) return;
        char *end = nullptr;
        const double x = std::strtod(val.c_str(), &end);
        if (end == val.c_str()) return;
        v[name] = x;
    }
    double get(const std::string &name) const {
        const auto it = v.find(name);
        if (it == v.end()) {
            std::cerr << "[error] parameter " << name << " not found" << std::endl;
            std::exit(1);
        }
        return it->second;
    }
    double get(const std::string &name, double dflt) const {
        const auto it = v.find(name);
        return it == v.end() ? dflt : it->second;
    }
};

void die(const std::string &msg) {
    std::cerr << "[error] hbt_fast_analysis: " << msg << std::endl;
    std::exit(1);
}

void check(hbt_ctx *c, int rc, const char *what) {
    if (rc == HBT_OK) return;
    die(std::string(what) + " failed: " + hbt_last_error(c));
}

double seconds_since(const std::chrono::steady_clock::time_point &t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

int main(int argc, char *argv[]) {
    const auto t_start = std::chrono::steady_clock::now();
    const std::string path = "results";
    Params P;
    {
        std::ifstream f("parameters.dat");
        if (!f) die("parameters.dat not found in the working directory");
        std::string line;
        while (std::getline(f, line)) P.set(line);
    }
    for (int i = 1; i < argc; i++) P.set(argv[i]);

    if (P.get("analyze_HBT", 0) != 1) die("analyze_HBT is not 1: nothing to do (the other analyses are the reference program's)");
    const int read_in_mode = static_cast<int>(P.get("read_in_mode"));
    const bool known_mode = read_in_mode == 10 || read_in_mode == 9 || read_in_mode == 2 || read_in_mode == 21 || read_in_mode == 0 ||
                            read_in_mode == 1 || read_in_mode == 3 || read_in_mode == 4 || read_in_mode == 5 || read_in_mode == 7 ||
                            read_in_mode == 8;
    if (!known_mode) die("unknown read_in_mode (the reference has 0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 21)");
    // (modes 2 and 21 do not force these off as mode 10 does, src/particleSamples.cpp:409-412)
    if (read_in_mode != 10 && read_in_mode != 0 && read_in_mode != 9 && P.get("resonance_weak_feed_down_flag", 0) == 1)
        die("resonance_weak_feed_down_flag = 1 is not supported here");
    const bool real_mixed = P.get("read_in_real_mixed_events", 0) == 1;
    if (P.get("resonance_feed_down_flag", 0) == 1) die("resonance_feed_down_flag = 1 is not supported here");
    if (P.get("readRapidityShiftFromFile", 0) == 1) die("readRapidityShiftFromFile = 1 is not supported here");
    // particle_monval = 333: for the UrQMD / SMASH / JAM formats the reference keeps reconst_flag = 1 and appends phi
    // mesons RECONSTRUCTED from K+ K- pairs to the list (src/particleSamples.cpp:116-119, 1303-1307, 2200); hbt_reader
    // would hand out the primary phi(1020) only
    if (static_cast<int>(P.get("particle_monval")) == 333 && read_in_mode != 10 && read_in_mode != 9 && read_in_mode != 0)
        die("particle_monval = 333 with this read_in_mode needs the reference's phi-meson reconstruction: use the drop-in binary");

    hbt_params hp;
    hp.qnpts = static_cast<int>(P.get("qnpts"));
    hp.n_KT = static_cast<int>(P.get("n_KT"));
    hp.n_Kphi = static_cast<int>(P.get("n_Kphi"));
    hp.azimuthal_flag = static_cast<int>(P.get("azimuthal_flag"));
    hp.invariant_radius_flag = static_cast<int>(P.get("invariant_radius_flag"));
    hp.long_comoving_boost = P.get("long_comoving_boost") == 1 ? 1 : 0;
    hp.q_min = P.get("q_min");
    hp.q_max = P.get("q_max");
    hp.KT_min = P.get("KT_min");
    hp.KT_max = P.get("KT_max");
    hp.HBTrap_min = P.get("HBTrap_min");
    hp.HBTrap_max = P.get("HBTrap_max");
    hp.needed_number_of_pairs = P.get("needed_number_of_pairs");

    // HBT_B200_DEVICES as in the drop-in class: "all", a count, or a list of devices; one context per entry, the
    // batches go to them in turn and the ordered pair cap stays exact across them (hbt_group_*)
    std::vector<int32_t> devs;
    if (const char *e = std::getenv("HBT_B200_DEVICES")) {
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
    }
    if (devs.empty()) devs.push_back(0);
    hbt_group *grp = nullptr;
    if (hbt_group_create(&hp, static_cast<int32_t>(devs.size()), devs.data(), &grp) != HBT_OK)
        die(std::string("cannot create the GPU engine: ") + hbt_group_last_error(nullptr) + hbt_last_error(nullptr));
    hbt_ctx *ctx = hbt_group_ctx(grp, 0);
    hbt_rng *rng = nullptr;
    if (hbt_rng_create(static_cast<int32_t>(P.get("randomSeed")), &rng) != HBT_OK) die("cannot create the random number generator");
    hbt_reader *rd = nullptr;
    // file names of src/particleSamples.cpp:133-149
    const bool bin_file = read_in_mode == 21 || read_in_mode == 9;
    const std::string file = path + (read_in_mode == 10 ? "/particle_samples.gz" : bin_file ? "/particle_list.bin"
                                     : read_in_mode == 8 ? "/particles_binary.bin"
                                     : read_in_mode == 0 ? "/OSCAR.DAT" : "/particle_list.dat");
    if (hbt_reader_open(file.c_str(), read_in_mode, static_cast<int>(P.get("particle_monval")), static_cast<int64_t>(P.get("event_buffer_size")),
                        P.get("rapidity_shift", 0), nullptr, &rd) != HBT_OK)
        die("cannot open " + file + " for particle_monval " + std::to_string(static_cast<int>(P.get("particle_monval"))));
    hbt_reader *rd2 = nullptr;  // the mixed-event file (src/particleSamples.cpp:133-149, :527-535)
    if (real_mixed) {
        const std::string file2 = path + (read_in_mode == 10 ? "/particle_samples_mixed_event.gz"
                                          : bin_file ? "/particle_list_mixed_event.bin"
                                          : read_in_mode == 8 ? "/particles_binary_mixed_event.bin"
                                          : read_in_mode == 0 ? "/OSCAR_mixed_event.DAT" : "/particle_list_mixed_event.dat");
        if (hbt_reader_open(file2.c_str(), read_in_mode, static_cast<int>(P.get("particle_monval")),
                            static_cast<int64_t>(P.get("event_buffer_size")), P.get("rapidity_shift", 0), nullptr, &rd2) != HBT_OK)
            die("cannot open " + file2);
    }

    std::vector<double> cut, cut2;
    std::vector<int64_t> off, off2;
    std::vector<int32_t> ids;
    std::vector<double> cs;
    long long n_batches = 0, n_events = 0;
    unsigned long long pairs_same = 0, pairs_mixed = 0;
    double t_wait = 0.0;
    for (;;) {
        const double *p = nullptr;
        const int64_t *o = nullptr;
        const auto tw = std::chrono::steady_clock::now();
        const int nev = hbt_reader_next(rd, &p, &o, nullptr);
        t_wait += seconds_since(tw);
        if (nev < 0) die(std::string("reader: ") + hbt_reader_error(rd));
        if (nev == 0) break;
        // Psi_2 over every particle of the species in the batch (src/HBT_correlation.cpp:233-249)
        const double psi_ref = hp.azimuthal_flag == 1 ? hbt_psi_ref(p, o[nev], 2) : 0.0;
        // the gather of each event with its rapidity cut (:255-281)
        cut.resize(static_cast<size_t>(o[nev]) * 8 + 8);
        off.assign(1, 0);
        for (int e = 0; e < nev; e++) {
            const int64_t kept = hbt_gather_rapidity(&hp, p + 8 * o[e], o[e + 1] - o[e], cut.data() + 8 * off.back());
            off.push_back(off.back() + kept);
        }
        // the partner events: the batch itself, or the next batch of the mixed-event file (same gather)
        int nev2 = nev;
        const int64_t *o2 = off.data();
        if (rd2) {
            const double *p2 = nullptr;
            const int64_t *r2 = nullptr;
            nev2 = hbt_reader_next(rd2, &p2, &r2, nullptr);
            if (nev2 < 0) die(std::string("reader (mixed events): ") + hbt_reader_error(rd2));
            off2.assign(1, 0);
            if (nev2 > 0) {
                cut2.resize(static_cast<size_t>(r2[nev2]) * 8 + 8);
                for (int e = 0; e < nev2; e++) {
                    const int64_t kept = hbt_gather_rapidity(&hp, p2 + 8 * r2[e], r2[e + 1] - r2[e], cut2.data() + 8 * off2.back());
                    off2.push_back(off2.back() + kept);
                }
            }
            o2 = off2.data();
        }
        // the draws of the batch in the reference's order (:202-217, :493-497); a batch without partner events
        // has no mixed-event loop (the reference would take "% 0" there)
        const int nmix = nev2 > 0 ? nev2 / 2 + 1 : 0;
        ids.resize(static_cast<size_t>(nev) * nmix);
        cs.resize(static_cast<size_t>(nev) * nmix * 2);
        if (nmix) hbt_rng_mixed_plan(rng, nev, nev2, ids.data(), cs.data(), nullptr);
        if (hbt_group_accumulate_batch(grp, cut.data(), off.data(), nev, rd2 ? cut2.data() : nullptr, rd2 ? off2.data() : nullptr,
                                       rd2 ? nev2 : 0, ids.data(), cs.data(), nmix, psi_ref, 1, nmix > 0 ? 1 : 0) != HBT_OK)
            die(std::string("hbt_group_accumulate_batch failed: ") + hbt_group_last_error(grp));
        const unsigned long long n1 = static_cast<unsigned long long>(off.back());
        pairs_same += n1 > 1 ? n1 * (n1 - 1) / 2 : 0;
        for (int e = 0; e < nev; e++)
            for (int c = 0; c < nmix; c++) {
                const int id = ids[static_cast<size_t>(e) * nmix + c];
                pairs_mixed += static_cast<unsigned long long>(off[e + 1] - off[e]) * (o2[id + 1] - o2[id]);
            }
        n_batches++;
        n_events += nev;
    }
    if (hbt_group_reduce(grp) != HBT_OK) die(std::string("hbt_group_reduce failed: ") + hbt_group_last_error(grp));
    const double t_loop = seconds_since(t_start);
    HbtHostResults res;
    check(ctx, hbt_fetch_results(ctx, hp, res), "hbt_read");
    HbtOutputWriter(hp, path, P.get("ecoOutput", 0) == 1).write_all(res);
    double same_ms = 0, mixed_ms = 0;
    for (int d = 0; d < hbt_group_size(grp); d++) {  // (sum over the GPUs: they work at the same time)
        double a = 0, b = 0;
        hbt_get_timers(hbt_group_ctx(grp, d), &a, &b, nullptr, nullptr);
        same_ms += a;
        mixed_ms += b;
    }
    std::printf("[info] hbt_fast_analysis: %lld batches, %lld events, %llu same + %llu mixed pairs; %.3f s in all "
                "(%.3f s waiting for the reader, %.1f MB of text, pair kernels %.3f s), output %.3f s\n",
                n_batches, n_events, pairs_same, pairs_mixed, t_loop, t_wait, hbt_reader_bytes(rd) / 1e6,
                (same_ms + mixed_ms) * 1e-3, seconds_since(t_start) - t_loop);
    hbt_reader_close(rd);
    if (rd2) hbt_reader_close(rd2);
    hbt_rng_destroy(rng);
    hbt_group_destroy(grp);
    return 0;
}
