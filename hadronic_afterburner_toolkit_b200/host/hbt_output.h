// Output stage of the HBT path, free of any reference type: host copies of the accumulators and
// the writers of results/HBT_correlation_function_*.dat in the reference's frozen format
// (/root/reference/src/HBT_correlation.cpp:694-855: file names, row order q_long outer / q_out
// middle / q_side inner, column layout, setw(18) setprecision(8) scientific, the count < 2 rule,
// ecoOutput).  Used by the drop-in class (HBT_correlation.cpp) and by the fast driver
// (hbt_fast_analysis.cpp).
#ifndef HBT_B200_HOST_OUTPUT_H_
#define HBT_B200_HOST_OUTPUT_H_

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/hbt_b200.h"

struct HbtHostResults {
    std::vector<uint64_t> num_count, den_count, npairs_num, npairs_den;
    std::vector<double> num_cos, sum_qo, sum_qs, sum_ql;
    std::vector<uint64_t> inv_count, inv_den, npairs_num_inv, npairs_den_inv;
    std::vector<double> inv_sum, inv_cos;
};

// hbt_read (+ hbt_read_qinv) into host vectors; returns the library's status
inline int hbt_fetch_results(hbt_ctx *c, const hbt_params &p, HbtHostResults &r) {
    const size_t nb = static_cast<size_t>(hbt_num_bins(c)), ns = static_cast<size_t>(hbt_num_slabs(c));
    r.num_count.resize(nb); r.den_count.resize(nb); r.num_cos.resize(nb);
    r.sum_qo.resize(nb); r.sum_qs.resize(nb); r.sum_ql.resize(nb);
    r.npairs_num.resize(ns); r.npairs_den.resize(ns);
    int rc = hbt_read(c, r.num_count.data(), r.num_cos.data(), r.sum_qo.data(), r.sum_qs.data(), r.sum_ql.data(),
                      r.den_count.data(), r.npairs_num.data(), r.npairs_den.data());
    if (rc != HBT_OK) return rc;
    if (p.invariant_radius_flag == 1) {
        const size_t n1 = static_cast<size_t>(p.n_KT) * p.qnpts;
        r.inv_count.resize(n1); r.inv_den.resize(n1); r.inv_sum.resize(n1); r.inv_cos.resize(n1);
        r.npairs_num_inv.resize(p.n_KT); r.npairs_den_inv.resize(p.n_KT);
        rc = hbt_read_qinv(c, r.inv_count.data(), r.inv_sum.data(), r.inv_cos.data(), r.inv_den.data(),
                           r.npairs_num_inv.data(), r.npairs_den_inv.data());
    }
    return rc;
}

struct HbtOutputWriter {
    hbt_params p;
    std::string path;
    bool eco = false;
    std::vector<double> KT_array, Kphi_array, q_axis;

    HbtOutputWriter(const hbt_params &params, const std::string &out_path, bool eco_output)
        : p(params), path(out_path), eco(eco_output) {
        // the grids of src/HBT_correlation.cpp:27-33, :47-62, evaluated the same way
        const double delta_q = (p.q_max - p.q_min) / (p.qnpts - 1);
        for (int i = 0; i < p.qnpts; i++) q_axis.push_back(p.q_min + i * delta_q);
        const double dKT = (p.KT_max - p.KT_min) / (p.n_KT - 1);
        const double dKphi = 2 * M_PI / p.n_Kphi;
        for (int i = 0; i < p.n_KT; i++) KT_array.push_back(p.KT_min + i * dKT);
        for (int i = 0; i < p.n_Kphi; i++) Kphi_array.push_back(i * dKphi);
    }

    //! One row as the reference's stream prints it: `output << std::scientific << std::setw(18) << std::setprecision(8)`
    //! in front of the first value only (setw does not persist), four blanks between values, std::endl.  libstdc++
    //! formats a double in scientific notation through the C library's "%.8e", so snprintf gives the same characters;
    //! the rows of a file are collected and written once (the reference's std::endl flushes every row: one write
    //! system call per row, 69 000 per 41^3 table).
    static void append_row(std::string &buf, const double *v, int n) {
        char line[256];
        int len = std::snprintf(line, sizeof(line), "%18.8e", v[0]);
        for (int k = 1; k < n; k++) len += std::snprintf(line + len, sizeof(line) - static_cast<size_t>(len), "    %.8e", v[k]);
        line[len++] = '\n';
        buf.append(line, static_cast<size_t>(len));
    }

    //! src/HBT_correlation.cpp:694-724
    void write_inv(const HbtHostResults &r) const {
        for (int iK = 0; iK < p.n_KT - 1; iK++) {
            const double npair_ratio = (static_cast<double>(r.npairs_num_inv[iK]) / static_cast<double>(r.npairs_den_inv[iK]));
            std::ostringstream filename;
            filename << path << "/HBT_correlation_function_inv_KT_" << KT_array[iK] << "_" << KT_array[iK + 1] << ".dat";
            std::ofstream output(filename.str().c_str());
            std::string buf;
            for (int iq = 0; iq < p.qnpts; iq++) {
                const size_t k = static_cast<size_t>(iK) * p.qnpts + iq;
                const double count = static_cast<double>(r.inv_count[k]);
                const double q_inv_local = r.inv_sum[k] / count;
                const double correl_fun_num = r.inv_cos[k];
                const double correl_fun_denorm = static_cast<double>(r.inv_den[k]) * npair_ratio;
                const double row[3] = {q_inv_local, correl_fun_num, correl_fun_denorm};
                if (eco) append_row(buf, row + 1, 2); else append_row(buf, row, 3);
            }
            output.write(buf.data(), static_cast<std::streamsize>(buf.size()));
            output.close();
        }
    }

    //! one (K_T[, K_phi]) slab: rows in the order q_long outer, q_out middle, q_side inner and the
    //! column layout of src/HBT_correlation.cpp:735-777
    void write_3d_file(const HbtHostResults &r, const std::string &filename, size_t slab, double npair_ratio) const {
        std::ofstream output(filename.c_str());
        const int qnpts = p.qnpts;
        const size_t q3 = static_cast<size_t>(qnpts) * qnpts * qnpts;
        std::string buf;
        buf.reserve(q3 * (eco ? 40 : 96));
        for (int iqlong = 0; iqlong < qnpts; iqlong++) {
            for (int iqout = 0; iqout < qnpts; iqout++) {
                for (int iqside = 0; iqside < qnpts; iqside++) {
                    const size_t bin = slab * q3 + (static_cast<size_t>(iqout) * qnpts + iqside) * qnpts + iqlong;
                    // the reference keeps the counts in doubles and truncates them to int here
                    const int npart_num = static_cast<int>(static_cast<double>(r.num_count[bin]));
                    const int npart_denorm = static_cast<int>(static_cast<double>(r.den_count[bin]));
                    double q_out_local, q_side_local, q_long_local, correl_fun_num, correl_fun_denorm;
                    if (npart_num < 2 || npart_denorm < 2) {
                        q_out_local = q_axis[iqout];
                        q_side_local = q_axis[iqside];
                        q_long_local = q_axis[iqlong];
                        correl_fun_num = 0.0;
                        correl_fun_denorm = npart_denorm;
                    } else {
                        q_out_local = r.sum_qo[bin] / npart_num;
                        q_side_local = r.sum_qs[bin] / npart_num;
                        q_long_local = r.sum_ql[bin] / npart_num;
                        correl_fun_num = r.num_cos[bin];
                        correl_fun_denorm = npair_ratio * static_cast<double>(r.den_count[bin]);
                    }
                    const double row[5] = {q_out_local, q_side_local, q_long_local, correl_fun_num, correl_fun_denorm};
                    if (eco) append_row(buf, row + 3, 2); else append_row(buf, row, 5);
                }
            }
        }
        output.write(buf.data(), static_cast<std::streamsize>(buf.size()));
        output.close();
    }

    //! src/HBT_correlation.cpp:726-783
    void write_KT(const HbtHostResults &r) const {
        for (int iK = 0; iK < p.n_KT - 1; iK++) {
            const double npair_ratio = (static_cast<double>(r.npairs_num[iK]) / static_cast<double>(r.npairs_den[iK]));
            std::ostringstream filename;
            filename << path << "/HBT_correlation_function_KT_" << KT_array[iK] << "_" << KT_array[iK + 1] << ".dat";
            write_3d_file(r, filename.str(), iK, npair_ratio);
        }
    }

    //! src/HBT_correlation.cpp:785-855
    void write_KT_Kphi(const HbtHostResults &r) const {
        for (int iK = 0; iK < p.n_KT - 1; iK++) {
            for (int iKphi = 0; iKphi < p.n_Kphi; iKphi++) {
                const size_t slab = static_cast<size_t>(iK) * p.n_Kphi + iKphi;
                const double npair_ratio = (static_cast<double>(r.npairs_num[slab]) / static_cast<double>(r.npairs_den[slab]));
                std::ostringstream filename;
                filename << path << "/HBT_correlation_function_KT_" << KT_array[iK] << "_" << KT_array[iK + 1] << "_Kphi_"
                         << Kphi_array[iKphi] << ".dat";
                write_3d_file(r, filename.str(), slab, npair_ratio);
            }
        }
    }

    //! HBT_correlation::output_HBTcorrelation (the dispatch at the end of Analysis::HBTAnalysis)
    void write_all(const HbtHostResults &r) const {
        if (p.invariant_radius_flag == 1) write_inv(r);
        if (p.azimuthal_flag == 0) {
            write_KT(r);
        } else {
            write_KT_Kphi(r);
        }
    }
};

#endif  // HBT_B200_HOST_OUTPUT_H_
