// Drop-in replacement of the reference's class BalanceFunction
// (/root/reference/src/BalanceFunction.h:16-75): same include guard, same constructor, same public
// methods, same parameters.dat switches, same three output files — the pair loops
// (src/BalanceFunction.cpp:120-197) run on a B200 through the C ABI hbt_bf_* of libhbt_b200.so.
// Compiled against the reference's own headers, like HBT_correlation.h.
#ifndef BALANCEFUNCTION_H_
#define BALANCEFUNCTION_H_

#include <memory>
#include <string>
#include <vector>

#include "ParameterReader.h"
#include "Random.h"
#include "particleSamples.h"
#include "pretty_ostream.h"

struct hbt_bf;

class BalanceFunction {
  private:
    const ParameterReader paraRdr_;
    const std::string path_;
    std::shared_ptr<particleSamples> particle_list;
    std::weak_ptr<RandomUtil::Random> ran_gen_ptr;
    pretty_ostream messager;

    int particle_monval_a;
    int particle_monval_b;
    bool same_species;
    long int N_b, N_bbar;
    int rap_type_;
    int Bnpts;
    int Bnphi;
    double dphi;
    double Bphi_min;
    double Brap_min;
    double Brap_max;
    double drap;
    double BpT_min, BpT_max;

    hbt_bf *bf_;

    typedef std::vector<std::vector<particle_info> *> plist_t;
    // (phi_p, rapidity) of the particles inside the pT cut, flat, with per-event offsets
    struct Flat {
        std::vector<double> v;
        std::vector<long long> off;
    };
    void gather(const plist_t *plist, Flat &out);
    void run(int hist, const Flat &a, const Flat &b, const std::vector<int> &partner, const std::vector<double> &rotation);

  public:
    BalanceFunction(const ParameterReader &paraRdr, const std::string path, std::shared_ptr<RandomUtil::Random> ran_gen);
    ~BalanceFunction();
    BalanceFunction(const BalanceFunction &) = delete;
    BalanceFunction &operator=(const BalanceFunction &) = delete;

    void set_particle_list(std::shared_ptr<particleSamples> particle_list_in) { particle_list = particle_list_in; }
    bool check_same_particle(const particle_info &lhs, const particle_info &rhs);
    void calculate_balance_function(std::shared_ptr<particleSamples> particle_list_in);
    int get_number_of_particles(const std::vector<std::vector<particle_info> *> *plist_b);
    void output_balance_function();
};

#endif  // BALANCEFUNCTION_H_
