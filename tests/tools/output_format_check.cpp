// row formatting of host/hbt_output.h against the reference's stream expression, character for character
#include <cstring>
#include <iostream>
#include <limits>
#include <random>
#include "hbt_output.h"
int main() {
    std::mt19937_64 g(7);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    std::vector<double> vals = {0.0, -0.0, 1.0, -1.0, 1e-300, -1e300, 9.999999995e-5, 9.9999999949e7, 123456789.0, 0.5e-8,
                                std::numeric_limits<double>::infinity(), -std::numeric_limits<double>::infinity(),
                                std::numeric_limits<double>::quiet_NaN(), 5e-324, 1.7976931348623157e308, 2.5, 0.125, 1e22, 1e23};
    for (int k = 0; k < 200000; k++) vals.push_back(u(g) * std::pow(10.0, static_cast<int>(u(g) * 30)));
    size_t bad = 0;
    for (size_t i = 0; i + 5 <= vals.size(); i += 5) {
        for (int n : {2, 3, 5}) {
            std::ostringstream o;
            o << std::scientific << std::setw(18) << std::setprecision(8);
            o << vals[i];
            for (int k = 1; k < n; k++) o << "    " << vals[i + k];
            o << std::endl;
            std::string b;
            HbtOutputWriter::append_row(b, &vals[i], n);
            if (b != o.str()) { if (bad < 5) std::cerr << "differs: [" << o.str() << "] vs [" << b << "]\n"; bad++; }
        }
    }
    std::cout << (bad ? "MISMATCH " : "identical ") << bad << " of " << 3 * (vals.size() / 5) << " rows\n";
    return bad != 0;
}
