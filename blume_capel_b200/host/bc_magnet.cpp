// bc_magnet -- the reference's susceptibility table (corner_contribution/magnet.cpp) straight from the
// engine's fused magnetisation, without the round trip through one cluster file per sample.
//
//   bc_magnet out.csv T D J [--L 48,64] [--runs 100] [--samples 200] [--burn-sweeps n]
//             [--sweeps-per-sample n] [--seed s] [--device d] [--boundary periodic|open]
//       every run is one replica of a batched GPU launch; M is measured in the sweep kernel.
//   bc_magnet out.csv --from-m samples.txt
//       CPU only: reads lines "L run m" (what magnet.cpp derives from the files of one sample) and writes
//       the table; used to check byte-compatibility with the reference tool.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "../../include/bc_engine.h"
#include "fss_stats.hpp"

static std::vector<int> parse_list(const std::string& s) {
    std::vector<int> v;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) v.push_back(std::atoi(tok.c_str()));
    return v;
}

int main(int argc, const char* argv[]) {
    if (argc < 4) {
        std::cerr << "usage: bc_magnet out.csv T D J [flags] | bc_magnet out.csv --from-m samples.txt\n";
        return 2;
    }
    const std::string out_path = argv[1];
    std::vector<int> Ls = {48, 64};  // magnet.cpp:102
    std::vector<std::vector<std::vector<double>>> M;
    if (std::string(argv[2]) == "--from-m") {
        std::ifstream in(argv[3]);
        if (!in) { std::cerr << "bc_magnet: cannot read " << argv[3] << std::endl; return 1; }
        std::map<int, std::map<int, std::vector<double>>> data;
        int L, run;
        double m;
        while (in >> L >> run >> m) data[L][run].push_back(m);
        Ls.clear();
        for (auto& kv : data) {
            Ls.push_back(kv.first);
            M.emplace_back();
            for (auto& rv : kv.second) M.back().push_back(rv.second);
        }
    } else {
        if (argc < 5) return 2;
        const double T = std::atof(argv[2]), D = std::atof(argv[3]), J = std::atof(argv[4]);
        int runs = 100, samples = 200, device = 0, boundary = BC_PERIODIC;
        int64_t burn = -1, sps = -1;
        uint64_t seed = 1;
        for (int i = 5; i < argc; ++i) {
            const std::string a = argv[i];
            auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
            if (a == "--L") Ls = parse_list(next());
            else if (a == "--runs") runs = std::atoi(next());
            else if (a == "--samples") samples = std::atoi(next());
            else if (a == "--burn-sweeps") burn = std::atoll(next());
            else if (a == "--sweeps-per-sample") sps = std::atoll(next());
            else if (a == "--seed") seed = std::strtoull(next(), nullptr, 0);
            else if (a == "--device") device = std::atoi(next());
            else if (a == "--boundary") boundary = std::string(next()) == "open" ? BC_OPEN : BC_PERIODIC;
            else { std::cerr << "bc_magnet: unknown flag " << a << std::endl; return 2; }
        }
        for (int L : Ls) {
            const int64_t N = (int64_t)L * L;
            const int64_t b = burn >= 0 ? burn : 50 * N, s = sps > 0 ? sps : N;  // ~ L^2 sweeps decorrelate (z ~ 2.2)
            bc_engine* e = nullptr;
            if (bc_create(&e, L, boundary, runs, BC_STORAGE_INT8, device, seed + (uint64_t)L) != BC_OK) {
                std::cerr << "bc_magnet: " << bc_last_error(nullptr) << std::endl;
                return 1;
            }
            std::vector<bc_obs> obs((size_t)samples * runs);
            if (bc_set_params(e, -1, T, D, J) != BC_OK || bc_init_random(e, -1) != BC_OK || bc_sweep(e, b, 0, nullptr) < 0 ||
                bc_set_sweep_counter(e, (b + s - 1) / s * s) != BC_OK || bc_sweep(e, (int64_t)samples * s, s, obs.data()) < 0) {
                std::cerr << "bc_magnet: " << bc_last_error(e) << std::endl;
                bc_destroy(e);
                return 1;
            }
            M.emplace_back((size_t)runs);
            for (int k = 0; k < samples; ++k)
                for (int r = 0; r < runs; ++r) M.back()[(size_t)r].push_back((double)obs[(size_t)k * runs + r].M);
            bc_destroy(e);
        }
    }
    std::ofstream f(out_path);
    if (!f) { std::cerr << "bc_magnet: cannot write " << out_path << std::endl; return 1; }
    f << bcfss::magnet_csv(Ls, M);
    return 0;
}
