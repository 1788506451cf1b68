// bc_gamma_nu -- the reference's cluster-size susceptibility table (corner_contribution/gamma_nu.cpp) from the
// engine's on-GPU cluster labelling instead of one cluster file per sample.
//   bc_gamma_nu out.txt T D J [--kind spin|fk] [--L 12,16,24,32] [--runs 24] [--samples 100] [--burn-sweeps n]
//               [--sweeps-per-sample n] [--cluster-every k] [--seed s] [--device d]
//   bc_gamma_nu out.txt --from-s samples.txt      (CPU only: lines "L run sum_size_sq max_size")
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "../../include/bc_engine.h"
#include "fss_stats.hpp"

int main(int argc, const char* argv[]) {
    if (argc < 4) {
        std::cerr << "usage: bc_gamma_nu out.txt T D J [flags] | bc_gamma_nu out.txt --from-s samples.txt\n";
        return 2;
    }
    std::vector<int> Ls = {12, 16, 24, 32};  // gamma_nu.cpp:101
    std::vector<std::vector<std::vector<bcfss::ClusterSample>>> S;
    if (std::string(argv[2]) == "--from-s") {
        std::ifstream in(argv[3]);
        if (!in) { std::cerr << "bc_gamma_nu: cannot read " << argv[3] << std::endl; return 1; }
        std::map<int, std::map<int, std::vector<bcfss::ClusterSample>>> data;
        int L, run;
        double a, b;
        while (in >> L >> run >> a >> b) data[L][run].push_back({a, b});
        Ls.clear();
        for (auto& kv : data) {
            Ls.push_back(kv.first);
            S.emplace_back();
            for (auto& rv : kv.second) S.back().push_back(rv.second);
        }
    } else {
        if (argc < 5) return 2;
        const double T = std::atof(argv[2]), D = std::atof(argv[3]), J = std::atof(argv[4]);
        int runs = 24, samples = 100, device = 0, kind = 0;
        int64_t burn = -1, sps = -1, cluster_every = 3;
        uint64_t seed = 1;
        for (int i = 5; i < argc; ++i) {
            const std::string a = argv[i];
            auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
            if (a == "--L") { Ls.clear(); std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) Ls.push_back(std::atoi(t.c_str())); }
            else if (a == "--kind") kind = std::string(next()) == "fk" ? 1 : 0;
            else if (a == "--runs") runs = std::atoi(next());
            else if (a == "--samples") samples = std::atoi(next());
            else if (a == "--burn-sweeps") burn = std::atoll(next());
            else if (a == "--sweeps-per-sample") sps = std::atoll(next());
            else if (a == "--cluster-every") cluster_every = std::atoll(next());
            else if (a == "--seed") seed = std::strtoull(next(), nullptr, 0);
            else if (a == "--device") device = std::atoi(next());
            else { std::cerr << "bc_gamma_nu: unknown flag " << a << std::endl; return 2; }
        }
        for (int L : Ls) {
            const int64_t N = (int64_t)L * L, b = burn >= 0 ? burn : 20 * N, s = sps > 0 ? sps : 30;
            bc_engine* e = nullptr;
            if (bc_create(&e, L, BC_PERIODIC, runs, BC_STORAGE_INT8, device, seed + (uint64_t)L) != BC_OK) {
                std::cerr << "bc_gamma_nu: " << bc_last_error(nullptr) << std::endl;
                return 1;
            }
            auto evolve = [&](int64_t n) {
                while (n > 0) {
                    const int64_t k = (cluster_every > 0 && cluster_every < n) ? cluster_every : n;
                    if (bc_sweep(e, k, 0, nullptr) < 0) return false;
                    if (cluster_every > 0 && k == cluster_every && bc_cluster_update(e, 1) != BC_OK) return false;
                    n -= k;
                }
                return true;
            };
            std::vector<bc_cluster_stats> st((size_t)runs);
            bool ok = bc_set_params(e, -1, T, D, J) == BC_OK && bc_init_random(e, -1) == BC_OK && evolve(b);
            S.emplace_back((size_t)runs);
            for (int k = 0; ok && k < samples; ++k) {
                ok = evolve(s) && bc_cluster_moments(e, kind, k, st.data()) == BC_OK;
                for (int r = 0; ok && r < runs; ++r) S.back()[(size_t)r].push_back({(double)st[(size_t)r].sum_size_sq, (double)st[(size_t)r].max_size});
            }
            if (!ok) { std::cerr << "bc_gamma_nu: " << bc_last_error(e) << std::endl; bc_destroy(e); return 1; }
            bc_destroy(e);
        }
    }
    std::ofstream f(argv[1]);
    if (!f) { std::cerr << "bc_gamma_nu: cannot write " << argv[1] << std::endl; return 1; }
    f << bcfss::gamma_nu_table(Ls, S);
    return 0;
}
