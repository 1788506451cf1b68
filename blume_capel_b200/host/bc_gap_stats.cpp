// bc_gap_stats -- the reference's gap-size histogram per run (corner_contribution/gap_size_statistics.cpp)
//   bc_gap_stats hist out.txt L file1 [file2 ...]       CPU only: cluster files of one run -> histogram file
//   bc_gap_stats run out_root T D J [--kind spin|fk] [--L 48,64] [--runs 100] [--samples n] [--burn-sweeps n]
//                [--sweeps-per-sample n] [--cluster-every k] [--seed s] [--device d] [--boundary periodic|open]
//       GPU: every run is one replica; per sample the lattice is downloaded, its clusters are formed on the host
//       exactly as bc_run writes them, and the histogram is accumulated in memory; writes
//       {out_root}/{L}/{run}.txt like the reference tool.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <sys/stat.h>

#include "../../include/bc_engine.h"
#include "gap_format.hpp"
#include "gap_stats.hpp"

static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

int main(int argc, const char* argv[]) {
    if (argc < 5) {
        std::cerr << "usage: bc_gap_stats hist out.txt L files... | bc_gap_stats run out_root T D J [flags]\n";
        return 2;
    }
    const std::string mode = argv[1];
    if (mode == "hist") {
        const int L = std::atoi(argv[3]);
        std::vector<long long> hist((size_t)L / 2, 0);
        long long n = 0;
        for (int i = 4; i < argc; ++i, ++n) {
            std::string text;
            if (!bcgap::read_file(argv[i], text)) { std::cerr << "bc_gap_stats: cannot read " << argv[i] << std::endl; return 1; }
            bcgapstats::accumulate(hist, text, L);
        }
        return bcgap::write_file(argv[2], bcgapstats::run_file(hist, n)) ? 0 : 1;
    }
    if (mode != "run" || argc < 6) return 2;
    const std::string out_root = argv[2];
    const double T = std::atof(argv[3]), D = std::atof(argv[4]), J = std::atof(argv[5]);
    std::vector<int> Ls = {48, 64};  // gap_size_statistics.cpp:173
    int runs = 100, samples = 20, device = 0, kind = 0, boundary = BC_PERIODIC;
    int64_t burn = -1, sps = 30, cluster_every = 3;
    uint64_t seed = 1;
    for (int i = 6; i < argc; ++i) {
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
        else if (a == "--boundary") boundary = std::string(next()) == "open" ? BC_OPEN : BC_PERIODIC;
        else { std::cerr << "bc_gap_stats: unknown flag " << a << std::endl; return 2; }
    }
    const double p_fk = 1 - std::exp(-2 * (1 / T) * J);
    const uint32_t thr = (uint32_t)std::llround(std::fmin(std::fmax(p_fk, 0.0), 1.0) * 4294967295.0);
    mkdir(out_root.c_str(), 0777);
    for (int L : Ls) {
        const int64_t N = (int64_t)L * L, b = burn >= 0 ? burn : 20 * N;
        mkdir((out_root + "/" + std::to_string(L)).c_str(), 0777);
        bc_engine* e = nullptr;
        if (bc_create(&e, L, boundary, runs, BC_STORAGE_INT8, device, seed + (uint64_t)L) != BC_OK) {
            std::cerr << "bc_gap_stats: " << bc_last_error(nullptr) << std::endl;
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
        std::vector<std::vector<long long>> hist((size_t)runs, std::vector<long long>((size_t)L / 2, 0));
        std::vector<int8_t> spins((size_t)runs * N);
        bool ok = bc_set_params(e, -1, T, D, J) == BC_OK && bc_init_random(e, -1) == BC_OK && evolve(b);
        for (int k = 0; ok && k < samples; ++k) {
            ok = evolve(sps) && bc_download_all(e, spins.data()) == BC_OK;
            for (int r = 0; ok && r < runs; ++r) {
                const int8_t* s = spins.data() + (size_t)r * N;
                std::string text;
                if (kind == 0) text = bcgap::spin_clusters_to_gap(s, L, boundary == BC_PERIODIC);
                else {
                    const uint32_t k0 = (uint32_t)(seed + (uint64_t)L), k1 = (uint32_t)((seed + (uint64_t)L) >> 32) ^ (uint32_t)r;
                    text = bcgap::clusters_to_gap(s, L, boundary == BC_PERIODIC, [&](int64_t site, int dir) {
                        uint32_t w[4];
                        philox((uint32_t)site, (uint32_t)((uint64_t)site >> 32), (uint32_t)k, (3u << 1) | (uint32_t)dir, k0, k1, w);
                        return w[0] < thr;
                    });
                }
                bcgapstats::accumulate(hist[(size_t)r], text, L);
            }
        }
        if (!ok) { std::cerr << "bc_gap_stats: " << bc_last_error(e) << std::endl; bc_destroy(e); return 1; }
        bc_destroy(e);
        for (int r = 0; r < runs; ++r)
            bcgap::write_file(out_root + "/" + std::to_string(L) + "/" + std::to_string(r) + ".txt", bcgapstats::run_file(hist[(size_t)r], samples));
    }
    return 0;
}
