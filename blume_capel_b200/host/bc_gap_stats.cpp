// bc_gap_stats -- the reference's gap-size histogram per run (corner_contribution/gap_size_statistics.cpp)
//   bc_gap_stats hist out.txt L file1 [file2 ...]       CPU only: cluster files of one run -> histogram file
//   bc_gap_stats run out_root T D J [--kind spin|fk] [--L 48,64] [--runs 100] [--samples n] [--burn-sweeps n]
//                [--sweeps-per-sample n] [--cluster-every k] [--seed s] [--device d] [--boundary periodic|open]
//                [--corner out.csv]
//       GPU: every run is one replica; per sample the clusters are labelled, sorted and walked ON THE GPU
//       (bc_gap_histogram: no lattice download, no cluster text) and one histogram per run comes back; writes
//       {out_root}/{L}/{run}.txt like the reference tool.  --boundary open selects the open-boundary variant of the
//       histogram (unfolded gaps, no wrap-around gap, L-1 bins).  --corner also writes the table of
//       gap_size_analysis.py:24-48 (Batch,L,Corner_Contribution,Error; batches of 10 runs) from the histograms.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <sys/stat.h>

#include "../../include/bc_engine.h"
#include "gap_format.hpp"
#include "gap_stats.hpp"

// gap_size_analysis.py:5-11: corner contribution of one run from its gap histogram,
//   sum_{i < L/2} sum_{j >= i} hist[j]  /  (L^2 * nsamples)
static double corner_contribution(const std::vector<long long>& hist, int L, long long nsamples) {
    double tail = 0, total = 0;
    for (size_t i = hist.size(); i-- > 0;) { tail += (double)hist[i]; total += tail; }
    return total / ((double)L * L * (double)nsamples);
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
    std::string corner_path;
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
        else if (a == "--corner") corner_path = next();
        else { std::cerr << "bc_gap_stats: unknown flag " << a << std::endl; return 2; }
    }
    mkdir(out_root.c_str(), 0777);
    const int variant = boundary == BC_OPEN ? 1 : 0;
    std::string corner = "Batch,L,Corner_Contribution,Error\n";
    for (int L : Ls) {
        const int64_t N = (int64_t)L * L, b = burn >= 0 ? burn : 20 * N;
        const size_t bins = variant == 0 ? (size_t)L / 2 : (size_t)L - 1;
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
        std::vector<int64_t> hist((size_t)runs * bins, 0);
        bool ok = bc_set_params(e, -1, T, D, J) == BC_OK && bc_init_random(e, -1) == BC_OK && evolve(b);
        for (int k = 0; ok && k < samples; ++k)  // FK bonds: measurement counter = the sample index, like bc_run's fk/ dumps
            ok = evolve(sps) && bc_gap_histogram(e, kind, k, variant, hist.data()) == BC_OK;
        if (!ok) { std::cerr << "bc_gap_stats: " << bc_last_error(e) << std::endl; bc_destroy(e); return 1; }
        bc_destroy(e);
        std::vector<double> cc;
        for (int r = 0; r < runs; ++r) {
            const std::vector<long long> h(hist.begin() + (size_t)r * bins, hist.begin() + (size_t)(r + 1) * bins);
            bcgap::write_file(out_root + "/" + std::to_string(L) + "/" + std::to_string(r) + ".txt", bcgapstats::run_file(h, samples));
            cc.push_back(corner_contribution(h, L, samples));
        }
        for (size_t b0 = 0; b0 + 10 <= cc.size(); b0 += 10) {  // gap_size_analysis.py:24-48: batches of 10 runs
            double mean = 0, var = 0;
            for (size_t i = b0; i < b0 + 10; ++i) mean += cc[i] / 10;
            for (size_t i = b0; i < b0 + 10; ++i) var += (cc[i] - mean) * (cc[i] - mean) / 9;
            corner += std::to_string(b0 / 10) + "," + std::to_string(L) + "," + std::to_string(mean) + "," + std::to_string(std::sqrt(var / 10)) + "\n";
        }
    }
    if (!corner_path.empty()) bcgap::write_file(corner_path, corner);
    return 0;
}
