// bc_run -- host driver with the reference's run interface, on top of the C ABI (bc_engine.h).
//
// Same seven positionals as /root/reference/main.cpp:316-324 (documented in job_scripts/run.sh:10-11):
//     bc_run run root burn nsteps T D J [flags]
//   burn = 0: load ./{root}/burn/{L}_burn.txt and collect; 1: burn in, write the burn file, stop;
//   >= 2: burn in, then collect (main.cpp:337-355).  With no arguments the reference's defaults
//   apply (main.cpp:325-333).  Files, all in the reference's gap format (main.cpp:232-257):
//     ./{root}/burn/{L}_burn.txt, ./{root}/spin/{L}/{run}/{i}.txt, ./{root}/fk/{L}/{run}/{i}.txt
// What differs: L is a flag instead of -DL_MACRO; the lattice update is the GPU checkerboard
// Metropolis sweep (no Wolff move); sweeps per burn-in / per sample default to the reference's
// attempt counts (1500*N resp. 9*N step() calls of 3 sweeps each, main.cpp:342,362,159) and can be
// set with flags; --cluster-every k adds one GPU Swendsen-Wang cluster move per k sweeps (the reference
// interleaves Wolff clusters, main.cpp:162); --seed makes the run reproducible (the reference seeds from
// random_device).
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "../../include/bc_engine.h"
#include "gap_format.hpp"
#include "host_util.hpp"

namespace {

int die(bc_engine* e, const char* what) {
    std::cerr << "bc_run: " << what << ": " << bc_last_error(e) << std::endl;
    if (e) bc_destroy(e);
    return 1;
}

}  // namespace

int main(int argc, const char* argv[]) {
    // ---- the reference's positionals and defaults (main.cpp:316-333)
    int run = 0, burn = 2, nsteps = 1500;
    std::string root = "NONE";
    double T = 0.608, D = 1.966, J = 1.0;
    int argi = 1;
    if (argc > 1 && std::strncmp(argv[1], "--", 2) != 0) {
        if (argc < 8) {
            std::cerr << "usage: bc_run run root burn nsteps T D J [--L n] [--boundary periodic|open] [--seed s]\n"
                         "              [--device d] [--burn-sweeps n] [--sweeps-per-sample n] [--obs file.csv]\n"
                         "              [--cluster-every k] [--no-fk] [--mkdir] [--save-state f] [--load-state f] [--time-dumps]\n";
            return 2;
        }
        run = std::atoi(argv[1]); root = argv[2]; burn = std::atoi(argv[3]); nsteps = std::atoi(argv[4]);
        T = std::atof(argv[5]); D = std::atof(argv[6]); J = std::atof(argv[7]);
        argi = 8;
    }
    int L = 64, device = 0, boundary = BC_PERIODIC;
    bool have_seed = false, write_fk = true, do_mkdir = false, time_dumps = false;
    uint64_t seed = 0;
    int64_t burn_sweeps = -1, sweeps_per_sample = -1, cluster_every = 0;
    std::string obs_path, save_state, load_state;
    for (; argi < argc; ++argi) {
        const std::string a = argv[argi];
        auto next = [&]() -> const char* { return argi + 1 < argc ? argv[++argi] : ""; };
        if (a == "--L") L = std::atoi(next());
        else if (a == "--boundary") boundary = std::string(next()) == "open" ? BC_OPEN : BC_PERIODIC;
        else if (a == "--seed") { seed = std::strtoull(next(), nullptr, 0); have_seed = true; }
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--burn-sweeps") burn_sweeps = std::atoll(next());
        else if (a == "--sweeps-per-sample") sweeps_per_sample = std::atoll(next());
        else if (a == "--cluster-every") cluster_every = std::atoll(next());
        else if (a == "--obs") obs_path = next();
        else if (a == "--save-state") save_state = next();
        else if (a == "--load-state") load_state = next();
        else if (a == "--no-fk") write_fk = false;
        else if (a == "--mkdir") do_mkdir = true;
        else if (a == "--time-dumps") time_dumps = true;
        else { std::cerr << "bc_run: unknown flag " << a << std::endl; return 2; }
    }
    if (!have_seed) { std::random_device rd; seed = ((uint64_t)rd() << 32) | rd(); }  // like main.cpp:307-309
    const int64_t N = (int64_t)L * L;
    if (burn_sweeps < 0) burn_sweeps = 3 * 1500 * N;        // 1500*N step() calls (main.cpp:342), 3 sweeps each
    if (sweeps_per_sample < 0) sweeps_per_sample = 3 * 9 * N;  // 9*N step() calls (main.cpp:362)

    const std::string sL = std::to_string(L), sRun = std::to_string(run);
    const std::string burn_path = "./" + root + "/burn/" + sL + "_burn.txt";
    const std::string spin_dir = "./" + root + "/spin/" + sL + "/" + sRun + "/";
    const std::string fk_dir = "./" + root + "/fk/" + sL + "/" + sRun + "/";
    if (do_mkdir) {  // gen_files.py:5-22 creates these for the reference
        mkdirs("./" + root + "/burn");
        mkdirs(spin_dir);
        mkdirs(fk_dir);
    }

    bc_engine* eng = nullptr;
    if (bc_create(&eng, L, boundary, 1, BC_STORAGE_INT8, device, seed) != BC_OK) return die(nullptr, "bc_create");
    // n sweeps, with one Swendsen-Wang cluster move after every `cluster_every` sweeps if requested: the
    // counterpart of the Wolff move the reference interleaves in step() (main.cpp:158-163)
    auto evolve = [&](int64_t n) -> bool {
        if (cluster_every <= 0) return bc_sweep(eng, n, 0, nullptr) >= 0;
        while (n > 0) {
            const int64_t k = n < cluster_every ? n : cluster_every;
            if (bc_sweep(eng, k, 0, nullptr) < 0) return false;
            if (k == cluster_every && bc_cluster_update(eng, 1) != BC_OK) return false;
            n -= k;
        }
        return true;
    };
    if (bc_set_params(eng, -1, T, D, J) != BC_OK) return die(eng, "bc_set_params");
    std::vector<int8_t> spins((size_t)N);
    const bool periodic = boundary == BC_PERIODIC;

    if (burn >= 1) {
        std::cout << "burning " + sL << std::endl;  // main.cpp:338
        if (bc_init_random(eng, 0) != BC_OK) return die(eng, "bc_init_random");
        if (!evolve(burn_sweeps)) return die(eng, "bc_sweep");
        if (bc_download(eng, 0, spins.data()) != BC_OK) return die(eng, "bc_download");
        if (!bcgap::write_file(burn_path, bcgap::spin_clusters_to_gap(spins.data(), L, periodic)))
            std::cerr << "bc_run: cannot write " << burn_path << " (the reference fails silently here)" << std::endl;
        std::cout << "burnt " + sL << std::endl;  // main.cpp:351
        if (burn == 1) { bc_destroy(eng); return 0; }
    }

    // If not burning in, grab the burned-in lattice (main.cpp:358) -- or an exact checkpoint: the reference saves no
    // RNG state; here (lattice, seed, sweep counter, cluster counter) is the whole state of the run
    struct StateHeader { char magic[8]; int32_t L, boundary; uint64_t seed; int64_t sweep, cluster; };
    int first_sample = 0;
    if (!load_state.empty()) {
        std::string raw;
        StateHeader h;
        if (!bcgap::read_file(load_state, raw) || raw.size() != sizeof(h) + sizeof(int32_t) + (size_t)N) {
            std::cerr << "bc_run: cannot read state file " << load_state << std::endl;
            bc_destroy(eng);
            return 1;
        }
        std::memcpy(&h, raw.data(), sizeof(h));
        int32_t done = 0;
        std::memcpy(&done, raw.data() + sizeof(h), sizeof(done));
        if (std::memcmp(h.magic, "BCSTATE1", 8) != 0 || h.L != L || h.boundary != boundary) {
            std::cerr << "bc_run: state file does not match this run" << std::endl;
            bc_destroy(eng);
            return 1;
        }
        bc_destroy(eng);  // the seed is part of the state: recreate the handle with it
        seed = h.seed;
        if (bc_create(&eng, L, boundary, 1, BC_STORAGE_INT8, device, seed) != BC_OK) return die(nullptr, "bc_create");
        if (bc_set_params(eng, -1, T, D, J) != BC_OK) return die(eng, "bc_set_params");
        std::memcpy(spins.data(), raw.data() + sizeof(h) + sizeof(done), (size_t)N);
        if (bc_upload(eng, 0, spins.data()) != BC_OK || bc_set_sweep_counter(eng, h.sweep) != BC_OK ||
            bc_set_cluster_counter(eng, h.cluster) != BC_OK) return die(eng, "restore state");
        first_sample = done;
    } else {
        std::string text;
        if (!bcgap::read_file(burn_path, text) || !bcgap::gap_to_lattice(text, L, spins.data())) {
            std::cerr << "bc_run: cannot read burn file " << burn_path << std::endl;
            bc_destroy(eng);
            return 1;
        }
        if (bc_upload(eng, 0, spins.data()) != BC_OK) return die(eng, "bc_upload");
        // burn = 0: continue the sweep counter after the nominal burn-in so that a collection run started
        // with the same --seed does not replay the random numbers the burn-in consumed
        if (burn < 1 && bc_set_sweep_counter(eng, burn_sweeps) != BC_OK) return die(eng, "bc_set_sweep_counter");
        // ... nor the bond and flip streams of the burn-in's cluster moves
        if (burn < 1 && cluster_every > 0 && bc_set_cluster_counter(eng, burn_sweeps / cluster_every) != BC_OK)
            return die(eng, "bc_set_cluster_counter");
    }

    FILE* obs_file = nullptr;
    if (!obs_path.empty()) {
        obs_file = std::fopen(obs_path.c_str(), "w");
        if (obs_file) std::fprintf(obs_file, "sample,sweep,M,Q,B,E_per_site,m,q\n");
    }
    double sE = 0, sAm = 0, sQ = 0, sM2 = 0, sM4 = 0;
    std::vector<uint32_t> l_sites((size_t)N), l_starts((size_t)N + 1);
    std::vector<int8_t> l_signs((size_t)N);
    for (int i = first_sample; i < nsteps; ++i) {
        bc_obs o;
        // measure_every = sweeps_per_sample: one fused measurement at the end of the block when aligned
        if (!evolve(sweeps_per_sample)) return die(eng, "bc_sweep");
        if (bc_measure(eng, &o) != BC_OK) return die(eng, "bc_measure");
        const double E = (-J * (double)o.B + D * (double)o.Q) / (double)N, m = (double)o.M / (double)N,
                     q = (double)o.Q / (double)N;
        sE += E; sAm += std::fabs(m); sQ += q; sM2 += m * m; sM4 += m * m * m * m;
        if (obs_file)
            std::fprintf(obs_file, "%d,%lld,%lld,%lld,%lld,%.12g,%.12g,%.12g\n", i, (long long)o.sweep, (long long)o.M,
                         (long long)o.Q, (long long)o.B, E, m, q);
        // the clusters are formed on the GPU (labelling + sort by cluster); the host only prints decimals
        const std::string name = std::to_string(i) + ".txt";
        for (int kind = 0; kind < (write_fk ? 2 : 1); ++kind) {  // main.cpp:369-370 (spin/), :371-372 (fk/, bond counter = sample index)
            int64_t n_sites = 0, n_clusters = 0;
            const auto t0 = std::chrono::steady_clock::now();
            if (bc_cluster_lists(eng, 0, kind, kind ? i : 0, l_sites.data(), l_starts.data(), l_signs.data(), &n_sites, &n_clusters) != BC_OK)
                return die(eng, "bc_cluster_lists");
            const auto t1 = std::chrono::steady_clock::now();
            const std::string text = bcgap::lists_to_gap(l_sites.data(), l_starts.data(), l_signs.data(), n_clusters);
            const auto t2 = std::chrono::steady_clock::now();
            bcgap::write_file((kind ? fk_dir : spin_dir) + name, text);
            const auto t3 = std::chrono::steady_clock::now();
            if (time_dumps) {
                auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
                std::fprintf(stderr, "{\"dump\": \"%s/%d\", \"L\": %d, \"clusters\": %lld, \"sites\": %lld, \"bytes\": %zu, "
                                     "\"gpu_lists_ms\": %.2f, \"format_ms\": %.2f, \"write_ms\": %.2f}\n", kind ? "fk" : "spin", i, L,
                             (long long)n_clusters, (long long)n_sites, text.size(), ms(t0, t1), ms(t1, t2), ms(t2, t3));
            }
        }
    }
    if (obs_file) std::fclose(obs_file);
    if (!save_state.empty()) {
        StateHeader h;
        std::memcpy(h.magic, "BCSTATE1", 8);
        h.L = L; h.boundary = boundary; h.seed = seed;
        if (bc_get_sweep_counter(eng, &h.sweep) != BC_OK || bc_get_cluster_counter(eng, &h.cluster) != BC_OK ||
            bc_download(eng, 0, spins.data()) != BC_OK) return die(eng, "save state");
        const int32_t done = nsteps;
        std::string raw((const char*)&h, sizeof(h));
        raw.append((const char*)&done, sizeof(done));
        raw.append((const char*)spins.data(), (size_t)N);
        if (!bcgap::write_file(save_state, raw)) std::cerr << "bc_run: cannot write " << save_state << std::endl;
    }
    if (nsteps > first_sample) {
        const double n = nsteps - first_sample, m2 = sM2 / n, am = sAm / n;
        std::printf("{\"L\": %d, \"T\": %.10g, \"D\": %.10g, \"J\": %.10g, \"samples\": %d, \"sweeps_per_sample\": %lld, "
                    "\"E_per_site\": %.10g, \"abs_m\": %.10g, \"q\": %.10g, \"U4\": %.10g, \"chi\": %.10g, \"seed\": %llu}\n",
                    L, T, D, J, nsteps, (long long)sweeps_per_sample, sE / n, am, sQ / n, 1 - (sM4 / n) / (3 * m2 * m2),
                    (double)N * (m2 - am * am), (unsigned long long)seed);
    }
    bc_destroy(eng);
    return 0;
}
