// bc_runner -- the reference's whole data-collection campaign in one process on one GPU.
//
// The reference collects its data with job_scripts/runner.sh: one binary per lattice size
// (-DL_MACRO, runner.sh:12-15), one Slurm array task per run index (runner.sh:17-20), every task executing
//     compiled/<name>-<L> run root 0 nsamples T D J                 (job_scripts/run.sh:10-11)
// after burner.sh produced one burn file per size.  Here every size is one engine handle whose replicas are the
// run indices (replica r draws from Philox key (seed_lo, seed_hi ^ r)), one host thread per size drives its
// handle (distinct handles may be driven from distinct threads, bc_engine.h), and the files are the reference's:
//     ./{root}/burn/{L}_burn.txt   ./{root}/spin/{L}/{run}/{i}.txt   ./{root}/fk/{L}/{run}/{i}.txt
// in the gap format of main.cpp:232-257, so magnet.cpp, gamma_nu.cpp, gap_size_statistics.cpp and square.py
// read them unchanged.  Run r of a size is bit-identical to
//     bc_run r root burn nsamples T D J --L L --seed (seed ^ (r << 32))   with the same flags.
//
//     bc_runner root T D J [--L 8,12,16,24,32,48,64] [--runs n|n1,n2,...] [--nsamples n|n1,n2,...]
//               [--burn 0|1|2] [--burn-sweeps n] [--sweeps-per-sample n] [--cluster-every k] [--seed s]
//               [--boundary periodic|open] [--device d] [--mkdir] [--no-fk] [--no-dumps] [--obs]
//   --L, --runs, --nsamples default to runner.sh:4-6 (sizes 8..64, 100/100/50/50/50/50/50 runs,
//   15000/15000/15000/15000/5000/5000/1500 samples; a size runner.sh does not list gets 50 runs of 1500).  --burn as in main.cpp:337-355: 0 = every run starts from
//   the size's burn file (what the reference's array tasks do), 1 = burn in and write the burn file (replica 0's
//   lattice), 2 = burn every run in independently, then collect.  Sweeps per burn-in / per sample default to the
//   reference's attempt counts (1500*N resp. 9*N step() calls of 3 sweeps each, main.cpp:342,362,159).
//   --lockstep (needs --sweeps-per-sample, every L >= 128 and a multiple of 32, equal --nsamples): the sizes
//   advance together with ONE kernel launch per half-sweep (bc_sweep_group) instead of one thread per size.
//   --obs writes ./{root}/obs/{L}.csv (run,sample,sweep,M,Q,B).  One JSON line per size (and point) goes to stdout.
//   --T-list a,b,... / --D-list x,y,...: a grid of (T, D) points instead of the positional pair (BASELINE config 3:
//   the finite-size-scaling ladder over a grid around the tricritical point); every point gets its own tree
//   ./{root}/p{k}/... (k = index in the T-major grid), the replicas of a size are points x runs (replica =
//   point * runs + run), and --shard k/n keeps the points with index % n == k for this process (one process per
//   GPU, --device k; the Philox keys follow the process's own replica numbering, so a sharded campaign is
//   reproducible for the same --shard, not file-identical to the unsharded one).  --cluster-every needs one common T (the bond probability is per handle).
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "../../include/bc_engine.h"
#include "gap_format.hpp"
#include "host_util.hpp"

namespace {

std::vector<long long> int_list(const std::string& s) {
    std::vector<long long> v;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) if (!tok.empty()) v.push_back(std::atoll(tok.c_str()));
    return v;
}

struct Point {
    double T, D;
    int index;          // position in the full T-major grid (names the point's directory)
    std::string root;   // ./{root} for a single point, ./{root}/p{index} on a grid
};

struct Options {
    std::vector<Point> points;
    std::string root = "NONE";
    double T = 0.608, D = 1.966, J = 1.0;  // main.cpp:325-333
    std::vector<long long> sizes{8, 12, 16, 24, 32, 48, 64}, runs{100, 100, 50, 50, 50, 50, 50},
        nsamples{15000, 15000, 15000, 15000, 5000, 5000, 1500};  // runner.sh:4-6
    int burn = 2, device = 0, boundary = BC_PERIODIC;
    long long burn_sweeps = -1, sweeps_per_sample = -1, cluster_every = 0;
    uint64_t seed = 0;
    bool write_fk = true, do_mkdir = false, dumps = true, obs = false, lockstep = false;
};

// one lattice size: its handle, its runs (replicas) and its running sums
struct Size {
    int L = 0, n_runs = 0, n_samples = 0;
    long long burn_sweeps = 0, sweeps_per_sample = 0;
    bc_engine* eng = nullptr;
    std::vector<int8_t> spins;   // n_runs lattices, replica-major
    std::vector<bc_obs> obs;     // n_runs records
    std::vector<FILE*> obs_file;                 // per point
    std::vector<double> sE, sAm, sQ, sM2, sM4;   // per point
    std::string error;
};

bool fail(Size& s, const char* what) {
    s.error = std::string(what) + ": " + bc_last_error(s.eng);
    return false;
}

// n sweeps with one cluster move after every cluster_every sweeps (the counterpart of main.cpp:158-163), as bc_run
bool evolve(const Options& o, Size& s, long long n) {
    if (o.cluster_every <= 0) return bc_sweep(s.eng, n, 0, nullptr) >= 0 || fail(s, "bc_sweep");
    while (n > 0) {
        const long long k = n < o.cluster_every ? n : o.cluster_every;
        if (bc_sweep(s.eng, k, 0, nullptr) < 0) return fail(s, "bc_sweep");
        if (k == o.cluster_every && bc_cluster_update(s.eng, 1) != BC_OK) return fail(s, "bc_cluster_update");
        n -= k;
    }
    return true;
}

std::string burn_path(const Point& pt, int L) { return "./" + pt.root + "/burn/" + std::to_string(L) + "_burn.txt"; }

bool setup(const Options& o, Size& s) {
    const long long N = (long long)s.L * s.L;
    s.burn_sweeps = o.burn_sweeps >= 0 ? o.burn_sweeps : 3 * 1500 * N;               // main.cpp:342
    s.sweeps_per_sample = o.sweeps_per_sample >= 0 ? o.sweeps_per_sample : 3 * 9 * N;  // main.cpp:362
    const int n_points = (int)o.points.size(), n_rep = n_points * s.n_runs;
    if (bc_create(&s.eng, s.L, o.boundary, n_rep, BC_STORAGE_INT8, o.device, o.seed) != BC_OK) {
        s.error = std::string("bc_create: ") + bc_last_error(nullptr);
        return false;
    }
    if (n_points == 1) {
        if (bc_set_params(s.eng, -1, o.points[0].T, o.points[0].D, o.J) != BC_OK) return fail(s, "bc_set_params");
    } else {
        for (int p = 0; p < n_points; ++p)
            for (int r = 0; r < s.n_runs; ++r)
                if (bc_set_params(s.eng, p * s.n_runs + r, o.points[p].T, o.points[p].D, o.J) != BC_OK) return fail(s, "bc_set_params");
    }
    s.spins.resize((size_t)n_rep * N);
    s.obs.resize((size_t)n_rep);
    s.obs_file.assign((size_t)n_points, nullptr);
    s.sE.assign((size_t)n_points, 0.0); s.sAm = s.sQ = s.sM2 = s.sM4 = s.sE;
    if (o.do_mkdir) {  // gen_files.py:5-22 creates these for the reference
        for (const Point& pt : o.points) {
            mkdirs("./" + pt.root + "/burn");
            for (int r = 0; r < s.n_runs; ++r) {
                mkdirs("./" + pt.root + "/spin/" + std::to_string(s.L) + "/" + std::to_string(r));
                if (o.write_fk) mkdirs("./" + pt.root + "/fk/" + std::to_string(s.L) + "/" + std::to_string(r));
            }
            if (o.obs) mkdirs("./" + pt.root + "/obs");
        }
    }
    const bool periodic = o.boundary == BC_PERIODIC;
    if (o.burn >= 1) {
        if (bc_init_random(s.eng, -1) != BC_OK) return fail(s, "bc_init_random");
        if (!evolve(o, s, s.burn_sweeps)) return false;
        for (int p = 0; p < n_points; ++p) {  // the burn file of a point is its run 0
            if (bc_download(s.eng, p * s.n_runs, s.spins.data()) != BC_OK) return fail(s, "bc_download");
            if (!bcgap::write_file(burn_path(o.points[p], s.L), bcgap::spin_clusters_to_gap(s.spins.data(), s.L, periodic)))
                std::cerr << "bc_runner: cannot write " << burn_path(o.points[p], s.L) << std::endl;
        }
    } else {
        // every run starts from the size's burn file, like the reference's array tasks (main.cpp:358); the sweep
        // counter continues after the nominal burn-in so that no run replays the burn-in's random numbers
        for (int p = 0; p < n_points; ++p) {
            int8_t* first = &s.spins[(size_t)p * s.n_runs * N];
            std::string text;
            if (!bcgap::read_file(burn_path(o.points[p], s.L), text) || !bcgap::gap_to_lattice(text, s.L, first)) {
                s.error = "cannot read burn file " + burn_path(o.points[p], s.L);
                return false;
            }
            for (int r = 1; r < s.n_runs; ++r) std::memcpy(first + (size_t)r * N, first, (size_t)N);
        }
        if (bc_upload_all(s.eng, s.spins.data()) != BC_OK) return fail(s, "bc_upload_all");
        if (bc_set_sweep_counter(s.eng, s.burn_sweeps) != BC_OK) return fail(s, "bc_set_sweep_counter");
        if (o.cluster_every > 0 && bc_set_cluster_counter(s.eng, s.burn_sweeps / o.cluster_every) != BC_OK)
            return fail(s, "bc_set_cluster_counter");
    }
    if (o.obs)
        for (int p = 0; p < n_points; ++p) {
            s.obs_file[(size_t)p] = std::fopen(("./" + o.points[p].root + "/obs/" + std::to_string(s.L) + ".csv").c_str(), "w");
            if (s.obs_file[(size_t)p]) std::fprintf(s.obs_file[(size_t)p], "run,sample,sweep,M,Q,B\n");
        }
    return true;
}

// after the sweeps of sample i: observables, lattices, dumps of every run
bool record(const Options& o, Size& s, int i) {
    const long long N = (long long)s.L * s.L;
    if (bc_measure(s.eng, s.obs.data()) != BC_OK) return fail(s, "bc_measure");
    const std::string name = std::to_string(i) + ".txt", sL = std::to_string(s.L);
    std::vector<uint32_t> l_sites, l_starts;
    std::vector<int8_t> l_signs;
    if (o.dumps) { l_sites.resize((size_t)N); l_starts.resize((size_t)N + 1); l_signs.resize((size_t)N); }
    for (int p = 0; p < (int)o.points.size(); ++p) {
        const Point& pt = o.points[(size_t)p];
        for (int r = 0; r < s.n_runs; ++r) {
            const int replica = p * s.n_runs + r;
            const bc_obs& ob = s.obs[(size_t)replica];
            const double E = (-o.J * (double)ob.B + pt.D * (double)ob.Q) / (double)N, m = (double)ob.M / (double)N,
                         q = (double)ob.Q / (double)N;
            s.sE[p] += E; s.sAm[p] += std::fabs(m); s.sQ[p] += q; s.sM2[p] += m * m; s.sM4[p] += m * m * m * m;
            if (s.obs_file[(size_t)p])
                std::fprintf(s.obs_file[(size_t)p], "%d,%d,%lld,%lld,%lld,%lld\n", r, i, (long long)ob.sweep, (long long)ob.M,
                             (long long)ob.Q, (long long)ob.B);
        }
    }
    if (!o.dumps) return true;
    // Dumps: the clusters of every run are formed on the GPU (one labelling + one sort for the whole handle, cached
    // by the library across the per-run fetches); the host prints decimals.  spin/ first for all runs, then fk/.
    for (int kind = 0; kind < (o.write_fk ? 2 : 1); ++kind)
        for (int p = 0; p < (int)o.points.size(); ++p)
            for (int r = 0; r < s.n_runs; ++r) {
                int64_t n_sites = 0, n_clusters = 0;
                if (bc_cluster_lists(s.eng, p * s.n_runs + r, kind, kind ? i : 0, l_sites.data(), l_starts.data(), l_signs.data(),
                                     &n_sites, &n_clusters) != BC_OK) return fail(s, "bc_cluster_lists");
                bcgap::write_file("./" + o.points[(size_t)p].root + (kind ? "/fk/" : "/spin/") + sL + "/" + std::to_string(r) + "/" + name,
                                  bcgap::lists_to_gap(l_sites.data(), l_starts.data(), l_signs.data(), n_clusters, 1));  // main.cpp:369-372
            }
    return true;
}

void summary(const Options& o, const Size& s) {
    const double n = (double)s.n_samples * s.n_runs;
    if (n <= 0) return;
    for (size_t p = 0; p < o.points.size(); ++p) {
        const double m2 = s.sM2[p] / n, am = s.sAm[p] / n, N = (double)s.L * s.L;
        std::printf("{\"L\": %d, \"point\": %d, \"T\": %.10g, \"D\": %.10g, \"J\": %.10g, \"runs\": %d, \"samples\": %d, "
                    "\"sweeps_per_sample\": %lld, \"E_per_site\": %.10g, \"abs_m\": %.10g, \"q\": %.10g, \"U4\": %.10g, "
                    "\"chi\": %.10g, \"seed\": %llu}\n",
                    s.L, o.points[p].index, o.points[p].T, o.points[p].D, o.J, s.n_runs, s.n_samples, s.sweeps_per_sample,
                    s.sE[p] / n, am, s.sQ[p] / n, 1 - (s.sM4[p] / n) / (3 * m2 * m2), N * (m2 - am * am),
                    (unsigned long long)o.seed);
    }
}

}  // namespace

int main(int argc, const char* argv[]) {
    Options o;
    if (argc < 5 || std::strncmp(argv[1], "--", 2) == 0) {
        std::cerr << "usage: bc_runner root T D J [--L a,b,...] [--runs n|a,b,...] [--nsamples n|a,b,...] [--burn 0|1|2]\n"
                     "                 [--burn-sweeps n] [--sweeps-per-sample n] [--cluster-every k] [--seed s]\n"
                     "                 [--boundary periodic|open] [--device d] [--mkdir] [--no-fk] [--no-dumps] [--obs]\n"
                     "                 [--lockstep] [--T-list a,b,...] [--D-list x,y,...] [--shard k/n]\n";
        return 2;
    }
    o.root = argv[1]; o.T = std::atof(argv[2]); o.D = std::atof(argv[3]); o.J = std::atof(argv[4]);
    bool have_seed = false, have_runs = false, have_nsamples = false;
    std::string t_list, d_list;
    int shard_k = 0, shard_n = 1;
    for (int argi = 5; argi < argc; ++argi) {
        const std::string a = argv[argi];
        auto next = [&]() -> const char* { return argi + 1 < argc ? argv[++argi] : ""; };
        if (a == "--L") o.sizes = int_list(next());
        else if (a == "--runs") { o.runs = int_list(next()); have_runs = true; }
        else if (a == "--nsamples") { o.nsamples = int_list(next()); have_nsamples = true; }
        else if (a == "--burn") o.burn = std::atoi(next());
        else if (a == "--burn-sweeps") o.burn_sweeps = std::atoll(next());
        else if (a == "--sweeps-per-sample") o.sweeps_per_sample = std::atoll(next());
        else if (a == "--cluster-every") o.cluster_every = std::atoll(next());
        else if (a == "--seed") { o.seed = std::strtoull(next(), nullptr, 0); have_seed = true; }
        else if (a == "--boundary") o.boundary = std::string(next()) == "open" ? BC_OPEN : BC_PERIODIC;
        else if (a == "--device") o.device = std::atoi(next());
        else if (a == "--mkdir") o.do_mkdir = true;
        else if (a == "--no-fk") o.write_fk = false;
        else if (a == "--no-dumps") o.dumps = false;
        else if (a == "--obs") o.obs = true;
        else if (a == "--lockstep") o.lockstep = true;
        else if (a == "--T-list") t_list = next();
        else if (a == "--D-list") d_list = next();
        else if (a == "--shard") { if (std::sscanf(next(), "%d/%d", &shard_k, &shard_n) != 2) shard_n = 0; }
        else { std::cerr << "bc_runner: unknown flag " << a << std::endl; return 2; }
    }
    if (!have_seed) { std::random_device rd; o.seed = ((uint64_t)rd() << 32) | rd(); }  // like main.cpp:307-309
    {  // the (T, D) points of this process
        std::vector<double> Ts, Ds;
        auto dbl_list = [](const std::string& str, std::vector<double>& v) {
            std::stringstream ss(str);
            std::string tok;
            while (std::getline(ss, tok, ',')) if (!tok.empty()) v.push_back(std::atof(tok.c_str()));
        };
        dbl_list(t_list, Ts); dbl_list(d_list, Ds);
        if (Ts.empty()) Ts.push_back(o.T);
        if (Ds.empty()) Ds.push_back(o.D);
        if (shard_n < 1 || shard_k < 0 || shard_k >= shard_n) { std::cerr << "bc_runner: --shard k/n needs 0 <= k < n" << std::endl; return 2; }
        const bool grid = Ts.size() * Ds.size() > 1;
        int index = 0;
        for (double t : Ts)
            for (double d : Ds) {
                if (index % shard_n == shard_k) o.points.push_back({t, d, index, grid ? o.root + "/p" + std::to_string(index) : o.root});
                ++index;
            }
        if (o.points.empty()) { std::cerr << "bc_runner: no (T, D) point falls to shard " << shard_k << "/" << shard_n << std::endl; return 2; }
    }
    const size_t n_sizes = o.sizes.size();
    if (!have_runs || !have_nsamples) {  // per-size defaults of runner.sh:4-6; other sizes as its largest one
        static const long long table[7][3] = {{8, 100, 15000}, {12, 100, 15000}, {16, 50, 15000}, {24, 50, 15000},
                                              {32, 50, 5000}, {48, 50, 5000}, {64, 50, 1500}};
        std::vector<long long> runs, nsamples;
        for (long long L : o.sizes) {
            long long r = 50, n = 1500;
            for (const auto& row : table) if (row[0] == L) { r = row[1]; n = row[2]; }
            runs.push_back(r); nsamples.push_back(n);
        }
        if (!have_runs) o.runs = runs;
        if (!have_nsamples) o.nsamples = nsamples;
    }
    auto per_size = [&](std::vector<long long>& v, const char* what) {
        if (v.size() == 1) v.assign(n_sizes, v[0]);
        if (v.size() != n_sizes) { std::cerr << "bc_runner: " << what << " needs one value or one per size" << std::endl; return false; }
        return true;
    };
    if (n_sizes == 0 || !per_size(o.runs, "--runs") || !per_size(o.nsamples, "--nsamples")) return 2;

    std::vector<Size> sizes(n_sizes);
    for (size_t k = 0; k < n_sizes; ++k) {
        sizes[k].L = (int)o.sizes[k]; sizes[k].n_runs = (int)o.runs[k]; sizes[k].n_samples = (int)o.nsamples[k];
        if (sizes[k].n_runs < 1 || sizes[k].n_samples < 0) { std::cerr << "bc_runner: bad --runs / --nsamples" << std::endl; return 2; }
    }
    if (o.lockstep) {
        if (o.sweeps_per_sample < 0 || o.cluster_every > 0 || n_sizes > 8) {
            std::cerr << "bc_runner: --lockstep needs --sweeps-per-sample, no --cluster-every and at most 8 sizes" << std::endl;
            return 2;
        }
        for (const Size& s : sizes)
            if (s.L % 32 || s.L < 128 || s.n_samples != sizes[0].n_samples) {
                std::cerr << "bc_runner: --lockstep needs L >= 128, L % 32 == 0 and equal --nsamples" << std::endl;
                return 2;
            }
    }

    // ---- burn-in (or burn-file load) of every size, one host thread per size
    std::atomic<bool> ok{true};
    {
        std::vector<std::thread> th;
        for (Size& s : sizes) th.emplace_back([&o, &s, &ok]() { if (!setup(o, s)) ok = false; });
        for (std::thread& t : th) t.join();
    }
    auto finish = [&](int rc) {
        for (Size& s : sizes) {
            if (!s.error.empty()) std::cerr << "bc_runner: L=" << s.L << ": " << s.error << std::endl;
            for (FILE* f : s.obs_file) if (f) std::fclose(f);
            if (s.eng) bc_destroy(s.eng);
        }
        return rc;
    };
    if (!ok) return finish(1);
    if (o.burn == 1) return finish(0);  // burn only (burner.sh)

    // ---- collection
    if (o.lockstep) {
        std::vector<bc_engine*> group;
        for (Size& s : sizes) group.push_back(s.eng);
        for (int i = 0; i < sizes[0].n_samples && ok; ++i) {
            if (bc_sweep_group(group.data(), (int)group.size(), o.sweeps_per_sample, 0) < 0) {
                sizes[0].error = std::string("bc_sweep_group: ") + bc_last_error(group[0]);
                ok = false;
                break;
            }
            std::vector<std::thread> th;  // dumps of the sizes in parallel
            for (Size& s : sizes) th.emplace_back([&o, &s, &ok, i]() { if (!record(o, s, i)) ok = false; });
            for (std::thread& t : th) t.join();
        }
    } else {
        std::vector<std::thread> th;
        for (Size& s : sizes)
            th.emplace_back([&o, &s, &ok]() {
                for (int i = 0; i < s.n_samples; ++i)
                    if (!evolve(o, s, s.sweeps_per_sample) || !record(o, s, i)) { ok = false; return; }
            });
        for (std::thread& t : th) t.join();
    }
    if (!ok) return finish(1);
    for (const Size& s : sizes) summary(o, s);
    return finish(0);
}
