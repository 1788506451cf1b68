// fss_stats.hpp -- the reference's magnetic-susceptibility table, computed from magnetisation samples
// instead of from cluster files.  Restates corner_contribution/magnet.cpp:
//   per run   chi = (<m^2> - <|m|>^2) / L^2  with m = sum of spins of a sample      (magnet.cpp:57-74)
//   per batch of 10 runs: mean and population stdev / sqrt(10)                       (magnet.cpp:76-98,121-130)
//   CSV "batch,L,magnetic_susceptibility,standard_error", std::to_string formatting, (magnet.cpp:101,126-129)
//   the file ends with one extra newline (`file << output << endl`, magnet.cpp:133).
#pragma once
#include <cmath>
#include <string>
#include <vector>

namespace bcfss {

inline double run_chi(const std::vector<double>& m, int L) {
    double m1 = 0, m2 = 0;
    for (double v : m) { m1 += std::fabs(v); m2 += v * v; }
    const double n = (double)m.size();
    const double m1_avg = m1 / n, m2_avg = m2 / n;
    return (m2_avg - m1_avg * m1_avg) / ((double)L * L);
}

inline double mean(const std::vector<double>& d) {
    double s = 0;
    for (double v : d) s += v;
    return s / d.size();
}

inline double stdev(const std::vector<double>& d) {
    const double mu = mean(d);
    double s = 0;
    for (double v : d) s += (v - mu) * (v - mu);
    return std::sqrt(s / d.size());
}

// M[l][run] = magnetisation samples of that run; runs are batched by tens like the reference.
inline std::string magnet_csv(const std::vector<int>& Ls, const std::vector<std::vector<std::vector<double>>>& M) {
    std::string out = "batch,L,magnetic_susceptibility,standard_error\n";
    for (size_t li = 0; li < Ls.size(); ++li) {
        const int L = Ls[li];
        const size_t nruns = M[li].size();
        std::vector<double> chi(nruns);
        for (size_t r = 0; r < nruns; ++r) chi[r] = run_chi(M[li][r], L);
        for (size_t b = 0; b + 10 <= nruns; b += 10) {
            std::vector<double> batch(chi.begin() + b, chi.begin() + b + 10);
            out += std::to_string(b / 10) + "," + std::to_string(L) + "," + std::to_string(mean(batch)) + "," +
                   std::to_string(stdev(batch) / std::sqrt(10.0)) + "\n";
        }
    }
    out += "\n";
    return out;
}

// Cluster-size susceptibility table of corner_contribution/gamma_nu.cpp:
//   per sample S = (sum size^2 - max^2) / L^2 (gamma_nu.cpp:35-60), per run the mean over samples (:62-73),
//   per L the mean over runs and population stdev / sqrt(nruns) (:99-117); "L Chi SE" header, to_string
//   formatting, one extra newline at the end (:99,117,121).
struct ClusterSample { double sum_sq, max_size; };
inline std::string gamma_nu_table(const std::vector<int>& Ls, const std::vector<std::vector<std::vector<ClusterSample>>>& S) {
    std::string out = "L Chi SE\n";
    for (size_t li = 0; li < Ls.size(); ++li) {
        const int L = Ls[li];
        const size_t nruns = S[li].size();
        std::vector<double> stat(nruns);
        for (size_t r = 0; r < nruns; ++r) {
            double tot = 0;
            // +1: the reference counts the blank last line of every cluster file as a "cluster" of size -1
            // (count_size("") = -1 passes the `size == 0` filter, gamma_nu.cpp:24-33,44-49); kept for parity
            for (const ClusterSample& c : S[li][r]) tot += (c.sum_sq + 1.0 - c.max_size * c.max_size) / (L * L);
            stat[r] = tot / (double)S[li][r].size();
        }
        out += std::to_string(L) + " " + std::to_string(mean(stat)) + " " + std::to_string(stdev(stat) / std::sqrt((double)nruns)) + "\n";
    }
    out += "\n";
    return out;
}

}  // namespace bcfss
