// gap_format.hpp -- the reference's on-disk cluster "gap format", host side.
//
// Format (written by export_clusters, /root/reference/main.cpp:232-257; read by
// get_lattice_from_burn main.cpp:281-299, magnet.cpp:24-55, gamma_nu.cpp:24-60,
// gap_size_statistics.cpp:84-98, square.py:184-196): one line per cluster, "+ " or "- ", then
// the cluster's site indices (idx = x + y*L) sorted ascending and delta-encoded from 0, each
// followed by one space, then '\n'; the file ends with one extra '\n'.  Clusters appear in
// order of their smallest site.  Zero spins are never written.
//
// New code (union-find instead of the reference's explicit-stack DFS); same bytes for bond
// probability 1 ("spin" clusters), same distribution for FK clusters (each same-sign bond is
// open independently with probability p, which is what the DFS of main.cpp:179-230 samples).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <functional>
#include <numeric>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace bcgap {

// Cluster lists formed on the GPU (bc_cluster_lists: clusters in the order of their smallest site, members
// ascending) -> the file text of export_clusters (main.cpp:236-256).  Only decimals are printed here; clusters are
// dealt out to host threads in contiguous ranges of about equal member counts and the pieces are concatenated.
inline std::string lists_to_gap(const uint32_t* sites, const uint32_t* starts, const int8_t* signs, int64_t n_clusters,
                                int n_threads = 0) {
    const int64_t n_sites = n_clusters ? (int64_t)starts[n_clusters] : 0;
    if (n_threads <= 0) n_threads = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), 1 + n_sites / 200000);
    std::vector<std::string> piece((size_t)n_threads);
    auto work = [&](int t) {
        // cluster range of thread t: member offsets [t, t+1) * n_sites / n_threads, rounded to cluster starts
        const int64_t lo_off = n_sites * t / n_threads, hi_off = n_sites * (t + 1) / n_threads;
        const int64_t c0 = std::lower_bound(starts, starts + n_clusters, (uint32_t)lo_off) - starts;
        const int64_t c1 = t + 1 == n_threads ? n_clusters : std::lower_bound(starts, starts + n_clusters, (uint32_t)hi_off) - starts;
        std::string& out = piece[(size_t)t];
        if (c1 > c0) out.reserve((size_t)(starts[c1] - starts[c0]) * 3 + (size_t)(c1 - c0) * 4 + 16);
        char buf[16];
        for (int64_t c = c0; c < c1; ++c) {
            out += signs[c] > 0 ? "+ " : "- ";
            uint32_t prev = 0;
            for (uint32_t k = starts[c]; k < starts[c + 1]; ++k) {
                uint32_t v = sites[k] - prev;  // members ascend from 0
                char* e = buf + sizeof buf;
                *--e = ' ';
                do { *--e = (char)('0' + v % 10); v /= 10; } while (v);
                out.append(e, (size_t)(buf + sizeof buf - e));
                prev = sites[k];
            }
            out += '\n';
        }
    };
    if (n_threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    size_t total = 1;
    for (const auto& x : piece) total += x.size();
    std::string out;
    out.reserve(total);
    for (const auto& x : piece) out += x;
    out += '\n';  // `file << lines << endl` (main.cpp:255)
    return out;
}

struct UnionFind {
    std::vector<int32_t> parent;
    explicit UnionFind(size_t n) : parent(n) { std::iota(parent.begin(), parent.end(), 0); }
    int32_t find(int32_t a) {
        while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
        return a;
    }
    void unite(int32_t a, int32_t b) {
        a = find(a); b = find(b);
        if (a == b) return;
        if (a < b) parent[b] = a; else parent[a] = b;  // root = smallest site of the cluster
    }
};

// bond_open(site, dir) with dir 0 = +x, 1 = +y decides whether a same-sign bond is kept (any callable).
template <class BondOpen>
inline std::string clusters_to_gap(const int8_t* spins, int L, bool periodic, const BondOpen& bond_open) {
    const int64_t N = (int64_t)L * L;
    UnionFind uf((size_t)N);
    for (int y = 0; y < L; ++y)
        for (int x = 0; x < L; ++x) {
            const int64_t i = x + (int64_t)y * L;
            const int8_t s = spins[i];
            if (s == 0) continue;
            int xr = x + 1, yd = y + 1;
            const bool has_r = xr < L || periodic, has_d = yd < L || periodic;
            if (xr == L) xr = 0;
            if (yd == L) yd = 0;
            if (has_r && L > 1) {
                const int64_t j = xr + (int64_t)y * L;
                if (j != i && spins[j] == s && bond_open(i, 0)) uf.unite((int32_t)i, (int32_t)j);
            }
            if (has_d && L > 1) {
                const int64_t j = x + (int64_t)yd * L;
                if (j != i && spins[j] == s && bond_open(i, 1)) uf.unite((int32_t)i, (int32_t)j);
            }
        }
    // roots are the smallest site of each cluster -> clusters ordered by root, members ascending
    std::vector<int32_t> root((size_t)N, -1), count((size_t)N, 0), start((size_t)N, 0);
    for (int64_t i = 0; i < N; ++i)
        if (spins[i] != 0) ++count[(size_t)(root[(size_t)i] = uf.find((int32_t)i))];
    int64_t acc = 0;
    for (int64_t i = 0; i < N; ++i) { start[(size_t)i] = (int32_t)acc; acc += count[(size_t)i]; }
    std::vector<int32_t> members((size_t)acc), fill(start);
    for (int64_t i = 0; i < N; ++i)
        if (root[(size_t)i] >= 0) members[(size_t)fill[(size_t)root[(size_t)i]]++] = (int32_t)i;
    std::string out;
    out.reserve((size_t)acc * 3 + 16);
    char buf[16];
    for (int64_t r = 0; r < N; ++r) {
        if (count[(size_t)r] == 0) continue;
        out += spins[r] > 0 ? "+ " : "- ";
        int32_t prev = 0;
        for (int32_t k = 0; k < count[(size_t)r]; ++k) {
            const int32_t p = members[(size_t)(start[(size_t)r] + k)];
            // decimal digits of the gap (>= 0: members ascend from 0) followed by one space, written backwards
            uint32_t v = (uint32_t)(p - prev);
            char* e = buf + sizeof buf;
            *--e = ' ';
            do { *--e = (char)('0' + v % 10); v /= 10; } while (v);
            out.append(e, (size_t)(buf + sizeof buf - e));
            prev = p;
        }
        out += '\n';
    }
    out += '\n';  // `file << lines << endl` (main.cpp:255)
    return out;
}

inline std::string spin_clusters_to_gap(const int8_t* spins, int L, bool periodic) {
    return clusters_to_gap(spins, L, periodic, [](int64_t, int) { return true; });
}

// Gap file -> lattice.  Like get_lattice_from_burn (main.cpp:281-299) only listed sites are
// assigned; unlike it the lattice is zero-initialised first (the reference leaves unlisted sites
// at whatever the stack held, SURVEY section 2 "known defects").  Returns false on a bad file.
inline bool gap_to_lattice(const std::string& text, int L, int8_t* spins) {
    const int64_t N = (int64_t)L * L;
    for (int64_t i = 0; i < N; ++i) spins[i] = 0;
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        if (line[0] != '+' && line[0] != '-') continue;
        const int8_t v = line[0] == '+' ? 1 : -1;
        std::istringstream ls(line.substr(1));
        int64_t pos = 0, d;
        while (ls >> d) {
            pos += d;
            if (pos < 0 || pos >= N) return false;
            spins[pos] = v;
        }
    }
    return true;
}

inline bool read_file(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

inline bool write_file(const std::string& path, const std::string& data) {
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    f << data;
    return (bool)f;
}

}  // namespace bcgap
