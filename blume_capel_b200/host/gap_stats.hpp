// gap_stats.hpp -- the reference's gap-size histogram (corner_contribution/gap_size_statistics.cpp) computed
// from cluster text in memory instead of from files on disk.  Restates, quirks included, what the
// reference's tool does with one cluster file (gap_size_statistics.cpp:84-144):
//   * a cluster line is its delta-coded site list; the walk starts at posn = first_site / L (the ROW of the
//     first site, gap_size_statistics.cpp:104 -- SURVEY section 2 "known defects") and never records the first site;
//   * posn += delta; when posn >= L the collected "row" is closed (kept only if it has > 1 entries) and
//     posn %= L (:108-118);
//   * per row: gaps between consecutive entries, each folded to min(gap, L - gap), plus the wrap-around gap
//     last - first; a row of exactly two entries contributes one gap (:58-81).
// Byte-compatibility with the unmodified tool is what tests/test_magnet.py checks.
#pragma once
#include <algorithm>
#include <sstream>
#include <string>
#include <vector>

namespace bcgapstats {

inline void row_gaps(std::vector<long long>& hist, const std::vector<int>& row, int L) {
    auto fold = [&](int gap) { return std::min(gap, L - gap); };
    if (row.size() == 2) {
        ++hist[(size_t)fold(row[1] - row[0]) - 1];
        return;
    }
    for (size_t i = 0; i + 1 < row.size(); ++i) ++hist[(size_t)fold(row[i + 1] - row[i]) - 1];
    ++hist[(size_t)fold(row.back() - row.front()) - 1];
}

// hist has L/2 bins (bin g-1 counts gaps of size g); gap_text is the content of one cluster file
inline void accumulate(std::vector<long long>& hist, const std::string& gap_text, int L) {
    std::istringstream in(gap_text);
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tok;
        std::vector<int> cluster;
        bool first = true;
        while (std::getline(ls, tok, ' ')) {
            if (first) { first = false; continue; }  // the sign token
            if (!tok.empty()) cluster.push_back(std::stoi(tok));
        }
        if (cluster.empty()) continue;
        int posn = cluster[0] / L;
        std::vector<int> row;
        for (size_t i = 1; i < cluster.size(); ++i) {
            posn += cluster[i];
            if (posn >= L) {
                if (row.size() > 1) row_gaps(hist, row, L);
                row.clear();
                posn %= L;
            }
            row.push_back(posn);
        }
        if (row.size() > 1) row_gaps(hist, row, L);
    }
}

// the per-run output file: "<num_samples>\n<h1> <h2> ... \n" (gap_size_statistics.cpp:157-168)
inline std::string run_file(const std::vector<long long>& hist, long long num_samples) {
    std::string out = std::to_string(num_samples) + "\n";
    for (long long n : hist) out += std::to_string(n) + " ";
    out += "\n";
    return out;
}

}  // namespace bcgapstats
