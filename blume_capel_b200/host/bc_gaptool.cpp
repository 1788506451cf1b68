// bc_gaptool -- converts between raw int8 lattices and the reference's gap-format cluster files.
//   bc_gaptool write L [periodic|open] lattice.bin out.txt     (spin clusters, bond probability 1)
//   bc_gaptool read  L in.txt lattice.bin
// Used by the CPU tests to check byte-compatibility with files written by the unmodified reference.
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "gap_format.hpp"

int main(int argc, char** argv) {
    if (argc < 5) {
        std::cerr << "usage: bc_gaptool write L periodic|open lattice.bin out.txt | read L in.txt lattice.bin\n";
        return 2;
    }
    const std::string mode = argv[1];
    const int L = std::atoi(argv[2]);
    if (L < 1) return 2;
    std::vector<int8_t> spins((size_t)L * L);
    if (mode == "write" && argc == 6) {
        std::string raw;
        if (!bcgap::read_file(argv[4], raw) || raw.size() != spins.size()) { std::cerr << "bad lattice file\n"; return 1; }
        std::memcpy(spins.data(), raw.data(), raw.size());
        return bcgap::write_file(argv[5], bcgap::spin_clusters_to_gap(spins.data(), L, std::string(argv[3]) == "periodic")) ? 0 : 1;
    }
    if (mode == "read" && argc == 5) {
        std::string text;
        if (!bcgap::read_file(argv[3], text) || !bcgap::gap_to_lattice(text, L, spins.data())) { std::cerr << "bad gap file\n"; return 1; }
        return bcgap::write_file(argv[4], std::string((const char*)spins.data(), spins.size())) ? 0 : 1;
    }
    return 2;
}
