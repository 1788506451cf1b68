// cluster_emulation.cpp -- TEST INFRASTRUCTURE: runs the phases of the cluster kernels (bc_cluster_tile.cuh, every
// function of which is host + device) lane by lane on the CPU, in the order and with the synchronisation points of
// cl_label_kernel / cl_border_kernel / cl_flip_kernel / cl_labels_kernel (bc_cluster_kernels.cuh), so that the tile
// algorithm can be compared with the oracle in the development container, which has no GPU.  A CTA's __syncthreads()
// is "every thread has finished the phase" (threads of a phase run in a seeded random order); atomics are plain read-modify-writes.  Built by tests/test_cluster_cpu.py:
//     g++ -O2 -std=c++17 -shared -fPIC -o tests/cpu/libcluster_emulation.so tests/cpu/cluster_emulation.cpp
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../blume_capel_b200/csrc/bc_cluster_tile.cuh"

using namespace bc;

namespace {
struct Lattice {
    ClGeom g;
    std::vector<uint8_t> c0, c1;  // compact colour arrays of ONE replica (with or without ghost rows)
    std::vector<uint16_t> par;
    std::vector<uint64_t> hb, pm, vbb;
    std::vector<uint16_t> proot;
    std::vector<uint32_t> gparent;
    uint32_t n_tiles;
    ClTileOut out;
    ClGlobal G;
};

// rank `rank` of `n_ranks` row slabs of the L x L lattice `spins` (n_ranks = 1: the whole lattice); a slab of a multi-rank
// decomposition has ghost rows that hold the neighbours' boundary rows (periodic wrap; code 1 at an open edge)
void build(Lattice& lt, int L, int open, int ghost, const int8_t* spins, int n_ranks = 1, int rank = 0) {
    ClGeom& g = lt.g;
    g.L = L; g.open = open;
    g.Wc = (((L + 1) / 2) + 15) / 16 * 16;
    g.Ly = L / n_ranks; g.y_glob0 = rank * g.Ly; g.multi = n_ranks > 1;
    if (g.multi) ghost = 1;
    g.rows_arr = g.Ly + (ghost ? 2 : 0); g.row_first = ghost ? 1 : 0;
    g.tiles_x = (L + CL_T - 1) / CL_T; g.tiles_y = (g.Ly + CL_T - 1) / CL_T;
    lt.c0.assign((size_t)g.rows_arr * g.Wc, 1); lt.c1.assign((size_t)g.rows_arr * g.Wc, 1);
    for (int ya = 0; ya < g.rows_arr; ++ya) {
        int y = g.y_glob0 + ya - g.row_first;  // global row of array row ya
        if (y < 0 || y >= L) {
            if (open || !g.multi) continue;
            y = (y + L) % L;
        }
        if (!g.multi && (ya < g.row_first || ya >= g.row_first + g.Ly)) continue;  // one-rank slab: ghost rows never read
        for (int x = 0; x < L; ++x)
            (((x + y) & 1) ? lt.c1 : lt.c0)[(size_t)ya * g.Wc + (x >> 1)] = (uint8_t)(spins[(size_t)y * L + x] + 1);
    }
    lt.n_tiles = (uint32_t)(g.tiles_x * g.tiles_y);
    lt.par.assign((size_t)lt.n_tiles * CL_SITES, 0xEEEE);
    lt.hb.assign((size_t)lt.n_tiles * CL_T, 0); lt.pm.assign((size_t)lt.n_tiles * 2 * CL_T, 0); lt.vbb.assign(lt.n_tiles, 0);
    lt.proot.assign((size_t)lt.n_tiles * 256, CL_NONE);
    lt.gparent.assign((size_t)g.Ly * L, 0xDDDDDDDDu);
    lt.out = ClTileOut{lt.par.data(), lt.hb.data(), lt.pm.data(), lt.vbb.data(), lt.proot.data(), lt.gparent.data()};
    lt.G = ClGlobal{};
    lt.G.base[0] = lt.gparent.data();
    lt.G.sites_per_rank = 0;
}

// lanes of a phase in a seeded random order: a phase must not depend on the order its threads run in
struct Order {
    int idx[CL_LANES];
    uint64_t state;
    explicit Order(uint64_t seed) : state(seed * 0x9E3779B97F4A7C15ull + 1) { for (int i = 0; i < CL_LANES; ++i) idx[i] = i; }
    const int* next() {
        for (int i = CL_LANES - 1; i > 0; --i) {
            state = state * 6364136223846793005ull + 1442695040888963407ull;
            const int j = (int)((state >> 33) % (uint64_t)(i + 1));
            const int tmp = idx[i]; idx[i] = idx[j]; idx[j] = tmp;
        }
        return idx;
    }
};

void label(Lattice& lt, const ClRng& rng) {
    for (uint32_t tile = 0; tile < lt.n_tiles; ++tile) {
        ClTileSmem sm;
        std::memset(&sm, 0xAB, sizeof(sm));  // shared memory is not zero-initialised
        ClLane ln[CL_LANES];
        Order ord(tile);
        const ClTileId t = cl_tile_id(lt.g, tile);
        const int* o = ord.next();
        for (int l = 0; l < CL_LANES; ++l) cl_phase_load(o[l], lt.g, t, lt.c0.data(), lt.c1.data(), sm, ln[o[l]]);
        o = ord.next();
        for (int l = 0; l < CL_LANES; ++l) cl_phase_bonds(o[l], lt.g, t, lt.c0.data(), lt.c1.data(), rng, sm, ln[o[l]]);
        o = ord.next();
        for (int l = 0; l < CL_LANES; ++l) cl_phase_unions(o[l], t, sm, ln[o[l]]);
        for (bool any = true; any;) {  // while (__syncthreads_or(cl_phase_jump(...)))
            any = false;
            o = ord.next();
            for (int l = 0; l < CL_LANES; ++l) any |= cl_phase_jump(o[l], sm, ln[o[l]]);
        }
        o = ord.next();
        for (int l = 0; l < CL_LANES; ++l) cl_phase_perimeter(o[l], t, sm, ln[o[l]], lt.out.proot + (size_t)tile * 256);
        o = ord.next();
        for (int l = 0; l < CL_LANES; ++l) cl_phase_finish(o[l], lt.g, t, sm, ln[o[l]], lt.G);
        std::memcpy(lt.out.par + (size_t)tile * CL_SITES, sm.par, sizeof(sm.par));
        for (int l = 0; l < CL_LANES; ++l) {
            lt.out.hb[(size_t)tile * CL_T + l] = ln[l].HB;
            lt.out.pm[(size_t)tile * 2 * CL_T + l] = ln[l].P;
            lt.out.pm[(size_t)tile * 2 * CL_T + CL_T + l] = ln[l].M;
            if (l == t.h - 1) lt.out.vbb[tile] = ln[l].VB;
        }
    }
}

// cl_border_kernel; `below`: the rank below of a multi-rank decomposition (its perimeter roots and first global row)
void border(Lattice& lt, const Lattice* below) {
    for (uint32_t tile = 0; tile < lt.n_tiles; ++tile)
        for (int th = 127; th >= 0; --th) {
            const ClTileId t = cl_tile_id(lt.g, tile);
            const int j = th & 63, down = th >> 6;
            uint32_t a, b;
            ClTileId n = t;
            ClGeom gn = lt.g;
            if (!down) {
                if (j >= t.h || !((lt.out.hb[(size_t)tile * CL_T + j] >> (t.w - 1)) & 1ull)) continue;
                n = cl_tile_id(lt.g, tile - (uint32_t)t.tx + (uint32_t)((t.tx + 1) % lt.g.tiles_x));
                a = lt.out.proot[(size_t)tile * 256 + 192 + j];
                b = lt.out.proot[(size_t)n.tile * 256 + 128 + j];
            } else {
                if (j >= t.w || !((lt.out.vbb[tile] >> j) & 1ull)) continue;
                const bool last = t.ty + 1 == lt.g.tiles_y;
                n = cl_tile_id(lt.g, tile - (uint32_t)(t.ty * lt.g.tiles_x) + (uint32_t)(((t.ty + 1) % lt.g.tiles_y) * lt.g.tiles_x));
                a = lt.out.proot[(size_t)tile * 256 + 64 + j];
                if (last && lt.g.multi) {
                    gn.y_glob0 = below->g.y_glob0;
                    b = below->out.proot[(size_t)n.tile * 256 + j];
                } else {
                    b = lt.out.proot[(size_t)n.tile * 256 + j];
                }
            }
            cl_gunion(lt.G, cl_global_site(lt.g, t, a), cl_global_site(gn, n, b));
        }
}

void reload(Lattice& lt, const ClTileId& t, ClLabelSmem& sm, ClLane (&ln)[CL_LANES]) {
    std::memset(&sm, 0xCD, sizeof(sm));
    std::memcpy(sm.par, lt.out.par + (size_t)t.tile * CL_SITES, sizeof(sm.par));
    std::memset(sm.flag, 0, sizeof(sm.flag));
    for (int l = 0; l < CL_LANES; ++l) cl_phase_reload(l, t, lt.out.pm + (size_t)t.tile * 2 * CL_T, lt.out.hb + (size_t)t.tile * CL_T, ln[l]);
}
}  // namespace

void flip(Lattice& lt, const ClRng& rng) {
    for (uint32_t tile = 0; tile < lt.n_tiles; ++tile) {
        ClLabelSmem sm;
        ClLane ln[CL_LANES];
        const ClTileId t = cl_tile_id(lt.g, tile);
        reload(lt, t, sm, ln);
        for (int l = CL_LANES - 1; l >= 0; --l) cl_phase_decide(l, lt.g, t, rng, sm, ln[l], lt.G);
        for (int l = 0; l < CL_LANES; ++l)
            if (l < t.h) cl_apply_flips(lt.g, lt.c0.data(), lt.c1.data(), t.rep, t.x0, t.y0 + l, ln[l].P, ln[l].M, cl_flip_mask(l, sm, ln[l]));
    }
}

extern "C" {

// one Swendsen-Wang move of an L x L lattice (spins in place); ghost = 1 uses the array geometry of a one-rank slab;
// n_ranks > 1: the lattice as n_ranks row slabs, every rank with its own arrays and its own piece of the global
// union-find (the other ranks' pieces reached through the ClGlobal base pointers, as over peer memory), phases in the
// order the engine issues them: all ranks label | barrier | all ranks merge borders | barrier | all ranks flip
int emu_cluster_update(int L, int open, int ghost, int n_ranks, int8_t* spins, uint32_t thr, uint64_t seed, uint32_t replica,
                       int64_t cstep) {
    if (n_ranks < 1 || n_ranks > CL_MAX_RANKS || L % n_ranks) return -1;
    std::vector<Lattice> ranks((size_t)n_ranks);
    for (int k = 0; k < n_ranks; ++k) build(ranks[(size_t)k], L, open, ghost, spins, n_ranks, k);
    if (n_ranks > 1)
        for (Lattice& lt : ranks) {
            for (int k = 0; k < n_ranks; ++k) lt.G.base[k] = ranks[(size_t)k].gparent.data();
            lt.G.sites_per_rank = (uint32_t)(L / n_ranks) * (uint32_t)L;
        }
    // the kernels index thr by the tile's replica and xor it into the key; emulate replica `replica` as replica 0 of a
    // key that already carries it
    const ClRng rng{&thr, (uint32_t)seed, (uint32_t)(seed >> 32) ^ replica, (uint32_t)((uint64_t)cstep & 0xFFFFFFFFu),
                    (uint32_t)((uint64_t)cstep >> 32)};
    for (Lattice& lt : ranks) label(lt, rng);
    for (int k = n_ranks - 1; k >= 0; --k) border(ranks[(size_t)k], &ranks[(size_t)((k + 1) % n_ranks)]);
    for (Lattice& lt : ranks) flip(lt, rng);
    for (const Lattice& lt : ranks)
        for (int y = 0; y < lt.g.Ly; ++y)
            for (int x = 0; x < L; ++x)
                spins[(size_t)(lt.g.y_glob0 + y) * L + x] =
                    (int8_t)((int)(((x + y) & 1) ? lt.c1 : lt.c0)[(size_t)(lt.g.row_first + y) * lt.g.Wc + (x >> 1)] - 1);
    return 0;
}

// labels[i] = smallest site of the cluster of site i (0xFFFFFFFF for a zero spin); always = 1: every candidate bond open
int emu_cluster_labels(int L, int open, int ghost, const int8_t* spins, int always, uint32_t thr, uint64_t seed, uint32_t replica,
                       int64_t cstep, uint32_t* labels) {
    Lattice lt;
    build(lt, L, open, ghost, spins);
    const ClRng rng{always ? nullptr : &thr, (uint32_t)seed, (uint32_t)(seed >> 32) ^ replica,
                    (uint32_t)((uint64_t)cstep & 0xFFFFFFFFu), (uint32_t)((uint64_t)cstep >> 32)};
    label(lt, rng);
    border(lt, nullptr);
    for (uint32_t tile = 0; tile < lt.n_tiles; ++tile) {
        ClLabelSmem sm;
        ClLane ln[CL_LANES];
        const ClTileId t = cl_tile_id(lt.g, tile);
        reload(lt, t, sm, ln);
        for (int l = 0; l < CL_LANES; ++l)
            if (l < t.h) cl_row_labels(l, lt.g, t, sm, ln[l], lt.G, labels + (size_t)(t.y0 + l) * L + t.x0);
    }
    return 0;
}

}  // extern "C"
