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
};

void build(Lattice& lt, int L, int open, int ghost, const int8_t* spins) {
    ClGeom& g = lt.g;
    g.L = L; g.open = open;
    g.Wc = (((L + 1) / 2) + 15) / 16 * 16;
    g.rows_arr = L + (ghost ? 2 : 0); g.row_first = ghost ? 1 : 0;
    g.tiles_x = g.tiles_y = (L + CL_T - 1) / CL_T;
    lt.c0.assign((size_t)g.rows_arr * g.Wc, 1); lt.c1.assign((size_t)g.rows_arr * g.Wc, 1);
    for (int y = 0; y < L; ++y)
        for (int x = 0; x < L; ++x)
            (((x + y) & 1) ? lt.c1 : lt.c0)[(size_t)(g.row_first + y) * g.Wc + (x >> 1)] = (uint8_t)(spins[(size_t)y * L + x] + 1);
    lt.n_tiles = (uint32_t)(g.tiles_x * g.tiles_y);
    lt.par.assign((size_t)lt.n_tiles * CL_SITES, 0xEEEE);
    lt.hb.assign((size_t)lt.n_tiles * CL_T, 0); lt.pm.assign((size_t)lt.n_tiles * 2 * CL_T, 0); lt.vbb.assign(lt.n_tiles, 0);
    lt.proot.assign((size_t)lt.n_tiles * 256, CL_NONE);
    lt.gparent.assign((size_t)L * L, 0xDDDDDDDDu);
    lt.out = ClTileOut{lt.par.data(), lt.hb.data(), lt.pm.data(), lt.vbb.data(), lt.proot.data(), lt.gparent.data()};
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
        for (int l = 0; l < CL_LANES; ++l) cl_phase_finish(o[l], lt.g, t, sm, ln[o[l]], lt.out.gparent);
        std::memcpy(lt.out.par + (size_t)tile * CL_SITES, sm.par, sizeof(sm.par));
        for (int l = 0; l < CL_LANES; ++l) {
            lt.out.hb[(size_t)tile * CL_T + l] = ln[l].HB;
            lt.out.pm[(size_t)tile * 2 * CL_T + l] = ln[l].P;
            lt.out.pm[(size_t)tile * 2 * CL_T + CL_T + l] = ln[l].M;
            if (l == t.h - 1) lt.out.vbb[tile] = ln[l].VB;
        }
    }
    // cl_border_kernel
    for (uint32_t tile = 0; tile < lt.n_tiles; ++tile)
        for (int th = 127; th >= 0; --th) {
            const ClTileId t = cl_tile_id(lt.g, tile);
            const int j = th & 63, down = th >> 6;
            uint32_t a, b;
            ClTileId n = t;
            if (!down) {
                if (j >= t.h || !((lt.out.hb[(size_t)tile * CL_T + j] >> (t.w - 1)) & 1ull)) continue;
                n = cl_tile_id(lt.g, tile - (uint32_t)t.tx + (uint32_t)((t.tx + 1) % lt.g.tiles_x));
                a = lt.out.proot[(size_t)tile * 256 + 192 + j];
                b = lt.out.proot[(size_t)n.tile * 256 + 128 + j];
            } else {
                if (j >= t.w || !((lt.out.vbb[tile] >> j) & 1ull)) continue;
                n = cl_tile_id(lt.g, tile - (uint32_t)(t.ty * lt.g.tiles_x) + (uint32_t)(((t.ty + 1) % lt.g.tiles_y) * lt.g.tiles_x));
                a = lt.out.proot[(size_t)tile * 256 + 64 + j];
                b = lt.out.proot[(size_t)n.tile * 256 + j];
            }
            cl_gunion(lt.out.gparent, cl_global_site(lt.g, t, a), cl_global_site(lt.g, n, b));
        }
}

void reload(Lattice& lt, const ClTileId& t, ClLabelSmem& sm, ClLane (&ln)[CL_LANES]) {
    std::memset(&sm, 0xCD, sizeof(sm));
    std::memcpy(sm.par, lt.out.par + (size_t)t.tile * CL_SITES, sizeof(sm.par));
    std::memset(sm.flag, 0, sizeof(sm.flag));
    for (int l = 0; l < CL_LANES; ++l) cl_phase_reload(l, t, lt.out.pm + (size_t)t.tile * 2 * CL_T, lt.out.hb + (size_t)t.tile * CL_T, ln[l]);
}
}  // namespace

extern "C" {

// one Swendsen-Wang move of an L x L lattice (spins in place); ghost = 1 uses the array geometry of a one-rank slab
int emu_cluster_update(int L, int open, int ghost, int8_t* spins, uint32_t thr, uint64_t seed, uint32_t replica, int64_t cstep) {
    Lattice lt;
    build(lt, L, open, ghost, spins);
    // the kernels index thr by the tile's replica and xor it into the key; emulate replica `replica` as replica 0 of a
    // key that already carries it
    const ClRng rng{&thr, (uint32_t)seed, (uint32_t)(seed >> 32) ^ replica, (uint32_t)((uint64_t)cstep & 0xFFFFFFFFu),
                    (uint32_t)((uint64_t)cstep >> 32)};
    label(lt, rng);
    for (uint32_t tile = 0; tile < lt.n_tiles; ++tile) {
        ClLabelSmem sm;
        ClLane ln[CL_LANES];
        const ClTileId t = cl_tile_id(lt.g, tile);
        reload(lt, t, sm, ln);
        for (int l = CL_LANES - 1; l >= 0; --l) cl_phase_decide(l, lt.g, t, rng, sm, ln[l], lt.out.gparent);
        for (int l = 0; l < CL_LANES; ++l)
            if (l < t.h) cl_apply_flips(lt.g, lt.c0.data(), lt.c1.data(), t.rep, t.x0, t.y0 + l, ln[l].P, ln[l].M, cl_flip_mask(l, sm, ln[l]));
    }
    for (int y = 0; y < L; ++y)
        for (int x = 0; x < L; ++x)
            spins[(size_t)y * L + x] = (int8_t)((int)(((x + y) & 1) ? lt.c1 : lt.c0)[(size_t)(lt.g.row_first + y) * lt.g.Wc + (x >> 1)] - 1);
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
    for (uint32_t tile = 0; tile < lt.n_tiles; ++tile) {
        ClLabelSmem sm;
        ClLane ln[CL_LANES];
        const ClTileId t = cl_tile_id(lt.g, tile);
        reload(lt, t, sm, ln);
        for (int l = 0; l < CL_LANES; ++l)
            if (l < t.h) cl_row_labels(l, lt.g, t, sm, ln[l], lt.out.gparent, labels + (size_t)(t.y0 + l) * L + t.x0);
    }
    return 0;
}

}  // extern "C"
