// bc_cluster_tile.cuh -- bit-parallel cluster labelling of one 64 x 64 tile by 64 threads.
// Part of the sm_100a Blume-Capel engine; included through bc_cluster_kernels.cuh.
//
// What the reference does one cluster at a time with a BFS queue (wolff(), /root/reference/main.cpp:113-154, bond
// probability 1 - exp(-2J/T) at :115, zeros never join :123-130,144) and, for the dumps, with a DFS over the whole
// lattice (form_clusters, main.cpp:179-230), is done here for all clusters at once on 64-bit ROW MASKS:
//   * a tile row is four 64-bit words: plus sites, minus sites, open bonds to the right, open bonds downwards;
//   * thread l of the tile's 64 owns tile row l: it converts its 64 spin bytes (one 32-byte piece of each of the
//     two compact colour arrays) to masks, draws the random bonds of its row 32 at a time (bit-plane
//     uniforms, most significant plane first, until every candidate bond is decided: about two Philox calls per 32
//     bonds, see cl_draw_bonds), and finds the horizontal runs of its row with one AND-NOT;
//   * only RUN STARTS are union-find nodes (16-bit labels in shared memory): a vertical bond costs a union only when
//     the bond to its left does not already join the same two runs; the larger root is always linked under the
//     smaller one, so a cluster's root is its smallest site, independent of the order of the unions;
//   * clusters that touch the tile's perimeter are flagged and merged across tiles by cl_border_kernel in a global
//     union-find over their roots' site indices; everything else never leaves shared memory.
// Between kernels a tile is 8 KB of labels (raw copy of the shared-memory array, 2 B / site), 512 B of bond masks
// and 512 B of perimeter roots -- against 1 B + 4 B + 1 B per site written and 4 B chains re-read by round 1's
// one-thread-per-two-sites kernels (0.24 ms per L = 4096 update; profiles/r01_ncu_launches_cluster_v2.csv).
//
// Every function here is host + device: the phases of the kernels are plain functions of (lane, tile state), the
// kernels call them with __syncthreads() in between (one CTA of 64 threads per tile), and tests/cpu/cluster_emulation.cpp runs the same phases lane by
// lane on the CPU against the oracle (there is no GPU in the development container).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BC_HD __host__ __device__ __forceinline__
#define BC_UNROLL _Pragma("unroll")
#else
#define BC_UNROLL
#define BC_HD inline
struct uint4 { uint32_t x, y, z, w; };  // CPU emulation build (tests/cpu/cluster_emulation.cpp)
#endif

namespace bc {

constexpr int CL_T = 64;                 // tile edge (sites)
constexpr int CL_SITES = CL_T * CL_T;
constexpr uint16_t CL_FLAG = 0x8000u;    // label entry: the root's cluster touches the tile perimeter
constexpr uint16_t CL_IDX = 0x0FFFu;     // label entry: tile-local site index ly * 64 + lx of the root
constexpr uint16_t CL_NONE = 0xFFFFu;    // perimeter entry: empty site
constexpr uint32_t CL_STREAM_BOND = 3u, CL_STREAM_FLIP = 4u;

// ---- small portable helpers -------------------------------------------------------------------------------------
BC_HD int cl_ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
BC_HD int cl_clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}
// position of the run start at or before x: the highest set bit of rs among bits 0..x (there always is one for an
// occupied x, because an occupied site without a bond from its left neighbour starts a run)
BC_HD int cl_run_start(uint64_t rs, int x) { return 63 - cl_clz64(rs & (~0ull >> (63 - x))); }

BC_HD uint16_t cl_cas16(uint16_t* p, uint16_t cmp, uint16_t val) {
#if defined(__CUDA_ARCH__)
    return atomicCAS(reinterpret_cast<unsigned short*>(p), (unsigned short)cmp, (unsigned short)val);
#else
    const uint16_t old = *p;
    if (old == cmp) *p = val;
    return old;
#endif
}
BC_HD void cl_or32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
BC_HD uint32_t cl_cas32(uint32_t* p, uint32_t cmp, uint32_t val) {
#if defined(__CUDA_ARCH__)
    return atomicCAS_system(p, cmp, val);
#else
    const uint32_t old = *p;
    if (old == cmp) *p = val;
    return old;
#endif
}

BC_HD void cl_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
    BC_UNROLL
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ (k0 + (uint32_t)r * 0x9E3779B9u);
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ (k1 + (uint32_t)r * 0xBB67AE85u);
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// bits 0..31 of v to the even bit positions of a 64-bit word, and back
BC_HD uint64_t cl_spread(uint32_t v) {
    uint64_t s = v;
    s = (s | (s << 16)) & 0x0000FFFF0000FFFFull;
    s = (s | (s << 8)) & 0x00FF00FF00FF00FFull;
    s = (s | (s << 4)) & 0x0F0F0F0F0F0F0F0Full;
    s = (s | (s << 2)) & 0x3333333333333333ull;
    s = (s | (s << 1)) & 0x5555555555555555ull;
    return s;
}
BC_HD uint32_t cl_squeeze(uint64_t s) {
    s &= 0x5555555555555555ull;
    s = (s | (s >> 1)) & 0x3333333333333333ull;
    s = (s | (s >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    s = (s | (s >> 4)) & 0x00FF00FF00FF00FFull;
    s = (s | (s >> 8)) & 0x0000FFFF0000FFFFull;
    s = (s | (s >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)s;
}
// bit 0 of each of the four bytes of w as a nibble (no carries: the four products land on distinct bits)
BC_HD uint32_t cl_nibble(uint32_t w) { return ((w & 0x01010101u) * 0x00204081u >> 21) & 0xFu; }
// and back: bits 0..3 of n to bit 0 of the four bytes
BC_HD uint32_t cl_bytes(uint32_t n) { return ((n & 0xFu) * 0x00204081u) & 0x01010101u; }

// ---- geometry ---------------------------------------------------------------------------------------------------
// Colour arrays as the sweep kernels keep them: code = spin + 1, [replica][array row][Wc], site (x, y) lives in array
// (x + y) & 1 at byte x >> 1 of array row row_first + y.  A one-rank slab handle has ghost rows (row_first = 1);
// the labelling wraps by index and never reads them.
struct ClGeom {
    int L, Wc, rows_arr, row_first, open;
    int tiles_x, tiles_y;
    int Ly;        // rows this handle owns (L, or L / n_ranks for a slab of a multi-rank decomposition)
    int y_glob0;   // global row of its first row: bond draws and cluster names use GLOBAL coordinates
    int multi;     // slab of a multi-rank decomposition: the row below the last owned row is the bottom ghost row
};
struct ClTileId {
    uint32_t rep;
    int tx, ty, x0, y0, w, h;  // w, h: sites of the tile that exist (the last tiles of a lattice may be clipped)
    uint32_t tile;             // index of the tile among all tiles of all replicas
};
BC_HD ClTileId cl_tile_id(const ClGeom& g, uint32_t tile) {
    ClTileId t;
    const uint32_t per = (uint32_t)g.tiles_x * (uint32_t)g.tiles_y;
    t.tile = tile;
    t.rep = tile / per;
    const uint32_t in = tile - t.rep * per;
    t.ty = (int)(in / (uint32_t)g.tiles_x);
    t.tx = (int)(in - (uint32_t)t.ty * (uint32_t)g.tiles_x);
    t.x0 = t.tx * CL_T; t.y0 = t.ty * CL_T;
    t.w = g.L - t.x0 < CL_T ? g.L - t.x0 : CL_T;
    t.h = g.Ly - t.y0 < CL_T ? g.Ly - t.y0 : CL_T;
    return t;
}
BC_HD uint32_t cl_global_site(const ClGeom& g, const ClTileId& t, uint32_t local) {
    return (uint32_t)(g.y_glob0 + t.y0 + (int)(local >> 6)) * (uint32_t)g.L + (uint32_t)(t.x0 + (int)(local & 63u));
}
BC_HD size_t cl_row_offset(const ClGeom& g, uint32_t rep, int gy) {
    return ((size_t)rep * (size_t)g.rows_arr + (size_t)(g.row_first + gy)) * (size_t)g.Wc;
}
BC_HD uint32_t cl_site_code(const ClGeom& g, const uint8_t* c0, const uint8_t* c1, uint32_t rep, int gx, int gy) {
    return (((gx + gy) & 1) ? c1 : c0)[cl_row_offset(g, rep, gy) + (size_t)(gx >> 1)];
}

// Masks of the plus (code 2) and minus (code 0) sites of the 64 sites x0 .. x0+63 of lattice row gy.  Even x come
// from array gy & 1, odd x from the other one, 32 bytes each; bytes behind the lattice edge are pad bytes of code 1.
BC_HD void cl_load_row(const ClGeom& g, const uint8_t* c0, const uint8_t* c1, uint32_t rep, int x0, int gy, uint64_t* P,
                       uint64_t* M) {
    const size_t off = cl_row_offset(g, rep, gy) + (size_t)(x0 >> 1);
    const uint8_t* ev = ((gy & 1) ? c1 : c0) + off;
    const uint8_t* od = ((gy & 1) ? c0 : c1) + off;
    const int vecs = (g.Wc - (x0 >> 1)) >= 32 ? 2 : ((g.Wc - (x0 >> 1)) >= 16 ? 1 : 0);  // Wc is a multiple of 16
    uint32_t pe = 0, po = 0, me = 0, mo = 0;
    BC_UNROLL
    for (int v = 0; v < 2; ++v) {
        if (v >= vecs) break;
        const uint4 a = *reinterpret_cast<const uint4*>(ev + 16 * v), b = *reinterpret_cast<const uint4*>(od + 16 * v);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    BC_UNROLL
        for (int k = 0; k < 4; ++k) {
            const int sh = 16 * v + 4 * k;
            pe |= cl_nibble(aw[k] >> 1) << sh;
            po |= cl_nibble(bw[k] >> 1) << sh;
            me |= cl_nibble(~(aw[k] | (aw[k] >> 1))) << sh;
            mo |= cl_nibble(~(bw[k] | (bw[k] >> 1))) << sh;
        }
    }
    *P = cl_spread(pe) | (cl_spread(po) << 1);
    *M = cl_spread(me) | (cl_spread(mo) << 1);
}

// ---- random bonds -----------------------------------------------------------------------------------------------
// Spec (restated one bond at a time by the CPU checker): the 32 sites x = 32 grp .. 32 grp + 31 of lattice row gy share 32
// bit-plane words per direction; plane b (b = 0 most significant) is word b & 3 of
// Philox(gy << 15 | grp << 4 | dir << 3 | b >> 2, c_lo, c_hi, BOND << 1), the uniform of a site is the column of
// its bit through the planes, and the bond is open iff uniform < thr.  Evaluated for all candidates of a 32-site
// group at once: while some candidate is undecided, draw the next plane; where the plane's bit differs from the
// threshold's bit the comparison is decided.  Expected planes for 32 candidates: about seven (two Philox calls).
BC_HD uint64_t cl_draw_bonds(uint64_t cand, uint32_t gy, uint32_t grp0, uint32_t dir, uint32_t thr, uint32_t c_lo,
                             uint32_t c_hi, uint32_t k0, uint32_t k1) {
    uint64_t open = 0;
    for (uint32_t half = 0; half < 2; ++half) {
        uint32_t U = (uint32_t)(cand >> (32 * half)), R = 0;
        for (uint32_t c4 = 0; c4 < 8 && U; ++c4) {
            uint32_t w[4];
            cl_philox((gy << 15) | ((grp0 + half) << 4) | (dir << 3) | c4, c_lo, c_hi, CL_STREAM_BOND << 1, k0, k1, w);
    BC_UNROLL
            for (uint32_t j = 0; j < 4; ++j) {
                if ((thr >> (31u - (4u * c4 + j))) & 1u) { R |= U & ~w[j]; U &= w[j]; }
                else U &= ~w[j];
            }
        }
        open |= (uint64_t)R << (32 * half);
    }
    return open;
}

// flip decision of the cluster whose smallest site is `root` (oracle: cluster_flips); `blk` / `w` cache the last call
BC_HD uint32_t cl_flip_bit(uint32_t root, uint32_t c_lo, uint32_t c_hi, uint32_t k0, uint32_t k1, uint32_t* blk,
                           uint32_t (&w)[4]) {
    if ((root >> 7) != *blk) {
        *blk = root >> 7;
        cl_philox(*blk, c_lo, c_hi, CL_STREAM_FLIP << 1, k0, k1, w);
    }
    return (w[(root >> 5) & 3u] >> (root & 31u)) & 1u;
}

// ---- union-find over run starts (shared memory, 16-bit tile-local labels) ---------------------------------------
// find with path halving: every value ever stored in par[a] is an ancestor of a, so concurrent finds and unions only
// ever see valid (possibly longer) paths; roots are never written by the halving store.
BC_HD uint16_t cl_find(volatile uint16_t* par, uint16_t a) {
    uint16_t p = par[a];
    while (p != a) {
        const uint16_t gp = par[p];
        if (gp != p) par[a] = gp;
        a = p;
        p = gp;
    }
    return a;
}
// link the larger root under the smaller one; retry from the current roots if the larger one stopped being a root.
// Returns the root of the merged tree as far as this thread knows (a hint for its next union).
BC_HD uint16_t cl_union(uint16_t* par, uint16_t a, uint16_t b) {
    uint16_t ra = cl_find(par, a), rb = cl_find(par, b);
    while (ra != rb) {
        if (ra < rb) { const uint16_t t = ra; ra = rb; rb = t; }
        const uint16_t old = cl_cas16(par + ra, ra, rb);
        if (old == ra) break;
        ra = cl_find(par, old);
        rb = cl_find(par, rb);
    }
    return rb;
}
// The same over GLOBAL site indices (clusters that cross tile borders).  The array is indexed by site; when the lattice
// is split into row slabs over several GPUs every rank holds the entries of its own rows and maps the other ranks'
// pieces through peer memory (parents point at smaller sites, i.e. at the same or a lower rank; the compare-and-swap
// and the loads are system-scope and cross NVLink).
constexpr int CL_MAX_RANKS = 8;
struct ClGlobal {
    uint32_t* base[CL_MAX_RANKS];  // base[k]: rank k's piece (this replica's); a single piece: everything in base[0]
    uint32_t sites_per_rank;       // rows per rank * L; 0 = a single piece
};
BC_HD uint32_t* cl_gslot(const ClGlobal& G, uint32_t site) {
    if (G.sites_per_rank == 0u) return G.base[0] + site;
    const uint32_t k = site / G.sites_per_rank;
    return G.base[k] + (site - k * G.sites_per_rank);
}
BC_HD uint32_t cl_gfind(const ClGlobal& G, uint32_t a) {
    volatile uint32_t* sa = cl_gslot(G, a);
    uint32_t p = *sa;
    while (p != a) {
        volatile uint32_t* sp = cl_gslot(G, p);
        const uint32_t gp = *sp;
        if (gp != p) *sa = gp;
        a = p; sa = sp;
        p = gp;
    }
    return a;
}
BC_HD void cl_gunion(const ClGlobal& G, uint32_t a, uint32_t b) {
    uint32_t ra = cl_gfind(G, a), rb = cl_gfind(G, b);
    while (ra != rb) {
        if (ra < rb) { const uint32_t t = ra; ra = rb; rb = t; }
        const uint32_t old = cl_cas32(cl_gslot(G, ra), ra, rb);
        if (old == ra) break;
        ra = cl_gfind(G, old);
        rb = cl_gfind(G, rb);
    }
}

// f(x) for every set bit x of m, ascending: the two 32-bit halves separately (a 64-bit find-first-set is four times
// the instructions of a 32-bit one, and these loops are the kernels' inner loops)
template <class F>
BC_HD void cl_for_bits(uint64_t m, F f) {
    BC_UNROLL
    for (int half = 0; half < 2; ++half)
        for (uint32_t w = (uint32_t)(m >> (32 * half)); w; w &= w - 1) {
#if defined(__CUDA_ARCH__)
            f(32 * half + __ffs((int)w) - 1);
#else
            f(32 * half + __builtin_ctz(w));
#endif
        }
}

// ---- tile state -------------------------------------------------------------------------------------------------
constexpr int CL_LANES = CL_T;    // threads per tile: thread l owns tile row l
struct ClTileSmem {               // cl_label_kernel
    uint16_t par[CL_SITES];       // union-find over run starts; non-run-start entries are never read
    uint64_t plus[CL_T], minus[CL_T];
    uint64_t rs[CL_T];            // run starts per tile row
    uint64_t hb[CL_T];            // open bonds to the right (bit x: x -- x+1; bit w-1: into the next tile)
    uint32_t flag[CL_SITES / 32]; // roots that touch the perimeter
};
struct ClLabelSmem {              // consumers of a labelled tile (cl_flip_kernel, cl_labels_kernel)
    uint16_t par[CL_SITES];       // the tile's labels as cl_label_kernel left them
    uint32_t flag[CL_SITES / 32]; // roots that flip
};
struct ClLane {                   // registers of one thread: its tile row
    uint64_t P, M, HB, VB, RS;
};
// What a tile leaves in global memory for the kernels that follow (all tile-major)
struct ClTileOut {
    uint16_t* par;        // [tile][4096]  raw copy of ClTileSmem::par (run-start entries: root | CL_FLAG)
    uint64_t* hb;         // [tile][64]
    uint64_t* pm;         // [tile][2][64]  plus and minus masks of the rows (the flip kernel rewrites the spin bytes from
                          //                them and never reads the lattice)
    uint64_t* vbb;        // [tile]        open bonds from the tile's last row into the tile below
    uint16_t* proot;      // [tile][256]   root of the perimeter sites: top row, bottom row, left column, right column
    uint32_t* gparent;    // [replica][L*L] global union-find, touched only at the site indices of flagged roots
};
struct ClRng {
    const uint32_t* thr;  // [replica] bond threshold p * 2^32 (nullptr: every candidate bond is open)
    uint32_t k0, k1, c_lo, c_hi;
};

// ---- labelling phases (cl_label_kernel; `lane` = 0 .. 63 = the thread's tile row) -------------------------------
// Phase 1: masks of the thread's row.
BC_HD void cl_phase_load(int lane, const ClGeom& g, const ClTileId& t, const uint8_t* c0, const uint8_t* c1, ClTileSmem& sm,
                         ClLane& ln) {
    for (int i = 0; i < CL_SITES / 32 / CL_LANES; ++i) sm.flag[(CL_SITES / 32 / CL_LANES) * lane + i] = 0u;
    ln.P = ln.M = 0;
    if (lane < t.h) cl_load_row(g, c0, c1, t.rep, t.x0, t.y0 + lane, &ln.P, &ln.M);
    if (t.w < CL_T) { ln.P &= (1ull << t.w) - 1; ln.M &= (1ull << t.w) - 1; }
    sm.plus[lane] = ln.P;
    sm.minus[lane] = ln.M;
}

// Phase 2: candidate bonds (equal non-zero neighbours), random bonds, runs.
BC_HD void cl_phase_bonds(int lane, const ClGeom& g, const ClTileId& t, const uint8_t* c0, const uint8_t* c1, const ClRng& rng,
                          ClTileSmem& sm, ClLane& ln) {
    const int r = lane, gy = t.y0 + r;             // row of the handle's own piece (array row row_first + gy)
    const uint32_t yg = (uint32_t)(g.y_glob0 + gy);  // global row: the bond draws are functions of global coordinates
    ln.HB = ln.VB = ln.RS = 0;
    if (r < t.h) {
        const uint64_t P = ln.P, M = ln.M;
        // right neighbour of the tile's last column: the next tile's first column, or the wrap, or nothing
        uint64_t pr = 0, mr = 0;
        int nx = t.x0 + t.w;
        if (nx == g.L) nx = g.open ? -1 : 0;
        if (nx >= 0) {
            const uint32_t c = cl_site_code(g, c0, c1, t.rep, nx, gy);
            pr = c == 2u; mr = c == 0u;
        }
        uint64_t hc = (P & ((P >> 1) | (pr << (t.w - 1)))) | (M & ((M >> 1) | (mr << (t.w - 1))));
        // row below: inside the tile from shared memory, for the tile's last row from the lattice (or the wrap)
        uint64_t Pd = 0, Md = 0;
        if (r + 1 < t.h) { Pd = sm.plus[r + 1]; Md = sm.minus[r + 1]; }
        else {
            // below the tile: the handle's next row; below its last row the wrap (whole lattice) or the bottom ghost
            // row, array row row_first + Ly (slab of a multi-rank decomposition); nothing below an open lattice
            int ny = gy + 1;
            if (ny == g.Ly) ny = (g.open && yg + 1u == (uint32_t)g.L) ? -1 : (g.multi ? g.Ly : 0);
            if (ny >= 0 && ny != gy) {
                cl_load_row(g, c0, c1, t.rep, t.x0, ny, &Pd, &Md);
                if (t.w < CL_T) { Pd &= (1ull << t.w) - 1; Md &= (1ull << t.w) - 1; }
            }
        }
        uint64_t vc = (P & Pd) | (M & Md);
        if (rng.thr) {
            const uint32_t thr = rng.thr[t.rep];
            hc = cl_draw_bonds(hc, yg, (uint32_t)t.x0 >> 5, 0u, thr, rng.c_lo, rng.c_hi, rng.k0, rng.k1 ^ t.rep);
            vc = cl_draw_bonds(vc, yg, (uint32_t)t.x0 >> 5, 1u, thr, rng.c_lo, rng.c_hi, rng.k0, rng.k1 ^ t.rep);
        }
        ln.HB = hc; ln.VB = vc;
        ln.RS = (P | M) & ~(hc << 1);
        cl_for_bits(ln.RS, [&](int x) { sm.par[r * CL_T + x] = (uint16_t)(r * CL_T + x); });
    }
    sm.rs[r] = ln.RS;
    sm.hb[r] = ln.HB;
}

// Phase 3: one union per pair of runs joined by a vertical bond (skipped when the bond to the left joins the same
// pair).  Consecutive bonds often share a run: the root a union returned is the starting point of the next find.
BC_HD void cl_phase_unions(int lane, const ClTileId& t, ClTileSmem& sm, const ClLane& ln) {
    const int r = lane;
    if (r + 1 >= t.h) return;  // the last row's bonds leave the tile
    const uint64_t vb = ln.VB, rs_t = ln.RS, rs_d = sm.rs[r + 1];
    int last_a = -1, last_b = -1;
    uint16_t root = 0;
    cl_for_bits(vb & ~((vb & ln.HB & sm.hb[r + 1]) << 1), [&](int x) {
        const int a = cl_run_start(rs_t, x), b = cl_run_start(rs_d, x);
        const uint16_t ua = a == last_a ? root : (uint16_t)(r * CL_T + a), ub = b == last_b ? root : (uint16_t)((r + 1) * CL_T + b);
        root = cl_union(sm.par, ua, ub);
        last_a = a; last_b = b;
    });
}

// Phase 4, repeated until no thread reports a change: pointer jumping, every run start's entry moves to its
// grandparent.  The unions leave chains (a root that is linked under a smaller one takes its whole tree one level
// down, and the root of a spanning cluster changes dozens of times); walking them run start by run start kept five
// of 32 lanes busy for a third of the kernel.  Jumping halves every chain per round, all threads in step; any value
// read is an ancestor, so the rounds need no double buffering.  At the fixed point every entry is its ROOT.
BC_HD bool cl_phase_jump(int lane, ClTileSmem& sm, const ClLane& ln) {
    bool changed = false;
    cl_for_bits(ln.RS, [&](int x) {
        const int i = lane * CL_T + x;
        const uint16_t p = sm.par[i], gp = sm.par[p];
        if (gp != p) { sm.par[i] = gp; changed = true; }
    });
    return changed;
}

// Phase 5: roots of the perimeter sites (thread l: column l of the top and bottom row, its own row in the left and
// right column), flagged as "touches the perimeter"
BC_HD void cl_phase_perimeter(int lane, const ClTileId& t, ClTileSmem& sm, const ClLane& ln, uint16_t* proot) {
    BC_UNROLL
    for (int e = 0; e < 4; ++e) {
        // e = 0, 1: column x = lane of the top / bottom row; e = 2, 3: row y = lane of the left / right column
        const int row = e == 0 ? 0 : (e == 1 ? t.h - 1 : lane), col = e < 2 ? lane : (e == 2 ? 0 : t.w - 1);
        const uint64_t occ = e < 2 ? (sm.plus[row] | sm.minus[row]) : (ln.P | ln.M);
        const uint64_t rs = e < 2 ? sm.rs[row] : ln.RS;
        uint16_t root = CL_NONE;
        if (row < t.h && col < t.w && ((occ >> col) & 1ull)) {
            root = sm.par[row * CL_T + cl_run_start(rs, col)];
            cl_or32(&sm.flag[root >> 5], 1u << (root & 31u));
        }
        proot[64 * e + lane] = root;
    }
}

// Phase 6: the root's flag into every run start's entry (written by its owner, read by nobody else in this phase);
// flagged roots become nodes of the global union-find
BC_HD void cl_phase_finish(int lane, const ClGeom& g, const ClTileId& t, ClTileSmem& sm, const ClLane& ln, const ClGlobal& G) {
    cl_for_bits(ln.RS, [&](int x) {
        const uint16_t i = (uint16_t)(lane * CL_T + x);
        const uint16_t root = sm.par[i];
        if ((sm.flag[root >> 5] >> (root & 31u)) & 1u) {
            sm.par[i] = root | CL_FLAG;
            if (root == i) { const uint32_t s = cl_global_site(g, t, i); *cl_gslot(G, s) = s; }
        }
    });
}

// ---- consumers of a labelled tile -------------------------------------------------------------------------------
// Reload what cl_label_kernel left: labels into shared memory (done by the caller), masks and runs of the thread's row.
BC_HD void cl_phase_reload(int lane, const ClTileId& t, const uint64_t* pm_tile, const uint64_t* hb_tile, ClLane& ln) {
    ln.P = ln.M = ln.HB = ln.VB = ln.RS = 0;
    if (lane >= t.h) return;
    ln.P = pm_tile[lane];
    ln.M = pm_tile[CL_T + lane];
    ln.HB = hb_tile[lane];
    ln.RS = (ln.P | ln.M) & ~(ln.HB << 1);
}

// smallest site of the cluster of tile-local root entry `e` (root | flag)
BC_HD uint32_t cl_cluster_site(const ClGeom& g, const ClTileId& t, uint16_t e, const ClGlobal& G) {
    const uint32_t s = cl_global_site(g, t, e & CL_IDX);
    if (!(e & CL_FLAG)) return s;
    const uint32_t root = cl_gfind(G, s);
    uint32_t* slot = cl_gslot(G, s);
    if (*slot != root) *slot = root;  // later lookups of this tile root take one hop
    return root;
}

// Flip phase 1: the thread's roots decide.  One Philox call serves the 128 sites around a root: the call for the
// thread's own row is made up front by all threads together (an unflagged root is a site of the row; inside the loop
// over the run starts, where the threads meet their roots at different iterations, it ran at a seventh of the lanes).
BC_HD void cl_phase_decide(int lane, const ClGeom& g, const ClTileId& t, const ClRng& rng, ClLabelSmem& sm, const ClLane& ln,
                           const ClGlobal& G) {
    uint32_t blk = cl_global_site(g, t, (uint32_t)lane * CL_T) >> 7, w[4];
    cl_philox(blk, rng.c_lo, rng.c_hi, CL_STREAM_FLIP << 1, rng.k0, rng.k1 ^ t.rep, w);
    cl_for_bits(ln.RS, [&](int x) {
        const uint16_t i = (uint16_t)(lane * CL_T + x);
        const uint16_t e = sm.par[i];
        if ((e & CL_IDX) != i) return;
        const uint32_t root = cl_cluster_site(g, t, e, G);
        if (cl_flip_bit(root, rng.c_lo, rng.c_hi, rng.k0, rng.k1 ^ t.rep, &blk, w)) cl_or32(&sm.flag[i >> 5], 1u << (i & 31u));
    });
}

// Flip phase 2: mask of the sites of the thread's row whose cluster flips
BC_HD uint64_t cl_flip_mask(int lane, const ClLabelSmem& sm, const ClLane& ln) {
    uint64_t starts = 0;  // run starts whose cluster flips
    cl_for_bits(ln.RS, [&](int x) {
        const uint16_t root = sm.par[lane * CL_T + x] & CL_IDX;
        starts |= (uint64_t)((sm.flag[root >> 5] >> (root & 31u)) & 1u) << x;
    });
    // A run = its start and the sites up to the next start: fill(x) = starts(x) | (fill(x-1) & ~RS(x)) is the carry
    // chain of an addition with generate = starts and propagate = ~RS: in X + starts with X = starts | ~RS the carry
    // INTO position x + 1 is exactly fill(x).  (Zero sites behind a run inherit its bit; masked by the occupancy.)
    const uint64_t X = starts | ~ln.RS;
    const uint64_t carry = (X + starts) ^ X ^ starts;
    const uint64_t F = (carry >> 1) | ((starts | (~ln.RS & carry)) & (1ull << 63));
    return F & (ln.P | ln.M);
}

// Flip phase 3: the row's 64 spin bytes after the flips, written from the masks alone (code = 1 + plus - minus per
// byte, no borrows between bytes); a row without a flipped site is not touched.  Even x live in array gy & 1.
BC_HD void cl_apply_flips(const ClGeom& g, uint8_t* c0, uint8_t* c1, uint32_t rep, int x0, int gy, uint64_t P, uint64_t M, uint64_t F) {
    if (!F) return;
    const uint64_t Pn = (P & ~F) | (M & F), Mn = (M & ~F) | (P & F);
    const size_t off = cl_row_offset(g, rep, gy) + (size_t)(x0 >> 1);
    uint8_t* ev = ((gy & 1) ? c1 : c0) + off;
    uint8_t* od = ((gy & 1) ? c0 : c1) + off;
    const int vecs = (g.Wc - (x0 >> 1)) >= 32 ? 2 : ((g.Wc - (x0 >> 1)) >= 16 ? 1 : 0);
    const uint32_t fe = cl_squeeze(F), fo = cl_squeeze(F >> 1);
    const uint32_t pe = cl_squeeze(Pn), po = cl_squeeze(Pn >> 1), me = cl_squeeze(Mn), mo = cl_squeeze(Mn >> 1);
    BC_UNROLL
    for (int v = 0; v < 2; ++v) {
        if (v >= vecs) break;
        const int sh = 16 * v;
        if ((fe >> sh) & 0xFFFFu) {
            uint4 a;
            a.x = 0x01010101u + cl_bytes(pe >> sh) - cl_bytes(me >> sh);
            a.y = 0x01010101u + cl_bytes(pe >> (sh + 4)) - cl_bytes(me >> (sh + 4));
            a.z = 0x01010101u + cl_bytes(pe >> (sh + 8)) - cl_bytes(me >> (sh + 8));
            a.w = 0x01010101u + cl_bytes(pe >> (sh + 12)) - cl_bytes(me >> (sh + 12));
            *reinterpret_cast<uint4*>(ev + 16 * v) = a;
        }
        if ((fo >> sh) & 0xFFFFu) {
            uint4 b;
            b.x = 0x01010101u + cl_bytes(po >> sh) - cl_bytes(mo >> sh);
            b.y = 0x01010101u + cl_bytes(po >> (sh + 4)) - cl_bytes(mo >> (sh + 4));
            b.z = 0x01010101u + cl_bytes(po >> (sh + 8)) - cl_bytes(mo >> (sh + 8));
            b.w = 0x01010101u + cl_bytes(po >> (sh + 12)) - cl_bytes(mo >> (sh + 12));
            *reinterpret_cast<uint4*>(od + 16 * v) = b;
        }
    }
}

// Label phase for the measurements: the smallest site of the cluster of every site of the thread's row
// (0xFFFFFFFF for a zero spin) into `out` (the row's 64 entries of a row-major label array; clipped tiles: w entries)
BC_HD void cl_row_labels(int lane, const ClGeom& g, const ClTileId& t, const ClLabelSmem& sm, const ClLane& ln,
                         const ClGlobal& G, uint32_t* out) {
    uint32_t cur = 0xFFFFFFFFu;
    const uint64_t S = ln.P | ln.M;
    const bool vec = (g.L & 3) == 0;  // the row's entries are 16-byte aligned: four labels per store
    for (int x4 = 0; x4 < t.w; x4 += 4) {
        uint32_t v[4];
        BC_UNROLL
        for (int k = 0; k < 4; ++k) {
            const int x = x4 + k;
            if ((ln.RS >> x) & 1ull) cur = cl_cluster_site(g, t, sm.par[lane * CL_T + x], G);
            v[k] = ((S >> x) & 1ull) ? cur : 0xFFFFFFFFu;
        }
        if (vec && x4 + 4 <= t.w) {
            uint4 q; q.x = v[0]; q.y = v[1]; q.z = v[2]; q.w = v[3];
            *reinterpret_cast<uint4*>(out + x4) = q;
        } else {
            for (int k = 0; k < 4 && x4 + k < t.w; ++k) out[x4 + k] = v[k];
        }
    }
}

}  // namespace bc
