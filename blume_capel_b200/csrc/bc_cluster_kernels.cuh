// bc_cluster_kernels.cuh -- Swendsen-Wang cluster move and cluster-size moments.
// Part of the sm_100a Blume-Capel engine; included through bc_kernels.cuh.
#pragma once
#include "bc_common.cuh"

namespace bc {

// ------------------------------------------------------------------------------------------
// Cluster move (Swendsen-Wang on the +-1 spins; zeros never join, like wolff() of main.cpp:113-154; bond
// probability 1 - exp(-2J/T), main.cpp:115).  Works on the row-major int8 staging copy of the lattices:
// bonds from Philox stream 3, connected components by union-find where the larger root is always linked
// under the smaller one (the root of a cluster is its smallest site index, so the result does not depend
// on the order of the unions), one flip bit per cluster from Philox stream 4.
//
// Two levels: sw_local_kernel labels one SW_TX x SW_TY tile per CTA entirely in shared memory (bonds drawn
// in the same kernel) and writes parent[site] = the tile-local root as a global site index; the bonds that
// leave a tile (last column / last row of the tile, including the periodic wrap) are merged afterwards by
// sw_border_kernel with global compare-and-swap.  A union only returns once its two trees are one, so a
// single pass over the bonds is complete: no convergence loop, no host round trip.
// ------------------------------------------------------------------------------------------
constexpr int SW_TX = 64, SW_TY = 64, SW_THREADS = 512;

// find with path halving; every value ever stored in lab[a] is an ancestor of a, so concurrent finds and
// unions only ever see valid (possibly longer) paths.  Roots are never written by the halving store.
__device__ __forceinline__ int32_t sw_find_halve(volatile int32_t* lab, int32_t a) {
    int32_t p = lab[a];
    while (p != a) {
        const int32_t g = lab[p];
        if (g != p) lab[a] = g;
        a = p;
        p = g;
    }
    return a;
}

// link the larger root under the smaller one; retry from the current roots if the larger one stopped being a
// root meanwhile (trees only ever merge)
__device__ __forceinline__ void sw_union(volatile int32_t* lab, int32_t a, int32_t b) {
    int32_t ra = sw_find_halve(lab, a), rb = sw_find_halve(lab, b);
    while (ra != rb) {
        if (ra < rb) { const int32_t t = ra; ra = rb; rb = t; }
        const int32_t old = atomicCAS(const_cast<int32_t*>(lab) + ra, ra, rb);
        if (old == ra) break;
        ra = sw_find_halve(lab, old);
        rb = sw_find_halve(lab, rb);
    }
}

__device__ __forceinline__ int32_t sw_find(const int32_t* parent, int32_t a) {
    int32_t p = parent[a];
    while (p != a) { a = p; p = parent[a]; }
    return a;
}

// Root of site i once all unions are done (no concurrent linking): site -> tile-local root t (one hop) -> chain of
// tile roots.  The chain above t is shared by the whole tile; whoever walks it points t straight at the root
// (plain cached loads: a stale value is an older, still valid ancestor).
__device__ __forceinline__ int32_t sw_root_shorten(int32_t* par, int32_t i) {
    const int32_t t = par[i];
    const int32_t pt = par[t];
    if (pt == t) return t;
    const int32_t root = sw_find(par, pt);
    if (root != pt) par[t] = root;
    return root;
}

// bonds site (x, y) with spin s could have: bit 0 = +x neighbour equal, bit 1 = +y neighbour equal (periodic wrap
// unless open; a site is never bonded to itself, which only matters for L = 1)
__device__ __forceinline__ uint32_t sw_candidates(const int8_t* __restrict__ sp, int L, int open, int x, int y, int8_t s) {
    uint32_t cand = 0;
    if (s == 0) return 0;
    const int32_t i = x + y * L;
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        int nx = x + (dir == 0), ny = y + (dir == 1);
        if (nx == L || ny == L) {
            if (open) continue;
            nx = nx == L ? 0 : nx;
            ny = ny == L ? 0 : ny;
        }
        const int32_t j = nx + ny * L;
        if (j != i && sp[j] == s) cand |= 1u << dir;
    }
    return cand;
}

// grid = n_rep * tiles_x * tiles_y CTAs.  A warp owns whole tile rows (lane l = sites 2l, 2l+1 of the row):
//  1. bonds: the two sites of a lane are the two halves of ONE Philox call when L is even (site i uses words
//     2(i&1), 2(i&1)+1 of counter i>>1); bonds[] (bit 0 = bond to the +x neighbour open, bit 1 = +y) is only
//     written for the sites of a tile's last column / last row, the ones sw_border_kernel reads;
//  2. horizontal runs from two ballots: every site starts with the label of the first site of its run, so
//     the horizontal bonds cost no union at all;
//  3. vertical bonds: one union per bond unless the bond to its left joins the same two runs;
//  4. parent[site] = the tile-local root as a global site index.
__global__ void __launch_bounds__(SW_THREADS, 3)
sw_local_kernel(const int8_t* __restrict__ spins, uint8_t* __restrict__ bonds, int32_t* __restrict__ parent, int L,
                int tiles_x, int tiles_y, int open, int always, uint32_t bond_thr, uint32_t key0, uint32_t key1,
                uint32_t c_lo, uint32_t c_hi) {
    static_assert(SW_TX == 64 && SW_THREADS % 32 == 0, "a warp covers one 64-site tile row");
    __shared__ __align__(16) int32_t lab[SW_TX * SW_TY];
    __shared__ __align__(16) uint8_t bnd[SW_TX * SW_TY];
    __shared__ int s_next_row;
    if (threadIdx.x == 0) s_next_row = 0;
    constexpr int WARPS = SW_THREADS / 32;
    const int tiles = tiles_x * tiles_y;
    const int rep = (int)(blockIdx.x / (unsigned)tiles), tile = (int)(blockIdx.x - (unsigned)rep * tiles);
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int x0 = tx * SW_TX, y0 = ty * SW_TY;
    const int w = min(SW_TX, L - x0), h = min(SW_TY, L - y0);
    const size_t N = (size_t)L * L;
    const int8_t* sp = spins + (size_t)rep * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = 2 * lane;
    const bool v0 = lx < w, v1 = lx + 1 < w;
    const uint32_t k1r = key1 ^ (uint32_t)rep;

    for (int ly = warp; ly < h; ly += WARPS) {
        const int x = x0 + lx, y = y0 + ly;
        const int32_t i0 = x + y * L;
        uint32_t b0, b1;
        if (!(L & 1)) {
            // even L: x and i0 are even, a lane's two sites and the two below them are aligned 16-bit loads, the right
            // neighbour of the second site is the next lane's first site (a load only at the tile's last pair);
            // an absent neighbour (open edge) reads as spin 0, which never equals a non-zero spin
            const uint32_t own = v0 ? *reinterpret_cast<const uint16_t*>(sp + i0) : 0u;
            const int yd = y + 1 == L ? 0 : y + 1;
            const uint32_t dn = (v0 && !(open && y + 1 == L)) ? *reinterpret_cast<const uint16_t*>(sp + yd * L + x) : 0u;
            const uint32_t s0 = own & 0xFFu, s1 = own >> 8;
            uint32_t r1 = __shfl_down_sync(0xFFFFFFFFu, s0, 1);
            if (v0 && lx + 2 >= w) {
                const int xr = x + 2;
                r1 = xr < L ? (uint32_t)(uint8_t)sp[y * L + xr] : (open ? 0u : (uint32_t)(uint8_t)sp[y * L]);
            }
            b0 = (s0 != 0 && s1 == s0 ? 1u : 0u) | (s0 != 0 && (dn & 0xFFu) == s0 ? 2u : 0u);
            b1 = (s1 != 0 && r1 == s1 ? 1u : 0u) | (s1 != 0 && (dn >> 8) == s1 ? 2u : 0u);
        } else {
            const int8_t s0 = v0 ? sp[i0] : (int8_t)0, s1 = v1 ? sp[i0 + 1] : (int8_t)0;
            b0 = v0 ? sw_candidates(sp, L, open, x, y, s0) : 0u;
            b1 = v1 ? sw_candidates(sp, L, open, x + 1, y, s1) : 0u;
        }
        if ((b0 | b1) && !always) {
            uint32_t r[4];
            if (!(i0 & 1)) {  // both sites in one call
                philox4x32_10((uint32_t)(i0 >> 1), c_lo, c_hi, STREAM_BOND << 1, key0, k1r, r);
                b0 &= (r[0] < bond_thr ? 1u : 0u) | (r[1] < bond_thr ? 2u : 0u);
                b1 &= (r[2] < bond_thr ? 1u : 0u) | (r[3] < bond_thr ? 2u : 0u);
            } else {          // odd L, odd row start: i0 is the second half of one call, i0 + 1 the first of the next
                if (b0) {
                    philox4x32_10((uint32_t)(i0 >> 1), c_lo, c_hi, STREAM_BOND << 1, key0, k1r, r);
                    b0 &= (r[2] < bond_thr ? 1u : 0u) | (r[3] < bond_thr ? 2u : 0u);
                }
                if (b1) {
                    philox4x32_10((uint32_t)((i0 + 1) >> 1), c_lo, c_hi, STREAM_BOND << 1, key0, k1r, r);
                    b1 &= (r[0] < bond_thr ? 1u : 0u) | (r[1] < bond_thr ? 2u : 0u);
                }
            }
        }
        const int l0 = ly * SW_TX + lx;
        *reinterpret_cast<uchar2*>(&bnd[l0]) = make_uchar2((uint8_t)b0, (uint8_t)b1);
        const bool last_row = ly == h - 1;
        if (v0 && (last_row || lx == w - 1)) bonds[(size_t)rep * N + i0] = (uint8_t)b0;
        if (v1 && (last_row || lx + 1 == w - 1)) bonds[(size_t)rep * N + i0 + 1] = (uint8_t)b1;
        // runs: E bit l = sites 2l and 2l+1 joined, A bit l = sites 2l-1 and 2l joined (inside the tile)
        const uint32_t E = __ballot_sync(0xFFFFFFFFu, (b0 & 1u) && v1);
        const uint32_t A = __ballot_sync(0xFFFFFFFFu, (b1 & 1u) && lx + 2 < w) << 1;
        const uint32_t both = A & (E << 1);                 // bit l: site 2l joined to site 2l-2
        const uint32_t stop = ~both & ((2u << lane) - 1u);   // bit 0 is always a stop
        const int z = 31 - __clz(stop);                      // the run of site 2l reaches back to pair z
        const int start0 = 2 * z - (int)((A >> z) & 1u);
        const int start1 = ((E >> lane) & 1u) ? start0 : lx + 1;
        *reinterpret_cast<int2*>(&lab[l0]) = make_int2(ly * SW_TX + start0, ly * SW_TX + start1);
    }
    __syncthreads();
    // rows are handed out dynamically: the cost of a row's unions varies (the ncu source page showed a fifth of the
    // kernel's stall samples at the barrier behind this phase with a static row assignment)
    for (;;) {
        int ly = 0;
        if (lane == 0) ly = atomicAdd(&s_next_row, 1);
        ly = __shfl_sync(0xFFFFFFFFu, ly, 0);
        if (ly + 1 >= h) break;
        const int l0 = ly * SW_TX + lx;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int l = l0 + k;
            if (!(bnd[l] & 2u)) continue;
            // the bond to the left already joins the same two runs
            if (lx + k > 0 && (bnd[l - 1] & 3u) == 3u && (bnd[l + SW_TX - 1] & 1u)) continue;
            sw_union(lab, l, l + SW_TX);
        }
    }
    __syncthreads();
    // local index order == global index order inside a tile, so the local root is the tile's smallest site
    for (int ly = warp; ly < h; ly += WARPS) {
        const int l0 = ly * SW_TX + lx;
        int32_t* dst = parent + (size_t)rep * N + (size_t)(x0 + lx) + (size_t)(y0 + ly) * L;
        if (v0 && v1 && !(L & 1)) {  // even L: the pair is one aligned 8-byte store
            // site -> first site of its run (one hop) -> root; whoever walks the chain above a run start points the
            // run start at the root for the rest of the run (sw_root_shorten on the shared-memory labels)
            const int32_t ra = sw_root_shorten(lab, l0), rb = sw_root_shorten(lab, l0 + 1);
            *reinterpret_cast<int2*>(dst) = make_int2((x0 + (ra & (SW_TX - 1))) + (y0 + ra / SW_TX) * L,
                                                      (x0 + (rb & (SW_TX - 1))) + (y0 + rb / SW_TX) * L);
            continue;
        }
        if (v0) {
            const int32_t r = sw_find(lab, l0);
            dst[0] = (x0 + (r & (SW_TX - 1))) + (y0 + r / SW_TX) * L;
        }
        if (v1) {
            const int32_t r = sw_find(lab, l0 + 1);
            dst[1] = (x0 + (r & (SW_TX - 1))) + (y0 + r / SW_TX) * L;
        }
    }
}

// grid = n_rep * tiles_x * tiles_y CTAs of SW_TX + SW_TY threads: thread j < SW_TY owns the +x bond of row j of
// the tile's last column, thread SW_TY + j the +y bond of column j of the tile's last row.
__global__ void __launch_bounds__(SW_TX + SW_TY)
sw_border_kernel(const uint8_t* __restrict__ bonds, int32_t* __restrict__ parent, int L, int tiles_x, int tiles_y) {
    const int tiles = tiles_x * tiles_y;
    const int rep = (int)(blockIdx.x / (unsigned)tiles), tile = (int)(blockIdx.x - (unsigned)rep * tiles);
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int x0 = tx * SW_TX, y0 = ty * SW_TY;
    const int w = min(SW_TX, L - x0), h = min(SW_TY, L - y0);
    const size_t N = (size_t)L * L;
    const int t = threadIdx.x;
    const int dir = t >= SW_TY;
    const int lx = dir ? t - SW_TY : w - 1, ly = dir ? h - 1 : t;
    if (lx >= w || ly >= h) return;
    const int x = x0 + lx, y = y0 + ly;
    const int32_t i = x + y * L;
    if (!(bonds[(size_t)rep * N + i] & (1u << dir))) return;
    const int nx = dir ? x : (x + 1 == L ? 0 : x + 1), ny = dir ? (y + 1 == L ? 0 : y + 1) : y;
    sw_union(parent + (size_t)rep * N, i, nx + ny * L);
}

// parent[i] = root(i) for every site (needed by the size histogram; the flip kernel finds roots itself)
// (The per-site kernels below run on a 2-D grid: x = site index inside a replica in 32 bits, y = replica with a
// stride loop -- a 64-bit division per thread to split a flat index costs more than the rest of these kernels.)
__global__ void __launch_bounds__(256) sw_compress_kernel(int32_t* __restrict__ parent, int L, int n_rep) {
    const uint32_t N = (uint32_t)L * (uint32_t)L;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if ((N & 3u) == 0) {  // four sites per thread: the dependent loads of four chains in flight (see sw_flip_kernel)
        const uint32_t i = 4u * tid;
        if (i >= N) return;
        for (int rep = blockIdx.y; rep < n_rep; rep += gridDim.y) {
            int32_t* par = parent + (size_t)rep * N;
            const int4 t4 = *reinterpret_cast<const int4*>(par + i);
            const int32_t t[4] = {t4.x, t4.y, t4.z, t4.w};
            int32_t pt[4], root[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) pt[b] = par[t[b]];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                root[b] = pt[b];
                if (pt[b] != t[b]) {
                    root[b] = sw_find(par, pt[b]);
                    if (root[b] != pt[b]) par[t[b]] = root[b];
                }
            }
            if (root[0] != t[0] || root[1] != t[1] || root[2] != t[2] || root[3] != t[3])
                *reinterpret_cast<int4*>(par + i) = make_int4(root[0], root[1], root[2], root[3]);
        }
    } else {
        if (tid >= N) return;
        for (int rep = blockIdx.y; rep < n_rep; rep += gridDim.y) {
            int32_t* par = parent + (size_t)rep * N;
            const int32_t p = par[tid];  // tile-local root (or the site itself)
            const int32_t r = sw_root_shorten(par, (int32_t)tid);
            if (r != p) par[tid] = r;
        }
    }
}

// cluster sizes: size[root] += 1 for every non-zero site (parent compressed), then per replica
// sum of size^2, largest size and number of clusters (gamma_nu.cpp:35-60 needs sum size^2 - max^2)
__global__ void __launch_bounds__(256)
sw_count_kernel(const int8_t* __restrict__ spins, const int32_t* __restrict__ parent, int32_t* __restrict__ size, int L, int n_rep) {
    __shared__ int32_t s_root;  // one root per CTA is counted in shared memory first (a spanning cluster would
    __shared__ int s_count;     // otherwise take one global atomic per warp on a single address)
    const uint32_t N = (uint32_t)L * (uint32_t)L;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    for (int rep = blockIdx.y; rep < n_rep; rep += gridDim.y) {
        const size_t base = (size_t)rep * N;
        if (threadIdx.x == 0) { s_root = -1; s_count = 0; }
        __syncthreads();
        const bool live = i < N && spins[base + i] != 0;
        // consecutive sites mostly share a root: one atomic per run of equal roots inside a warp (the first lane of a
        // run adds the run's length), not one per site; -1 marks zero spins and never equals a root
        const int32_t root = live ? parent[base + i] : -1;
        const int32_t prev = __shfl_up_sync(0xFFFFFFFFu, root, 1);
        const bool head = lane == 0 || root != prev;
        const unsigned heads = __ballot_sync(0xFFFFFFFFu, head);
        const unsigned later = heads & ~((2u << lane) - 1u);  // run heads behind this lane
        const int run = (later ? __ffs(later) - 1 : 32) - lane;
        if (head && live && run >= 16) atomicCAS(&s_root, -1, root);  // the first long run names the CTA's root
        __syncthreads();
        if (head && live) {
            if (root == s_root) atomicAdd(&s_count, run);
            else atomicAdd(&size[base + root], run);
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_count) atomicAdd(&size[base + s_root], s_count);
    }
}

// One CTA reduces SW_MOM_PER_CTA consecutive entries: per-thread partial sums, warp shuffles, shared atomics, then
// three global atomics per CTA (a CTA whose range straddles two replicas falls back to one atomic per entry).
constexpr int SW_MOM_PER_CTA = 256 * 16;
__global__ void __launch_bounds__(256)
sw_moments_kernel(const int32_t* __restrict__ size, int L, int n_rep,
                  unsigned long long* __restrict__ out /* [rep][3]: sum sq, max, count */) {
    __shared__ unsigned long long acc[3];
    const long long N = (long long)L * L, total = N * n_rep;
    const long long first = (long long)blockIdx.x * SW_MOM_PER_CTA;
    const long long last = min(first + SW_MOM_PER_CTA, total) - 1;
    const long long rep = first / N;
    if (rep != last / N) {
        for (long long g = first + threadIdx.x; g <= last; g += 256) {
            const unsigned long long sz = (unsigned long long)size[g];
            if (sz == 0) continue;
            unsigned long long* o = out + (g / N) * 3;
            atomicAdd(o + 0, sz * sz);
            atomicMax(o + 1, sz);
            atomicAdd(o + 2, 1ull);
        }
        return;
    }
    if (threadIdx.x < 3) acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long sq = 0, cnt = 0;
    unsigned mx = 0;
    for (long long g = first + threadIdx.x; g <= last; g += 256) {
        const unsigned sz = (unsigned)size[g];
        sq += (unsigned long long)sz * sz;
        cnt += sz != 0;
        mx = max(mx, sz);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sq += __shfl_xor_sync(0xFFFFFFFFu, sq, d);
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(&acc[0], sq);
        atomicMax(&acc[1], (unsigned long long)mx);
        atomicAdd(&acc[2], cnt);
    }
    __syncthreads();
    if (threadIdx.x == 0 && acc[2]) {
        unsigned long long* o = out + rep * 3;
        atomicAdd(o + 0, acc[0]);
        atomicMax(o + 1, acc[1]);
        atomicAdd(o + 2, acc[2]);
    }
}

// Flip decisions, one per CLUSTER: the cluster's root draws Philox(root) on stream 4 and leaves the bit in
// flipbit[root].  Roots are a few per cent of the sites, so a warp first compacts the roots of SW_ROOT_SPAN
// consecutive sites into a shared-memory list (ballot + popc) and then draws for 32 roots at a time -- every lane
// of a Philox call works, instead of one call per site in which all lanes of a cluster repeat the same draw.
constexpr int SW_ROOT_SPAN = 512;  // sites per warp
__global__ void __launch_bounds__(256)
sw_rootbit_kernel(const int8_t* __restrict__ spins, const int32_t* __restrict__ parent, uint8_t* __restrict__ flipbit,
                  int L, int n_rep, uint32_t key0, uint32_t key1, uint32_t c_lo, uint32_t c_hi) {
    __shared__ uint16_t s_list[8][SW_ROOT_SPAN];  // offsets of the warp's roots from its first site
    const uint32_t N = (uint32_t)L * (uint32_t)L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = (blockIdx.x * 8u + (uint32_t)warp) * SW_ROOT_SPAN;
    if (first >= N) return;
    uint16_t* list = s_list[warp];
    for (int rep = blockIdx.y; rep < n_rep; rep += gridDim.y) {
        const size_t base = (size_t)rep * N;
        int count = 0;
        if ((N & 3u) == 0) {
            // four groups of 128 sites, a lane owns 4 consecutive sites of each: all 8 vector loads are issued before
            // the first ballot (the scan is pure memory latency otherwise)
            uint32_t sp4[4];
            int4 pa4[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t i = first + 128u * g + 4u * lane;
                const bool in = i < N;
                sp4[g] = in ? *reinterpret_cast<const uint32_t*>(spins + base + i) : 0u;
                pa4[g] = in ? *reinterpret_cast<const int4*>(parent + base + i) : make_int4(-1, -1, -1, -1);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int32_t i = (int32_t)(first + 128u * g + 4u * lane);
                const int32_t pv[4] = {pa4[g].x, pa4[g].y, pa4[g].z, pa4[g].w};
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const bool root = ((sp4[g] >> (8 * b)) & 0xFFu) != 0 && pv[b] == i + b;
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, root);
                    if (root) list[count + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(128 * g + 4 * lane + b);
                    count += __popc(m);
                }
            }
        } else {
            for (int k = 0; k < SW_ROOT_SPAN; k += 32) {
                const uint32_t i = first + k + lane;
                const bool root = i < N && spins[base + i] != 0 && parent[base + i] == (int32_t)i;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, root);
                if (root) list[count + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(k + lane);
                count += __popc(m);
            }
        }
        __syncwarp();
        for (int j = lane; j < count; j += 32) {
            const uint32_t i = first + list[j];
            uint32_t w[4];
            philox4x32_10(i, c_lo, c_hi, STREAM_FLIP << 1, key0, key1 ^ (uint32_t)rep, w);
            flipbit[base + i] = (uint8_t)(w[0] & 1u);
        }
        __syncwarp();
    }
}

// Every non-zero site looks its cluster's bit up: site -> tile root -> cluster root -> bit, three dependent loads.
// A thread owns four consecutive sites and keeps the four chains in flight side by side (one site per thread ran at
// a quarter of the speed, all warps parked on the long scoreboard).  grid.x covers N / 4 threads (N % 4 == 0) or N.
__global__ void __launch_bounds__(256)
sw_flip_kernel(int8_t* __restrict__ spins, int32_t* __restrict__ parent, const uint8_t* __restrict__ flipbit, int L, int n_rep) {
    const uint32_t N = (uint32_t)L * (uint32_t)L;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if ((N & 3u) == 0) {
        const uint32_t i = 4u * tid;
        if (i >= N) return;
        for (int rep = blockIdx.y; rep < n_rep; rep += gridDim.y) {
            const size_t base = (size_t)rep * N;
            int32_t* par = parent + base;
            const uint32_t s4 = *reinterpret_cast<const uint32_t*>(spins + base + i);
            if (s4 == 0) continue;
            const int4 t4 = *reinterpret_cast<const int4*>(par + i);
            const int32_t t[4] = {t4.x, t4.y, t4.z, t4.w};
            int32_t pt[4], root[4];
            uint32_t bit[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) pt[b] = par[t[b]];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                root[b] = pt[b];
                if (pt[b] != t[b]) {  // the tile root is not the cluster root: walk on and point it there
                    root[b] = sw_find(par, pt[b]);
                    if (root[b] != pt[b]) par[t[b]] = root[b];
                }
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) bit[b] = flipbit[base + root[b]];
            uint32_t out = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t sb = (s4 >> (8 * b)) & 0xFFu;
                const uint32_t nb = (bit[b] && sb) ? ((0u - sb) & 0xFFu) : sb;  // -1 <-> +1, 0 stays
                out |= nb << (8 * b);
            }
            if (out != s4) *reinterpret_cast<uint32_t*>(spins + base + i) = out;
        }
    } else {
        if (tid >= N) return;
        for (int rep = blockIdx.y; rep < n_rep; rep += gridDim.y) {
            const size_t base = (size_t)rep * N;
            const int8_t s = spins[base + tid];
            if (s == 0) continue;
            const int32_t root = sw_root_shorten(parent + base, (int32_t)tid);
            if (flipbit[base + root]) spins[base + tid] = (int8_t)-s;
        }
    }
}

}  // namespace bc
