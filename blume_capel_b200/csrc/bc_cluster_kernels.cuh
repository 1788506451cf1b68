// bc_cluster_kernels.cuh -- Swendsen-Wang cluster move, cluster labels, cluster-size moments.
// Part of the sm_100a Blume-Capel engine; included through bc_kernels.cuh.
//
// Cluster move (all clusters of the +-1 spins at once; zeros never join, like wolff() of main.cpp:113-154; bond
// probability 1 - exp(-2J/T), main.cpp:115) in three launches, straight on the compact colour arrays:
//   cl_label_kernel   64 threads per 64 x 64 tile: row masks, random bonds, runs, unions in shared memory
//                     (bc_cluster_tile.cuh); leaves 2 B / site of labels + 1 KB of masks and perimeter roots per tile
//   cl_border_kernel  one thread per bond that leaves a tile: global union-find over the flagged tile roots
//   cl_flip_kernel    64 threads per tile: one flip bit per cluster (keyed by the cluster's smallest site), flip masks
//                     per row, spin bytes rewritten in place
// Algorithmic traffic per site: 1 B spins read + 2 B labels written; 2 B labels + 1 B spins read, < 1 B spins
// written: about 7 B / site (round 1: about 22 B / site through a row-major staging copy and int32 labels).
#pragma once
#include "bc_cluster_tile.cuh"
#include "bc_common.cuh"

namespace bc {

struct ClArgs {
    ClGeom g;
    uint8_t* c0;
    uint8_t* c1;
    ClTileOut out;
    ClRng rng;
    uint32_t n_tiles;      // tiles of all replicas
    uint32_t* labels;      // cl_labels_kernel: [replica][L*L] smallest site of the cluster, 0xFFFFFFFF for a zero spin
    // the global union-find: one piece per rank ([replica][rows per rank * L] each; a whole lattice: one piece)
    uint32_t* gbase[CL_MAX_RANKS];
    uint32_t sites_per_rank;   // 0 = one piece
    size_t rep_stride;         // entries between the pieces of consecutive replicas
    // slab of a multi-rank decomposition: the perimeter roots of the rank below (peer memory) and its first global row
    const uint16_t* proot_below;
    int y_glob0_below;
};

__device__ __forceinline__ ClGlobal cl_global_of(const ClArgs& A, uint32_t rep) {
    ClGlobal G;
#pragma unroll
    for (int k = 0; k < CL_MAX_RANKS; ++k) G.base[k] = A.gbase[k] ? A.gbase[k] + (size_t)rep * A.rep_stride : nullptr;
    G.sites_per_rank = A.sites_per_rank;
    return G;
}

// One CTA of 64 threads per tile, thread l = tile row l (10.75 KB of shared memory: 21 tiles per SM)
__global__ void __launch_bounds__(CL_LANES) cl_label_kernel(const __grid_constant__ ClArgs A) {
    __shared__ __align__(16) ClTileSmem sm;
    const int lane = threadIdx.x;
    const uint32_t tile = blockIdx.x;
    const ClTileId t = cl_tile_id(A.g, tile);
    ClLane ln;
    cl_phase_load(lane, A.g, t, A.c0, A.c1, sm, ln);
    __syncthreads();
    cl_phase_bonds(lane, A.g, t, A.c0, A.c1, A.rng, sm, ln);
    __syncthreads();
    cl_phase_unions(lane, t, sm, ln);
    __syncthreads();
    while (__syncthreads_or(cl_phase_jump(lane, sm, ln))) {}
    cl_phase_perimeter(lane, t, sm, ln, A.out.proot + (size_t)tile * 256);
    __syncthreads();
    cl_phase_finish(lane, A.g, t, sm, ln, cl_global_of(A, t.rep));
    __syncthreads();
    // what the following kernels need: labels (raw 8 KB copy), bond masks of the rows, bonds into the tile below
    const uint4* src = reinterpret_cast<const uint4*>(sm.par);
    uint4* dst = reinterpret_cast<uint4*>(A.out.par + (size_t)tile * CL_SITES);
#pragma unroll
    for (int i = lane; i < CL_SITES * 2 / 16; i += CL_LANES) dst[i] = src[i];
    A.out.hb[(size_t)tile * CL_T + lane] = ln.HB;
    A.out.pm[(size_t)tile * 2 * CL_T + lane] = ln.P;
    A.out.pm[(size_t)tile * 2 * CL_T + CL_T + lane] = ln.M;
    if (lane == t.h - 1) A.out.vbb[tile] = ln.VB;
}

// One thread per bond that can leave a tile: thread j < 64 the bond of row j of the tile's last column to the right,
// thread 64 + j the bond of column j of the tile's last row downwards (periodic wrap included; an open lattice has
// no candidates there).  A union only returns once its two trees are one, so a single pass is complete.
__global__ void __launch_bounds__(128) cl_border_kernel(const __grid_constant__ ClArgs A) {
    const uint32_t tile = blockIdx.x;
    const ClTileId t = cl_tile_id(A.g, tile);
    const int j = threadIdx.x & 63, down = threadIdx.x >> 6;
    uint32_t a, b;
    ClTileId n = t;
    ClGeom gn = A.g;  // geometry the neighbour tile's site indices are computed with
    if (!down) {
        if (j >= t.h || !((A.out.hb[(size_t)tile * CL_T + j] >> (t.w - 1)) & 1ull)) return;
        n = cl_tile_id(A.g, tile - (uint32_t)t.tx + (uint32_t)((t.tx + 1) % A.g.tiles_x));
        a = A.out.proot[(size_t)tile * 256 + 192 + j];
        b = A.out.proot[(size_t)n.tile * 256 + 128 + j];
    } else {
        if (j >= t.w || !((A.out.vbb[tile] >> j) & 1ull)) return;
        const bool last = t.ty + 1 == A.g.tiles_y;
        n = cl_tile_id(A.g, tile - (uint32_t)(t.ty * A.g.tiles_x) + (uint32_t)(((t.ty + 1) % A.g.tiles_y) * A.g.tiles_x));
        a = A.out.proot[(size_t)tile * 256 + 64 + j];
        if (last && A.g.multi) {  // the tile below belongs to the next rank: same tile geometry, its rows, its perimeter roots
            gn.y_glob0 = A.y_glob0_below;
            b = A.proot_below[(size_t)n.tile * 256 + j];
        } else {
            b = A.out.proot[(size_t)n.tile * 256 + j];
        }
    }
    cl_gunion(cl_global_of(A, t.rep), cl_global_site(A.g, t, a), cl_global_site(gn, n, b));
}

// reload a labelled tile: labels into shared memory, masks and runs of the thread's row into registers
__device__ __forceinline__ void cl_reload_tile(const ClArgs& A, const ClTileId& t, int lane, ClLabelSmem& sm, ClLane& ln) {
    const uint4* src = reinterpret_cast<const uint4*>(A.out.par + (size_t)t.tile * CL_SITES);
    uint4* dst = reinterpret_cast<uint4*>(sm.par);
#pragma unroll
    for (int i = lane; i < CL_SITES * 2 / 16; i += CL_LANES) dst[i] = src[i];
    for (int i = 0; i < CL_SITES / 32 / CL_LANES; ++i) sm.flag[(CL_SITES / 32 / CL_LANES) * lane + i] = 0u;
    cl_phase_reload(lane, t, A.out.pm + (size_t)t.tile * 2 * CL_T, A.out.hb + (size_t)t.tile * CL_T, ln);
    __syncthreads();
}

// (8.7 KB of shared memory per tile: 26 tiles = 52 warps per SM; the kernel is latency-bound)
__global__ void __launch_bounds__(CL_LANES) cl_flip_kernel(const __grid_constant__ ClArgs A) {
    __shared__ __align__(16) ClLabelSmem sm;
    const int lane = threadIdx.x;
    const ClTileId t = cl_tile_id(A.g, blockIdx.x);
    ClLane ln;
    cl_reload_tile(A, t, lane, sm, ln);
    cl_phase_decide(lane, A.g, t, A.rng, sm, ln, cl_global_of(A, t.rep));
    __syncthreads();
    if (lane < t.h) cl_apply_flips(A.g, A.c0, A.c1, t.rep, t.x0, t.y0 + lane, ln.P, ln.M, cl_flip_mask(lane, sm, ln));
}

// Cluster labels for the measurements (cluster-size moments, gap-size histogram, cluster lists of the dumps):
// labels[site] = smallest site of the site's cluster, row-major like the reference's lattice (main.cpp:98-101).
__global__ void __launch_bounds__(CL_LANES) cl_labels_kernel(const __grid_constant__ ClArgs A) {
    __shared__ __align__(16) ClLabelSmem sm;
    const int lane = threadIdx.x;
    const ClTileId t = cl_tile_id(A.g, blockIdx.x);
    ClLane ln;
    cl_reload_tile(A, t, lane, sm, ln);
    const size_t N = (size_t)A.g.L * (size_t)A.g.L;
    if (lane < t.h)
        cl_row_labels(lane, A.g, t, sm, ln, cl_global_of(A, t.rep),
                      A.labels + (size_t)t.rep * N + (size_t)(t.y0 + lane) * (size_t)A.g.L + (size_t)t.x0);
}

// ------------------------------------------------------------------------------------------
// Cluster sizes from the labels: size[root] += 1 for every non-zero site, then per replica the sum of size^2, the
// largest size and the number of clusters (gamma_nu.cpp:35-60 needs sum size^2 - max^2).  2-D grids: x = site index
// inside a replica in 32 bits, y = replica with a stride loop.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cl_count_kernel(const uint32_t* __restrict__ labels, uint32_t* __restrict__ size, uint32_t N, int n_rep) {
    __shared__ uint32_t s_root;  // one root per CTA is counted in shared memory first (a spanning cluster would
    __shared__ int s_count;      // otherwise take one global atomic per warp on a single address)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    for (int rep = blockIdx.y; rep < n_rep; rep += gridDim.y) {
        const size_t base = (size_t)rep * N;
        if (threadIdx.x == 0) { s_root = 0xFFFFFFFFu; s_count = 0; }
        __syncthreads();
        // consecutive sites mostly share a root: one atomic per run of equal roots inside a warp (the first lane of a
        // run adds the run's length), not one per site; 0xFFFFFFFF marks zero spins and never equals a root
        const uint32_t root = i < N ? labels[base + i] : 0xFFFFFFFFu;
        const bool live = root != 0xFFFFFFFFu;
        const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, root, 1);
        const bool head = lane == 0 || root != prev;
        const unsigned heads = __ballot_sync(0xFFFFFFFFu, head);
        const unsigned later = heads & ~((2u << lane) - 1u);  // run heads behind this lane
        const int run = (later ? __ffs(later) - 1 : 32) - lane;
        if (head && live && run >= 16) atomicCAS(&s_root, 0xFFFFFFFFu, root);  // the first long run names the CTA's root
        __syncthreads();
        if (head && live) {
            if (root == s_root) atomicAdd(&s_count, run);
            else atomicAdd(&size[base + root], (uint32_t)run);
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_count) atomicAdd(&size[base + s_root], (uint32_t)s_count);
        __syncthreads();
    }
}

// One CTA reduces CL_MOM_PER_CTA consecutive entries: per-thread partial sums, warp shuffles, shared atomics, then
// three global atomics per CTA (a CTA whose range straddles two replicas falls back to one atomic per entry).
constexpr int CL_MOM_PER_CTA = 256 * 16;
__global__ void __launch_bounds__(256)
cl_moments_kernel(const uint32_t* __restrict__ size, long long N, int n_rep,
                  unsigned long long* __restrict__ out /* [rep][3]: sum sq, max, count */) {
    __shared__ unsigned long long acc[3];
    const long long total = N * n_rep;
    const long long first = (long long)blockIdx.x * CL_MOM_PER_CTA;
    const long long last = min(first + CL_MOM_PER_CTA, total) - 1;
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
        const unsigned sz = size[g];
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

}  // namespace bc
