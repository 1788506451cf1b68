// bc_cluster_sort.cuh -- cluster member lists and the gap-size histogram from the site labels.
// Part of the sm_100a Blume-Capel engine; included through bc_kernels.cuh.
//
// The reference builds every cluster as a sorted vector of sites (form_clusters + std::sort, main.cpp:179-230, 236-239)
// and its post-processor walks those vectors (gap_size_statistics.cpp:100-126).  On the GPU the same object is ONE
// array per lattice: the occupied sites sorted by (cluster label, site) -- labels are the clusters' smallest sites
// (cl_labels_kernel), so the clusters appear in the order of the reference's files and their members ascend.  It is
// produced by a stable least-significant-digit radix sort of the sites by label (8-bit digits, the sites enter in
// ascending order and stability keeps them so inside a cluster), and consumed by
//   gap_hist_kernel     the gap-size histogram of gap_size_statistics.cpp, quirks included, one thread per member;
//   cluster_heads_kernel / list compaction   cluster start offsets and signs for the dump writer (main.cpp:232-257).
#pragma once
#include "bc_common.cuh"

namespace bc {

constexpr int SORT_CHUNK = 2048;      // elements per warp (64 rounds of 32)
constexpr int SORT_WARPS = 8;         // warps per CTA
constexpr uint32_t SORT_EMPTY = 0xFFFFFFFFu;

struct SortArgs {
    const uint32_t* labels;   // pass 0: [replica][N] labels (key), the value is the site index; nullptr afterwards
    const uint2* in;          // later passes: [replica][cap] (key, site) pairs
    uint2* out;               // [replica][cap]
    uint32_t* hist;           // [replica][256][n_chunks] digit counts per chunk, exclusive-scanned in place
    const uint32_t* count_in; // [replica] number of pairs in `in` (later passes)
    uint32_t n;               // pass 0: N
    uint32_t cap;             // pairs per replica the buffers hold
    uint32_t n_chunks;
    int shift;                // digit = (key >> shift) & 255
};

__device__ __forceinline__ bool sort_fetch(const SortArgs& A, uint32_t rep, uint32_t i, uint32_t count, uint32_t* key, uint32_t* val) {
    if (i >= count) return false;
    if (A.labels) {
        *key = A.labels[(size_t)rep * A.n + i];
        *val = i;
        return *key != SORT_EMPTY;
    }
    const uint2 p = A.in[(size_t)rep * A.cap + i];
    *key = p.x; *val = p.y;
    return true;
}

// digit histogram of every chunk: grid = (ceil(n_chunks / SORT_WARPS), replicas)
__global__ void __launch_bounds__(32 * SORT_WARPS) sort_hist_kernel(const __grid_constant__ SortArgs A) {
    __shared__ uint32_t s_hist[SORT_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * SORT_WARPS + warp, rep = blockIdx.y;
    if (chunk >= A.n_chunks) return;
    const uint32_t count = A.labels ? A.n : A.count_in[rep];
    uint32_t* h = s_hist[warp];
    for (int d = lane; d < 256; d += 32) h[d] = 0;
    __syncwarp();
    const uint32_t first = chunk * SORT_CHUNK;
    for (uint32_t r = 0; r < SORT_CHUNK; r += 32) {
        uint32_t key = 0, val = 0;
        const bool live = sort_fetch(A, rep, first + r + lane, count, &key, &val);
        const unsigned act = __ballot_sync(0xFFFFFFFFu, live);
        if (live) {
            const uint32_t d = (key >> A.shift) & 255u;
            const unsigned same = __match_any_sync(act, d);
            if ((unsigned)lane == (unsigned)(__ffs(same) - 1)) h[d] += __popc(same);
        }
        __syncwarp();
    }
    for (int d = lane; d < 256; d += 32) A.hist[((size_t)rep * 256 + d) * A.n_chunks + chunk] = h[d];
}

// Exclusive prefix sum of n uint32 per replica in place, one CTA per replica (n <= a few million: the digit counts of
// a sort pass, digit-major, so that the prefix of entry (digit, chunk) is where the chunk's elements of that digit
// start; or the head flags of the sorted list).  total[replica] = the sum.
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t* __restrict__ data, uint32_t n, size_t stride, uint32_t* __restrict__ total) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry, s_round;
    uint32_t* d = data + (size_t)blockIdx.x * stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 4096) {  // 4 consecutive entries per thread per round
        const uint32_t i = base + 4 * threadIdx.x;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = i + k < n ? d[i + k] : 0u;
        const uint32_t mine = v[0] + v[1] + v[2] + v[3];
        uint32_t inc = mine;  // inclusive scan over the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the 32 warp totals
            const uint32_t w = s_warp[lane];
            uint32_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, winc, o);
                if (lane >= o) winc += up;
            }
            s_warp[lane] = winc - w;
            if (lane == 31) s_round = winc;
        }
        __syncthreads();
        uint32_t run = s_carry + s_warp[warp] + inc - mine;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i + k < n) d[i + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_round;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) total[blockIdx.x] = s_carry;
}

// stable scatter: every warp re-reads its chunk in order, 32 elements at a time; an element's destination is the
// scanned start of (its digit, this chunk) plus the number of earlier elements of the chunk with the same digit
__global__ void __launch_bounds__(32 * SORT_WARPS) sort_scatter_kernel(const __grid_constant__ SortArgs A) {
    __shared__ uint32_t s_base[SORT_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * SORT_WARPS + warp, rep = blockIdx.y;
    if (chunk >= A.n_chunks) return;
    const uint32_t count = A.labels ? A.n : A.count_in[rep];
    uint32_t* b = s_base[warp];
    for (int d = lane; d < 256; d += 32) b[d] = A.hist[((size_t)rep * 256 + d) * A.n_chunks + chunk];
    __syncwarp();
    const uint32_t first = chunk * SORT_CHUNK;
    for (uint32_t r = 0; r < SORT_CHUNK; r += 32) {
        uint32_t key = 0, val = 0;
        const bool live = sort_fetch(A, rep, first + r + lane, count, &key, &val);
        const unsigned act = __ballot_sync(0xFFFFFFFFu, live);
        if (live) {
            const uint32_t d = (key >> A.shift) & 255u;
            const unsigned same = __match_any_sync(act, d);
            const uint32_t pos = b[d] + __popc(same & ((1u << lane) - 1u));
            A.out[(size_t)rep * A.cap + pos] = make_uint2(key, val);
            __syncwarp(act);
            if ((unsigned)lane == (unsigned)(__ffs(same) - 1)) b[d] += __popc(same);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Gap-size histogram of one configuration per replica (gap_size_statistics.cpp:58-81, 100-144).
// `list` holds the occupied sites sorted by (cluster, site); a cluster's label is its
// first member m0.  The tool walks a cluster's members m1, m2, ... (never m0) at positions s' = m + m0 / L - m0 --
// its tracker starts at the ROW of the first site, gap_size_statistics.cpp:104 -- closes a "row" whenever s' / L
// changes, and counts per row of n > 1 entries the n - 1 gaps between neighbours plus, for n > 2, the wrap-around
// gap last - first, every gap folded to min(gap, L - gap) (variant 1, open boundaries: unfolded, no wrap-around
// gap).  One thread per member: the gap to its predecessor if both lie in one row; the first member of a row also
// finds the row's end by bisection and adds the wrap-around gap.
// ------------------------------------------------------------------------------------------
constexpr int GAP_SMEM_BINS = 8192;
__global__ void __launch_bounds__(256)
gap_hist_kernel(const uint2* __restrict__ list, const uint32_t* __restrict__ count, uint32_t cap, int L, int variant,
                unsigned long long* __restrict__ hist /* [replica][bins] */) {
    __shared__ uint32_t s_hist[GAP_SMEM_BINS];
    const uint32_t rep = blockIdx.y, M = count[rep];
    const int bins = variant == 0 ? L / 2 : L - 1;
    const bool use_smem = bins <= GAP_SMEM_BINS;
    if (use_smem) for (int b = threadIdx.x; b < bins; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    const uint2* ls = list + (size_t)rep * cap;
    unsigned long long* out = hist + (size_t)rep * bins;
    auto add = [&](long long gap) {
        if (variant == 0 && L - gap < gap) gap = L - gap;
        if (gap < 1 || gap > bins) return;  // cannot happen for a valid list
        if (use_smem) atomicAdd(&s_hist[gap - 1], 1u); else atomicAdd(out + (gap - 1), 1ull);
    };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const uint2 me = ls[i];
        if (me.y == me.x || i == 0) continue;  // the cluster's first site is never recorded (a list starts with one)
        const long long c = (long long)(me.x / (uint32_t)L) - (long long)me.x;
        const long long row = ((long long)me.y + c) / L;
        const uint2 prev = ls[i - 1];  // i > 0: the first element of a list is a cluster's first site
        const bool row_first = prev.y == prev.x || ((long long)prev.y + c) / L != row;
        if (!row_first) { add((long long)me.y - (long long)prev.y); continue; }
        if (variant != 0) continue;
        // end of the row: first index whose element is in another cluster or at or beyond position (row + 1) * L
        const long long limit = (row + 1) * L - c;  // in site units
        uint32_t lo = i + 1, hi = M;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            const uint2 e = ls[mid];
            if (e.x == me.x && (long long)e.y < limit) lo = mid + 1; else hi = mid;
        }
        if (lo - i > 2) add((long long)ls[lo - 1].y - (long long)me.y);
    }
    __syncthreads();
    if (use_smem)
        for (int b = threadIdx.x; b < bins; b += blockDim.x)
            if (s_hist[b]) atomicAdd(out + b, (unsigned long long)s_hist[b]);
}

// ------------------------------------------------------------------------------------------
// Cluster list for the dump writer (main.cpp:232-257 prints, per cluster in the order of the smallest site, the
// sign and the ascending members): head flag of every list element, scanned to the cluster's ordinal, then the
// start offset and the sign of every cluster.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cluster_heads_kernel(const uint2* __restrict__ list, const uint32_t* __restrict__ count, uint32_t cap, uint32_t* __restrict__ flags) {
    const uint32_t rep = blockIdx.y, M = count[rep];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const uint2 me = list[(size_t)rep * cap + i];
        flags[(size_t)rep * cap + i] = me.y == me.x ? 1u : 0u;
    }
}

// sites[i] = member i of the list; starts[ordinal] = i and sign[ordinal] = spin for every head (ordinal = scanned flag)
__global__ void __launch_bounds__(256)
cluster_lists_kernel(const uint2* __restrict__ list, const uint32_t* __restrict__ count, uint32_t cap, const uint32_t* __restrict__ ordinal,
                     const uint8_t* __restrict__ c0, const uint8_t* __restrict__ c1, int L, int Wc, int rows_arr, int row_first,
                     uint32_t* __restrict__ sites, uint32_t* __restrict__ starts, int8_t* __restrict__ sign) {
    const uint32_t rep = blockIdx.y, M = count[rep];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const uint2 me = list[(size_t)rep * cap + i];
        sites[(size_t)rep * cap + i] = me.y;
        if (me.y != me.x) continue;
        const uint32_t o = ordinal[(size_t)rep * cap + i];
        const uint32_t x = me.y % (uint32_t)L, y = me.y / (uint32_t)L;
        const uint8_t code = (((x + y) & 1u) ? c1 : c0)[((size_t)rep * rows_arr + row_first + y) * Wc + (x >> 1)];
        starts[(size_t)rep * cap + o] = i;
        sign[(size_t)rep * cap + o] = (int8_t)((int)code - 1);
    }
}

}  // namespace bc
