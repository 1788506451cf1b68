// bc_sweep_generic.cuh -- one-thread-per-site kernel for any L (fallback and independent second implementation).
// Part of the sm_100a Blume-Capel engine; included through bc_kernels.cuh.
#pragma once
#include "bc_update.cuh"  // exact_accept, warp_sum

namespace bc {

// ------------------------------------------------------------------------------------------
// Generic kernel: any L (even on a torus), one thread per same-colour site.  Used for sizes
// that are not a multiple of 32 (the reference's 8, 12, 16, 24, 48, 96) and as a second,
// independently written implementation the fast kernel is tested against.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t other_code(const uint8_t* oth_rep, int L, int Wc, int nx, int ny, bool open) {
    if (open) {
        if (nx < 0 || nx >= L || ny < 0 || ny >= L) return 1u;
    } else {
        nx = (nx < 0) ? nx + L : (nx >= L ? nx - L : nx);
        ny = (ny < 0) ? ny + L : (ny >= L ? ny - L : ny);
    }
    return oth_rep[(size_t)ny * Wc + (nx >> 1)];
}

template <bool MEASURE>
__global__ void __launch_bounds__(256)
sweep_generic_kernel(const __grid_constant__ SweepParams P, const int COL, const int open) {
    const int L = P.L, Wc = P.Wc;
    const int Kmax = (L + 1) >> 1;
    const long long per_rep = (long long)L * Kmax;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int rep = (int)(gid / per_rep);
    const bool live = rep < P.n_replicas;
    uint32_t a_c = 0, a_c2 = 0, a_cS = 0, a_S = 0;
    bool valid = false;
    if (live) {
        const long long rem = gid - (long long)rep * per_rep;
        const int y = (int)(rem / Kmax);
        const int k = (int)(rem - (long long)y * Kmax);
        const int par = (y + COL) & 1;
        const int x = 2 * k + par;
        if (x < L) {
            valid = true;
            int64_t t = P.t;
            if (P.t_ptr) t += *P.t_ptr;
            const uint32_t k0 = P.key0, k1 = P.key1 ^ (uint32_t)rep;
            const uint32_t ctr2 = (uint32_t)((uint64_t)t & 0xFFFFFFFFu);
            const size_t rep_base = (size_t)rep * (size_t)L * (size_t)Wc;
            uint8_t* own = P.own + rep_base + (size_t)y * Wc + k;
            const uint8_t* oth = P.other + rep_base;
            const uint32_t c = *own;
            const uint32_t S = other_code(oth, L, Wc, x + 1, y, open) + other_code(oth, L, Wc, x - 1, y, open) +
                               other_code(oth, L, Wc, x, y + 1, open) + other_code(oth, L, Wc, x, y - 1, open);
            uint32_t w[4];
            philox4x32_10((uint32_t)(k >> 3), (uint32_t)y, ctr2, ctr3_of(t, STREAM_COARSE, COL), k0, k1, w);
            const int jj = k & 7;
            const uint32_t h = (jj & 1) ? (w[jj >> 1] >> 16) : (w[jj >> 1] & 0xFFFFu);
            const uint32_t flip = h >> 15;
            const uint32_t e = 18u * c + 9u * flip + S;
            const uint32_t E = P.tab16[(size_t)rep * TABLE_ENTRIES + e];
            bool accept = (h & 0x7FFFu) < E;
            if ((h & 0x7FFFu) == E) {
                philox4x32_10((uint32_t)(k >> 3), (uint32_t)y, ctr2, ctr3_of(t, STREAM_FINE, COL), k0, k1, w);
                const uint32_t f16 = (jj & 1) ? (w[jj >> 1] >> 16) : (w[jj >> 1] & 0xFFFFu);
                accept = exact_accept(h, f16, P.tab31[(size_t)rep * TABLE_ENTRIES + e]);
                if (P.ties) atomicAdd(P.ties, 1ull);
            }
            uint32_t cn = c;
            if (accept) {
                cn = c + 1u + flip;  // main.cpp:94 in code space
                if (cn >= 3u) cn -= 3u;
                *own = (uint8_t)cn;
            }
            if (MEASURE) {
                a_c = cn; a_c2 = cn * cn;
                if (COL == 1) { a_cS = cn * S; a_S = S; }
            }
        }
    }
    if (MEASURE && P.obs) {
        // warp-uniform replica -> one atomic per warp; otherwise per thread
        const unsigned live_mask = __ballot_sync(0xFFFFFFFFu, live);
        if (!live) return;
        const int rep_lo = __shfl_sync(live_mask, rep, __ffs(live_mask) - 1);
        const bool uniform = __all_sync(live_mask, rep == rep_lo);
        int64_t slot = P.slot;
        if (P.slot_ptr) slot += *P.slot_ptr;
        unsigned long long* dst = P.obs + ((size_t)slot * P.n_replicas + rep) * OBS_RAW;
        if (uniform) {
            const uint32_t s_c = __reduce_add_sync(live_mask, a_c), s_c2 = __reduce_add_sync(live_mask, a_c2);
            const uint32_t s_cS = __reduce_add_sync(live_mask, a_cS), s_S = __reduce_add_sync(live_mask, a_S);
            if ((int)(threadIdx.x & 31) == __ffs(live_mask) - 1) {
                atomicAdd(dst + 2 * COL + 0, (unsigned long long)s_c);
                atomicAdd(dst + 2 * COL + 1, (unsigned long long)s_c2);
                if (COL == 1) { atomicAdd(dst + 4, (unsigned long long)s_cS); atomicAdd(dst + 5, (unsigned long long)s_S); }
            }
        } else if (valid) {
            atomicAdd(dst + 2 * COL + 0, (unsigned long long)a_c);
            atomicAdd(dst + 2 * COL + 1, (unsigned long long)a_c2);
            if (COL == 1) { atomicAdd(dst + 4, (unsigned long long)a_cS); atomicAdd(dst + 5, (unsigned long long)a_S); }
        }
    }
}

}  // namespace bc
