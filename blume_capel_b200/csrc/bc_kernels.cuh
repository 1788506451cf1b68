// bc_kernels.cuh -- hand-written sm_100a kernels of the checkerboard Metropolis half-sweep.
//
// Data layout in HBM: two compact per-colour arrays of spin codes c = s+1 in {0,1,2}, one
// byte per site, [replica][y][k] with k = x>>1 (site x = 2k + ((y+colour)&1)), row stride Wc.
// One 128-bit load = 16 same-colour sites of one row.  For a site of colour C in row y with
// p = (y+C)&1 its neighbours in the other colour's array O are O[y-1][k], O[y+1][k], O[y][k]
// and O[y][k+1] (p = 1) or O[y][k-1] (p = 0).
//
// Per site the kernels do what /root/reference/main.cpp:89-112 does per attempt, with the
// double-precision exp() replaced by the integer table of the spec.
#pragma once
#include "bc_common.cuh"
#include <type_traits>

namespace bc {

__device__ __forceinline__ uint4 ldg16(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint32_t ldg4(const uint8_t* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(0xFFFFFFFFu, v); }

// Exact rule for a site whose coarse compare tied: u31 = (u15 << 16 | f16) < thr31.
__device__ __forceinline__ bool exact_accept(uint32_t h16, uint32_t f16, uint32_t thr31) {
    return (((h16 & 0x7FFFu) << 16) | f16) < thr31;
}

// ------------------------------------------------------------------------------------------
// Fast kernel: L % 32 == 0.  One thread = one 16-byte chunk column j, marching down the
// L/strips rows of its strip with the three rows of the other colour held in registers.
// Work item id = blockIdx.x*128 + threadIdx.x -> (replica, strip, j), j fastest, so a warp
// reads 512 contiguous bytes per row; items/replica is a multiple of 32 (host guarantees),
// so a warp never spans replicas.
//   DYNPAR = false: rows per strip is even; the row parity is static per unrolled row.
//   DYNPAR = true : any strip height (used with 1 row per thread for small lattices).
// ------------------------------------------------------------------------------------------
constexpr int FAST_THREADS = 128;
constexpr int FAST_MAX_REPL_PER_CTA = FAST_THREADS / 32;

template <int COL, bool DYNPAR, bool MEASURE, bool OPEN>
__global__ void __launch_bounds__(FAST_THREADS)
sweep_fast_kernel(const __grid_constant__ SweepParams P) {
    __shared__ __align__(16) uint32_t s_tab[FAST_MAX_REPL_PER_CTA * TAB16_WORDS];

    const int ipr = P.strips * P.cpr;  // items per replica
    const long long gid0 = (long long)blockIdx.x * FAST_THREADS;
    const int rep0 = (int)(gid0 / ipr);
    for (int i = threadIdx.x; i < FAST_MAX_REPL_PER_CTA * TAB16_WORDS; i += FAST_THREADS) {
        const int r = rep0 + i / TAB16_WORDS;
        s_tab[i] = (r < P.n_replicas)
                       ? reinterpret_cast<const uint32_t*>(P.tab16)[(size_t)r * TAB16_WORDS + i % TAB16_WORDS]
                       : 0u;
    }
    __syncthreads();

    const long long gid = gid0 + threadIdx.x;
    const int rep = (int)(gid / ipr);
    if (rep >= P.n_replicas) return;  // whole warps only (ipr % 32 == 0)
    const int rem = (int)(gid - (long long)rep * ipr);
    const int strip = rem / P.cpr;
    const int j = rem - strip * P.cpr;
    const int L = P.L;
    const int rows = L / P.strips;
    const int y0 = strip * rows;
    const int Wc = P.Wc;

    int64_t t = P.t;
    if (P.t_ptr) t += *P.t_ptr;
    const uint32_t k0 = P.key0, k1 = P.key1 ^ (uint32_t)rep;
    const uint32_t ctr2 = (uint32_t)((uint64_t)t & 0xFFFFFFFFu);
    const uint32_t ctr3 = ctr3_of(t, STREAM_COARSE, COL);

    const char* tab = reinterpret_cast<const char*>(s_tab) + (rep - rep0) * (TAB16_WORDS * 4);
    const size_t rep_base = (size_t)rep * (size_t)L * (size_t)Wc;
    uint8_t* own_base = P.own + rep_base + (size_t)j * 16;
    const uint8_t* oth_base = P.other + rep_base + (size_t)j * 16;
    const uint8_t* oth_rep = P.other + rep_base;
    const int jr = (j + 1 == P.cpr) ? 0 : j + 1;   // chunk holding k+1 of our last site
    const int jl = (j == 0) ? P.cpr - 1 : j - 1;   // chunk holding k-1 of our first site
    const int off_r = jr * 16, off_l = jl * 16 + 12;
    const bool edge_r = OPEN && (j + 1 == P.cpr), edge_l = OPEN && (j == 0);

    const uint4 ones4 = make_uint4(ONES, ONES, ONES, ONES);
    auto load_other = [&](int y) -> uint4 {
        if (y < 0) { if (OPEN) return ones4; y += L; }
        if (y >= L) { if (OPEN) return ones4; y -= L; }
        return ldg16(oth_base + (size_t)y * Wc);
    };

    uint32_t a_c = 0, a_c2 = 0, a_cS = 0, a_S = 0, n_ties = 0;

    uint4 upv = load_other(y0 - 1);
    uint4 midv = load_other(y0);
    // one row of the strip; PARC::value is the row parity (y+COL)&1, or -1 for "compute it"
    auto do_row = [&](const int y, auto parc) {
        constexpr int PARV = decltype(parc)::value;
        const int par = (PARV < 0) ? ((y + COL) & 1) : PARV;
        const uint4 dnv = load_other(y + 1);
        const uint32_t side = par ? (edge_r ? 0x00000001u : ldg4(oth_rep + (size_t)y * Wc + off_r))
                                  : (edge_l ? 0x01000000u : ldg4(oth_rep + (size_t)y * Wc + off_l));
        uint8_t* own_ptr = own_base + (size_t)y * Wc;
        const uint4 ownv = *reinterpret_cast<const uint4*>(own_ptr);

        // coarse randoms: one Philox call per 8 sites; 16 bits per site
        uint32_t q[8];
        {
            uint32_t qa[4], qb[4];
            philox4x32_10(2u * (uint32_t)j, (uint32_t)y, ctr2, ctr3, k0, k1, qa);
            philox4x32_10(2u * (uint32_t)j + 1u, (uint32_t)y, ctr2, ctr3, k0, k1, qb);
#pragma unroll
            for (int i = 0; i < 4; ++i) { q[i] = qa[i]; q[4 + i] = qb[i]; }
        }

        const uint32_t o[4] = {ownv.x, ownv.y, ownv.z, ownv.w};
        const uint32_t u[4] = {upv.x, upv.y, upv.z, upv.w};
        const uint32_t m[4] = {midv.x, midv.y, midv.z, midv.w};
        const uint32_t d[4] = {dnv.x, dnv.y, dnv.z, dnv.w};
        uint32_t idx[4], xr[4], msk[4], S[4];
        bool tie = false;
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            // horizontal neighbour that is not at the same k: k+1 for par = 1, k-1 for par = 0
            const uint32_t sr = __funnelshift_r(m[wi], wi < 3 ? m[wi + 1] : side, 8);
            const uint32_t sl = __funnelshift_l(wi > 0 ? m[wi - 1] : side, m[wi], 8);
            const uint32_t sh = par ? sr : sl;
            const uint32_t S4 = u[wi] + d[wi] + m[wi] + sh;  // bytes 0..8, no carries (main.cpp:97-103)
            const uint32_t ra = q[2 * wi], rb = q[2 * wi + 1];
            // flip bits = bit 15 of each halfword -> one per byte
            const uint32_t f4 = (__byte_perm(ra, rb, 0x7531) >> 7) & ONES;
            const uint32_t c4 = o[wi];
            // byte offsets into the u16 table: 2*(18c + 9f + S) <= 106
            const uint32_t idx4 = c4 * 36u + f4 * 18u + S4 * 2u;
            // proposal (main.cpp:94 in code space): p = (c + 1 + f) mod 3
            const uint32_t t4 = c4 + f4 + ONES;                    // 1..4
            const uint32_t ge = ((t4 + 0x05050505u) >> 3) & ONES;  // t >= 3
            const uint32_t p4 = t4 - 3u * ge;
            // per-site compare with the coarse threshold (flip sits in bit 15 of both sides)
            const uint32_t a0 = idx4 & 0xFFu;
            const uint32_t a1 = __byte_perm(idx4, 0u, 0x4441);
            const uint32_t a2 = __byte_perm(idx4, 0u, 0x4442);
            const uint32_t a3 = idx4 >> 24;
            const uint32_t E0 = *reinterpret_cast<const uint16_t*>(tab + a0);
            const uint32_t E1 = *reinterpret_cast<const uint16_t*>(tab + a1);
            const uint32_t E2 = *reinterpret_cast<const uint16_t*>(tab + a2);
            const uint32_t E3 = *reinterpret_cast<const uint16_t*>(tab + a3);
            const uint32_t h0 = ra & 0xFFFFu, h1 = ra >> 16, h2 = rb & 0xFFFFu, h3 = rb >> 16;
            uint32_t m4 = 0;
            m4 += (h0 < E0) ? 0x00000001u : 0u;
            m4 += (h1 < E1) ? 0x00000100u : 0u;
            m4 += (h2 < E2) ? 0x00010000u : 0u;
            m4 += (h3 < E3) ? 0x01000000u : 0u;
            tie = tie || (h0 == E0) || (h1 == E1) || (h2 == E2) || (h3 == E3);
            idx[wi] = idx4;
            xr[wi] = c4 ^ p4;
            msk[wi] = m4;
            S[wi] = S4;
        }

        if (tie) {
            // rare (about 2^-15 per site): redo the tied sites exactly with the fine stream
            uint32_t fa[4], fb[4];
            const uint32_t ctr3f = ctr3_of(t, STREAM_FINE, COL);
            philox4x32_10(2u * (uint32_t)j, (uint32_t)y, ctr2, ctr3f, k0, k1, fa);
            philox4x32_10(2u * (uint32_t)j + 1u, (uint32_t)y, ctr2, ctr3f, k0, k1, fb);
            const uint32_t* t31 = P.tab31 + (size_t)rep * TABLE_ENTRIES;
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                const int wi = s >> 2, b = s & 3;
                const uint32_t a = (idx[wi] >> (8 * b)) & 0xFFu;
                const uint32_t rw = q[2 * wi + (b >> 1)];
                const uint32_t h = (b & 1) ? (rw >> 16) : (rw & 0xFFFFu);
                const uint32_t E = *reinterpret_cast<const uint16_t*>(tab + a);
                if (h == E) {
                    const uint32_t fw = (s < 8) ? fa[(s & 7) >> 1] : fb[(s & 7) >> 1];
                    const uint32_t f16 = (s & 1) ? (fw >> 16) : (fw & 0xFFFFu);
                    if (exact_accept(h, f16, __ldg(t31 + (a >> 1)))) msk[wi] |= (1u << (8 * b));
                    ++n_ties;
                }
            }
        }

        uint32_t nw[4];
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) nw[wi] = o[wi] ^ (xr[wi] & (msk[wi] * 0xFFu));
        *reinterpret_cast<uint4*>(own_ptr) = make_uint4(nw[0], nw[1], nw[2], nw[3]);

        if (MEASURE) {
#pragma unroll
            for (int wi = 0; wi < 4; ++wi) {
                a_c = __dp4a(nw[wi], ONES, a_c);
                a_c2 = __dp4a(nw[wi], nw[wi], a_c2);
                if (COL == 1) {
                    a_cS = __dp4a(nw[wi], S[wi], a_cS);
                    a_S = __dp4a(S[wi], ONES, a_S);
                }
            }
        }
        upv = midv;
        midv = dnv;
    };
    if (DYNPAR) {
        for (int r = 0; r < rows; ++r) do_row(y0 + r, std::integral_constant<int, -1>{});
    } else {
        // y0 and rows are even: parities alternate COL, 1-COL statically
        for (int r = 0; r < rows; r += 2) {
            do_row(y0 + r, std::integral_constant<int, COL>{});
            do_row(y0 + r + 1, std::integral_constant<int, 1 - COL>{});
        }
    }

    if (MEASURE && P.obs) {
        int64_t slot = P.slot;
        if (P.slot_ptr) slot += *P.slot_ptr;
        unsigned long long* dst = P.obs + ((size_t)slot * P.n_replicas + rep) * OBS_RAW;
        const uint32_t s_c = warp_sum(a_c), s_c2 = warp_sum(a_c2);
        const uint32_t s_cS = (COL == 1) ? warp_sum(a_cS) : 0u, s_S = (COL == 1) ? warp_sum(a_S) : 0u;
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(dst + 2 * COL + 0, (unsigned long long)s_c);
            atomicAdd(dst + 2 * COL + 1, (unsigned long long)s_c2);
            if (COL == 1) {
                atomicAdd(dst + 4, (unsigned long long)s_cS);
                atomicAdd(dst + 5, (unsigned long long)s_S);
            }
        }
    }
    if (n_ties && P.ties) atomicAdd(P.ties, (unsigned long long)n_ties);
}

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
            bool accept = h < E;
            if (h == E) {
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

// ------------------------------------------------------------------------------------------
// Layout conversion, init, standalone measurement
// ------------------------------------------------------------------------------------------
// row-major int8 spins [rep][y][x]  ->  compact codes (both colours); pad bytes = code 1 (spin 0)
__global__ void pack_kernel(const int8_t* __restrict__ spins, uint8_t* __restrict__ c0, uint8_t* __restrict__ c1,
                            int L, int Wc, int rep_first, int n_rep) {
    const long long per_rep = (long long)L * Wc;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const long long rem = gid - (long long)r * per_rep;
    const int y = (int)(rem / Wc), k = (int)(rem - (long long)y * Wc);
    const size_t dst = (size_t)(rep_first + r) * per_rep + (size_t)y * Wc + k;
    const int8_t* src = spins + (size_t)r * L * L + (size_t)y * L;
    const int xa = 2 * k + (y & 1), xb = 2 * k + ((y + 1) & 1);
    c0[dst] = (xa < L) ? (uint8_t)(src[xa] + 1) : (uint8_t)1;
    c1[dst] = (xb < L) ? (uint8_t)(src[xb] + 1) : (uint8_t)1;
}

__global__ void unpack_kernel(int8_t* __restrict__ spins, const uint8_t* __restrict__ c0,
                              const uint8_t* __restrict__ c1, int L, int Wc, int rep_first, int n_rep) {
    const long long per_rep = (long long)L * L;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const long long rem = gid - (long long)r * per_rep;
    const int y = (int)(rem / L), x = (int)(rem - (long long)y * L);
    const size_t src = (size_t)(rep_first + r) * L * Wc + (size_t)y * Wc + (x >> 1);
    const uint8_t code = ((x + y) & 1) ? c1[src] : c0[src];
    spins[gid] = (int8_t)((int)code - 1);
}

// iid uniform {-1,0,1} (main.cpp:166-171) from Philox stream INIT, indexed by the site index x + y*L
__global__ void init_kernel(uint8_t* __restrict__ c0, uint8_t* __restrict__ c1, int L, int Wc, int rep_first,
                            int n_rep, uint32_t key0, uint32_t key1) {
    const long long per_rep = (long long)L * L;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const int rep = rep_first + r;
    const long long i = gid - (long long)r * per_rep;
    const int y = (int)(i / L), x = (int)(i - (long long)y * L);
    uint32_t w[4];
    philox4x32_10((uint32_t)(i >> 2), (uint32_t)((unsigned long long)i >> 34), 0u, STREAM_INIT << 1, key0,
                  key1 ^ (uint32_t)rep, w);
    const uint32_t code = (uint32_t)(((unsigned long long)w[i & 3] * 3ull) >> 32);
    const size_t dst = (size_t)rep * L * Wc + (size_t)y * Wc + (x >> 1);
    if ((x + y) & 1) c1[dst] = (uint8_t)code; else c0[dst] = (uint8_t)code;
}

__global__ void fill_kernel(uint8_t* __restrict__ p, size_t n, uint8_t v) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) p[gid] = v;
}

// standalone observables of the current configuration into raw accumulators [rep][OBS_RAW]
__global__ void measure_kernel(const uint8_t* __restrict__ c0, const uint8_t* __restrict__ c1, int L, int Wc,
                               int n_rep, int open, unsigned long long* __restrict__ obs) {
    const long long per_rep = (long long)L * L;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int rep = (int)(gid / per_rep);
    const long long i = gid - (long long)rep * per_rep;
    const int y = (int)(i / L), x = (int)(i - (long long)y * L);
    const size_t rb = (size_t)rep * L * Wc;
    const int col = (x + y) & 1;
    const uint32_t c = (col ? c1 : c0)[rb + (size_t)y * Wc + (x >> 1)];
    uint32_t v[OBS_RAW] = {0, 0, 0, 0, 0, 0};
    v[2 * col + 0] = c;
    v[2 * col + 1] = c * c;
    if (col == 1) {
        const uint8_t* oth = c0 + rb;
        const uint32_t S = other_code(oth, L, Wc, x + 1, y, open) + other_code(oth, L, Wc, x - 1, y, open) +
                           other_code(oth, L, Wc, x, y + 1, open) + other_code(oth, L, Wc, x, y - 1, open);
        v[4] = c * S;
        v[5] = S;
    }
    const unsigned mask = __activemask();
    const int leader = __ffs(mask) - 1;
    const int rep_lo = __shfl_sync(mask, rep, leader);
    const bool uniform = __all_sync(mask, rep == rep_lo);
    unsigned long long* dst = obs + (size_t)rep * OBS_RAW;
#pragma unroll
    for (int i = 0; i < OBS_RAW; ++i) {
        if (uniform) {
            const uint32_t sum = __reduce_add_sync(mask, v[i]);
            if ((int)(threadIdx.x & 31) == leader && sum) atomicAdd(dst + i, (unsigned long long)sum);
        } else if (v[i]) {
            atomicAdd(dst + i, (unsigned long long)v[i]);
        }
    }
}

}  // namespace bc
