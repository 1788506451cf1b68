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

// PRMT with the full 4-bit selector nibbles: bit 3 of a nibble replicates the SIGN of the selected
// byte over the output byte (0x00 / 0xFF).  __byte_perm() masks that bit off, hence the PTX.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

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
//   SMALL = false: rows per strip is even (row parity static per unrolled row) and
//                  items/replica is a multiple of 128 (one replica, one table per CTA).
//   SMALL = true : any strip height, up to 4 replicas per CTA (small lattices).
//   VAR          : formulation of the per-site compare:
//                  0 = ISETP/SEL per site; 1 = IMAD difference + PRMT sign masks + SWAR tie test.
// The instruction budget, not HBM, limits this kernel: the ALU pipe
// (LOP3/PRMT/SHF/ISETP) is the scarce one, so VAR 1 moves work to the FMA pipe (IMAD).
// ------------------------------------------------------------------------------------------
constexpr int FAST_THREADS = 128;
constexpr int FAST_MAX_REPL_PER_CTA = FAST_THREADS / 32;
#ifndef BC_FAST_MIN_BLOCKS
#define BC_FAST_MIN_BLOCKS 8
#endif

template <int COL, bool SMALL, bool MEASURE, int VAR>
__global__ void __launch_bounds__(FAST_THREADS, BC_FAST_MIN_BLOCKS)
sweep_fast_kernel(const __grid_constant__ SweepParams P) {
    __shared__ __align__(16) uint32_t s_tab[(SMALL ? FAST_MAX_REPL_PER_CTA : 1) * TAB16_WORDS];

    const uint32_t ipr = (uint32_t)(P.strips * P.cpr);  // items per replica
    const unsigned long long gid0 = (unsigned long long)blockIdx.x * FAST_THREADS;
    const uint32_t rep0 = (uint32_t)(gid0 / ipr);
    for (int i = threadIdx.x; i < (SMALL ? FAST_MAX_REPL_PER_CTA : 1) * TAB16_WORDS; i += FAST_THREADS) {
        const uint32_t r = rep0 + i / TAB16_WORDS;
        s_tab[i] = (r < (uint32_t)P.n_replicas)
                       ? reinterpret_cast<const uint32_t*>(P.tab16)[(size_t)r * TAB16_WORDS + i % TAB16_WORDS]
                       : 0u;
    }
    __syncthreads();

    uint32_t rep, rem;
    if (SMALL) {
        const unsigned long long gid = gid0 + threadIdx.x;
        rep = (uint32_t)(gid / ipr);
        if (rep >= (uint32_t)P.n_replicas) return;  // whole warps only (ipr % 32 == 0)
        rem = (uint32_t)(gid - (unsigned long long)rep * ipr);
    } else {
        rep = rep0;  // ipr % 128 == 0: the CTA lies inside one replica
        rem = (uint32_t)(gid0 - (unsigned long long)rep0 * ipr) + threadIdx.x;
    }
    const uint32_t strip = rem / (uint32_t)P.cpr;
    const uint32_t j = rem - strip * (uint32_t)P.cpr;
    const uint32_t L = (uint32_t)P.L;
    const uint32_t rows = L / (uint32_t)P.strips;
    const uint32_t y0 = strip * rows;
    const uint32_t Wc = (uint32_t)P.Wc;
    const bool open = P.open != 0;

    int64_t t = P.t;
    if (P.t_ptr) t += *P.t_ptr;
    const uint32_t k0 = P.key0, k1 = P.key1 ^ rep;
    const uint32_t ctr2 = (uint32_t)((uint64_t)t & 0xFFFFFFFFu);
    const uint32_t ctr3 = ctr3_of(t, STREAM_COARSE, COL);

    // shared-memory table of this thread's replica: uniform for !SMALL
    const char* tab = reinterpret_cast<const char*>(s_tab) + (SMALL ? (rep - rep0) * (TAB16_WORDS * 4) : 0u);
    const size_t rep_base = (size_t)rep * (size_t)L * (size_t)Wc;
    uint8_t* own_col = P.own + rep_base + (size_t)j * 16;            // + y*Wc
    const uint8_t* oth_col = P.other + rep_base + (size_t)j * 16;    // + y*Wc
    const uint32_t jr = (j + 1 == (uint32_t)P.cpr) ? 0u : j + 1;     // chunk holding k+1 of our last site
    const uint32_t jl = (j == 0) ? (uint32_t)P.cpr - 1 : j - 1;      // chunk holding k-1 of our first site
    const uint8_t* side_r = P.other + rep_base + (size_t)jr * 16;       // first word of the right chunk
    const uint8_t* side_l = P.other + rep_base + (size_t)jl * 16 + 12;  // last word of the left chunk
    const bool edge_r = open && (j + 1 == (uint32_t)P.cpr), edge_l = open && (j == 0);

    const uint4 ones4 = make_uint4(ONES, ONES, ONES, ONES);
    const uint32_t total = L * Wc;  // bytes of one replica's colour array (< 2^32: L <= 65536)

    uint32_t a_c = 0, a_c2 = 0, a_cS = 0, a_S = 0, n_ties = 0;

    // rows y0-1 and y0 of the other colour (periodic wrap or open boundary = all-zero-spin row)
    uint32_t off = y0 * Wc;  // byte offset of row y inside the replica's colour array
    uint4 upv, midv;
    if (y0 == 0) upv = open ? ones4 : ldg16(oth_col + (total - Wc));
    else upv = ldg16(oth_col + (off - Wc));
    midv = ldg16(oth_col + off);

    // one row of the strip; PARV is the row parity (y+COL)&1, or -1 for "compute it"
    auto do_row = [&](const uint32_t y, auto parc) {
        constexpr int PARV = decltype(parc)::value;
        const int par = (PARV < 0) ? (int)((y + COL) & 1u) : PARV;
        uint32_t off_dn = off + Wc;
        uint4 dnv;
        if (off_dn == total) {
            dnv = open ? ones4 : ldg16(oth_col);
        } else {
            dnv = ldg16(oth_col + off_dn);
        }
        const uint32_t side = par ? (edge_r ? 0x00000001u : ldg4(side_r + off))
                                  : (edge_l ? 0x01000000u : ldg4(side_l + off));
        uint8_t* own_ptr = own_col + off;
        const uint4 ownv = *reinterpret_cast<const uint4*>(own_ptr);

        // coarse randoms: one Philox call per 8 sites; 16 bits per site
        uint32_t q[8];
        {
            uint32_t qa[4], qb[4];
            philox4x32_10(2u * j, y, ctr2, ctr3, k0, k1, qa);
            philox4x32_10(2u * j + 1u, y, ctr2, ctr3, k0, k1, qb);
#pragma unroll
            for (int i = 0; i < 4; ++i) { q[i] = qa[i]; q[4 + i] = qb[i]; }
        }

        const uint32_t o[4] = {ownv.x, ownv.y, ownv.z, ownv.w};
        const uint32_t u[4] = {upv.x, upv.y, upv.z, upv.w};
        const uint32_t m[4] = {midv.x, midv.y, midv.z, midv.w};
        const uint32_t d[4] = {dnv.x, dnv.y, dnv.z, dnv.w};
        uint32_t idx[4], nw[4], S[4];
        uint32_t xr[4], msk[4];  // proposal xor, accept byte masks (0xFF per accepted site)
        bool tie = false;
        uint32_t tieacc = 0;
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            // horizontal neighbour that is not at the same k: k+1 for par = 1, k-1 for par = 0
            const uint32_t sr = __funnelshift_r(m[wi], wi < 3 ? m[wi + 1] : side, 8);
            const uint32_t sl = __funnelshift_l(wi > 0 ? m[wi - 1] : side, m[wi], 8);
            const uint32_t sh = par ? sr : sl;
            const uint32_t S4 = u[wi] + d[wi] + m[wi] + sh;  // bytes 0..8, no carries (main.cpp:97-103)
            const uint32_t ra = q[2 * wi], rb = q[2 * wi + 1];
            const uint32_t c4 = o[wi];
            if (VAR == 0) {
                // flip bits = bit 15 of each halfword -> one per byte
                const uint32_t f4 = (__byte_perm(ra, rb, 0x7531) >> 7) & ONES;
                // byte offsets into the u16 table: 2*(18c + 9f + S) <= 106
                const uint32_t idx4 = c4 * 36u + f4 * 18u + S4 * 2u;
                // proposal (main.cpp:94 in code space): p = (c + 1 + f) mod 3
                const uint32_t t4 = c4 + f4 + ONES;                    // 1..4
                const uint32_t ge = ((t4 + 0x05050505u) >> 3) & ONES;  // t >= 3
                const uint32_t p4 = t4 - 3u * ge;
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
                msk[wi] = m4 * 0xFFu;
            } else {
                // flips as byte masks (PRMT sign replication of bytes 1,3 of ra and rb), then 0/1 bytes
                const uint32_t fFF = prmt(ra, rb, 0xFDB9u);
                const uint32_t f4 = fFF & ONES;
                const uint32_t idx4 = c4 * 36u + (f4 * 18u + (S4 + S4));
                // proposal xor x = c ^ p = ((2c + f) mod 3) + 1   (main.cpp:94 in code space)
                const uint32_t q1 = c4 * 2u + (f4 + ONES);                 // 2c + f + 1 in 1..6
                const uint32_t ge = ((q1 + 0x04040404u) >> 3) & ONES;      // 2c + f >= 3
                xr[wi] = q1 - 3u * ge;
                const uint32_t a0 = idx4 & 0xFFu;
                const uint32_t a1 = __byte_perm(idx4, 0u, 0x4441);
                const uint32_t a2 = __byte_perm(idx4, 0u, 0x4442);
                const uint32_t a3 = idx4 >> 24;
                const uint32_t E0 = *reinterpret_cast<const uint16_t*>(tab + a0);
                const uint32_t E1 = *reinterpret_cast<const uint16_t*>(tab + a1);
                const uint32_t E2 = *reinterpret_cast<const uint16_t*>(tab + a2);
                const uint32_t E3 = *reinterpret_cast<const uint16_t*>(tab + a3);
                // d = (h - E) * 65536 + (noise < 65536): sign(d) <=> h < E, top half zero <=> h == E.
                // h - E lies in [-32768, 32767], so the 32-bit difference cannot wrap.
                const uint32_t d0 = (ra << 16) - E0 * 65536u;
                const uint32_t d1 = ra - E1 * 65536u;
                const uint32_t d2 = (rb << 16) - E2 * 65536u;
                const uint32_t d3 = rb - E3 * 65536u;
                const uint32_t H01 = __byte_perm(d0, d1, 0x7632);  // [d0.hi16, d1.hi16]
                const uint32_t H23 = __byte_perm(d2, d3, 0x7632);
                msk[wi] = prmt(H01, H23, 0xFDB9u);            // 0xFF where d < 0
                tieacc |= (H01 - 0x00010001u) & ~H01;               // has-zero-halfword test (bit 15 / 31)
                tieacc |= (H23 - 0x00010001u) & ~H23;
                idx[wi] = idx4;
            }
            S[wi] = S4;
        }
        if (VAR != 0) tie = (tieacc & 0x80008000u) != 0u;

        if (tie) {
            // rare (about 2^-15 per site): redo the tied sites exactly with the fine stream
            uint32_t fa[4], fb[4];
            const uint32_t ctr3f = ctr3_of(t, STREAM_FINE, COL);
            philox4x32_10(2u * j, y, ctr2, ctr3f, k0, k1, fa);
            philox4x32_10(2u * j + 1u, y, ctr2, ctr3f, k0, k1, fb);
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
                    if (exact_accept(h, f16, __ldg(t31 + (a >> 1)))) msk[wi] |= (0xFFu << (8 * b));
                    ++n_ties;
                }
            }
        }

#pragma unroll
        for (int wi = 0; wi < 4; ++wi) nw[wi] = o[wi] ^ (xr[wi] & msk[wi]);
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
        off = off_dn;
    };
    if (SMALL) {
        for (uint32_t r = 0; r < rows; ++r) do_row(y0 + r, std::integral_constant<int, -1>{});
    } else {
        // y0 and rows are even: parities alternate COL, 1-COL statically
        for (uint32_t r = 0; r < rows; r += 2) {
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
