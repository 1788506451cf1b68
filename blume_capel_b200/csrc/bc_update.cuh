// bc_update.cuh -- the per-row update (16 same-colour sites of a row) shared by the streaming, group, persistent, wavefront and resident kernels.
// Part of the sm_100a Blume-Capel engine; included through bc_kernels.cuh.
#pragma once
#include "bc_common.cuh"

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

// 32-bit multiply-add kept as one IMAD (FMA pipe); plain C++ lets the compiler strength-reduce it
// into ALU-pipe adds/shifts, and the ALU pipe is the bottleneck of the sweep kernel.
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// cp.async (LDGSTS): global -> shared without staging registers; completion per thread via groups
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// (a & b) ^ c as ONE LOP3 (immLut 0x6A); c must live in a register, a LOP3 has room for one immediate only
__device__ __forceinline__ uint32_t and_xor(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// packed signed 16-bit minimum (VIMNMX.S16x2): 0x8000 is the most negative lane value
__device__ __forceinline__ uint32_t min_s16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("min.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// Exact rule for a site whose coarse compare tied: u31 = (u15 << 16 | f16) < thr31.
__device__ __forceinline__ bool exact_accept(uint32_t h16, uint32_t f16, uint32_t thr31) {
    return (((h16 & 0x7FFFu) << 16) | f16) < thr31;
}

// ------------------------------------------------------------------------------------------
// The per-row update shared by the streaming kernel and the resident (small-lattice) kernel:
// 16 same-colour sites of row y, chunk j.  Inputs: the other colour's rows y-1, y, y+1 (upv, midv,
// dnv), the word holding the horizontal neighbour that lives in the adjacent chunk (`side`), the
// own chunk; par = (y+COL)&1; tabm = shared-memory coarse table biased by -18 bytes; t31_rep = the
// replica's exact thresholds.  Returns the new chunk; accumulates the raw observables.
// ------------------------------------------------------------------------------------------
struct RowCtx {
    uint32_t j, ctr2, ctr3, k0, k1;
    int64_t t;
    const char* tabm;
    const uint32_t* t31_rep;
    uint32_t tab2s;  // PAIR: shared-space address of the pair table, biased by -TAB2_BIAS bytes
};

// Half of a chunk: the 8 sites served by one Philox call (HALF = 0: words 0,1; HALF = 1: words 2,3).
template <int COL, bool MEASURE, int HALF, bool PAIR>
__device__ __forceinline__ void update_half8(const RowCtx& C, const uint32_t y, const int par, const uint32_t (&o)[4],
                                             const uint32_t (&u)[4], const uint32_t (&m)[4], const uint32_t (&d)[4],
                                             const uint32_t side, const uint32_t ones_r, const uint32_t (&q)[4],
                                             uint32_t (&nw)[4], uint32_t& a_c, uint32_t& a_c2, uint32_t& a_cS,
                                             uint32_t& a_S, uint32_t& n_ties) {
    const uint32_t j = C.j, ctr2 = C.ctr2, k0 = C.k0, k1 = C.k1;
    const int64_t t = C.t;
    const char* tabm = C.tabm;
    const uint32_t* t31_rep = C.t31_rep;
    constexpr int half = HALF;
    // q = the coarse randoms of this half: one Philox call per 8 sites, 16 bits per site (coarse_draw)
    uint32_t idx[2], xr[2], msk[2] /* reject byte masks */, Ssave[2];
    uint32_t tiemin = 0x7FFF7FFFu;
#pragma unroll
    for (int w2 = 0; w2 < 2; ++w2) {
        const int wi = 2 * half + w2;
        // horizontal neighbour that is not at the same k: k+1 for par = 1, k-1 for par = 0
        const uint32_t sr = __funnelshift_r(m[wi], wi < 3 ? m[wi + 1] : side, 8);
        const uint32_t sl = __funnelshift_l(wi > 0 ? m[wi - 1] : side, m[wi], 8);
        const uint32_t sh = par ? sr : sl;
        const uint32_t S4 = u[wi] + d[wi] + m[wi] + sh;  // bytes 0..8, no carries (main.cpp:97-103)
        const uint32_t ra = q[2 * w2], rb = q[2 * w2 + 1];
        const uint32_t c4 = o[wi];
        // flips (bit 15 of each halfword) as byte masks via PRMT sign replication; fO = f + 1
        const uint32_t fO = and_xor(prmt(ra, rb, 0xFDB9u), 0x03030303u, ones_r);  // one LOP3
        // table entry numbers (biased by +9 through f+1): 18c + 9(f+1) + S <= 62; the factor 2 of the
        // u16 byte offset is applied by the IDP.4A that extracts each byte
        const uint32_t idx4 = imad(c4, 18u, imad(fO, 9u, S4));
        // proposal xor x = c ^ p = ((2c + f) mod 3) + 1   (main.cpp:94 in code space)
        const uint32_t q1 = imad(c4, 2u, fO);            // 2c + f + 1 in 1..6
        // (q1 + [q1 >= 4]) & 3: the shift drags the low bits of the next byte into bits 6..7, the byte sums stay
        // below 256 (no carry) and the mask drops them.  Shift + add fuse into one LEA.HI: two ALU-pipe
        // instructions, none on the FMA-heavy pipe that Philox saturates.
        xr[w2] = (q1 + (q1 >> 2)) & 0x03030303u;
        uint32_t Z01, Z23;
        if (PAIR) {
            // Pair table: ONE lookup per two sites.  Byte offset 4*(b0-9) + 216*(b1-9) by one IDP.4A (the
            // bias and the shared-space base are its accumulator), LDS.32 returns the packed word T with
            // ra - T = lanes 0x8000 + u15 - E15 exactly: the flip bits of ra (bits 15, 31) are known to the
            // table through the entry numbers and cancel, so no guard-bit OR and no PRMT to pack two
            // thresholds (see pair_table_kernel).
            const uint32_t a01 = __dp4a(idx4, 0x0000D804u, C.tab2s);
            const uint32_t a23 = __dp4a(idx4, 0xD8040000u, C.tab2s);
            uint32_t T01, T23;
            asm("ld.shared.u32 %0, [%1];" : "=r"(T01) : "r"(a01));
            asm("ld.shared.u32 %0, [%1];" : "=r"(T23) : "r"(a23));
            Z01 = ra - T01;
            Z23 = rb - T23;
        } else {
            // per-site table offsets: byte i of idx4, extracted on the FMA pipe (IDP.4A)
            const uint32_t a0 = __dp4a(idx4, 0x00000002u, 0u);
            const uint32_t a3 = __dp4a(idx4, 0x02000000u, 0u);
            const uint32_t a1 = __dp4a(idx4, 0x00000200u, 0u);
            const uint32_t a2 = __dp4a(idx4, 0x00020000u, 0u);
            const uint32_t E0 = *reinterpret_cast<const uint16_t*>(tabm + a0);
            const uint32_t E1 = *reinterpret_cast<const uint16_t*>(tabm + a1);
            const uint32_t E2 = *reinterpret_cast<const uint16_t*>(tabm + a2);
            const uint32_t E3 = *reinterpret_cast<const uint16_t*>(tabm + a3);
            // Packed 2-site compare: both 15-bit draws of a word get a guard bit, Z lane = 0x8000 + u15 - E15
            // in [0, 0xFFFF] (no cross-lane borrow): bit 15 clear <=> u15 < E15 (accept), lane == 0x8000 <=>
            // tie.  E15 = thr31 >> 16 (32768 = always accept: lane = u15, never a tie).
            Z01 = (ra | 0x80008000u) - __byte_perm(E0, E1, 0x5410);
            Z23 = (rb | 0x80008000u) - __byte_perm(E2, E3, 0x5410);
        }
        msk[w2] = prmt(Z01, Z23, 0xFDB9u);                  // 0xFF where u15 >= E15 (REJECT mask)
        // tie <=> some lane == 0x8000, the most negative signed 16-bit value: one packed min per word pair
        tiemin = min_s16x2(tiemin, min_s16x2(Z01, Z23));
        idx[w2] = idx4;
        Ssave[w2] = S4;
    }
    // high lane == 0x8000 <=> the word, as a signed number, is below 0x80010000; low lane: shifted to the top
    if ((int32_t)tiemin < (int32_t)0x80010000u || (tiemin << 16) == 0x80000000u) {
        // rare (about 2^-15 per uphill site): redo the tied sites exactly with the fine stream
        uint32_t f[4];
        philox4x32_10(2u * j + (uint32_t)half, y, ctr2, ctr3_of(t, STREAM_FINE, COL), k0, k1, f);
        const uint32_t* t31 = t31_rep;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int w2 = s >> 2, b = s & 3;
            const uint32_t a = ((idx[w2] >> (8 * b)) & 0xFFu) * 2u;
            const uint32_t rw = q[s >> 1];
            const uint32_t h = (s & 1) ? (rw >> 16) : (rw & 0xFFFFu);
            const uint32_t E = *reinterpret_cast<const uint16_t*>(tabm + a);
            if ((h & 0x7FFFu) == E) {
                const uint32_t fw = f[s >> 1];
                const uint32_t f16 = (s & 1) ? (fw >> 16) : (fw & 0xFFFFu);
                if (exact_accept(h, f16, __ldg(t31 + ((a - 18u) >> 1)))) msk[w2] &= ~(0xFFu << (8 * b));
                ++n_ties;
            }
        }
    }
#pragma unroll
    for (int w2 = 0; w2 < 2; ++w2) {
        const int wi = 2 * half + w2;
        const uint32_t v = o[wi] ^ (xr[w2] & ~msk[w2]);
        if (MEASURE) {
            a_c = __dp4a(v, ONES, a_c);
            a_c2 = __dp4a(v, v, a_c2);
            if (COL == 1) {
                a_cS = __dp4a(v, Ssave[w2], a_cS);
                a_S = __dp4a(Ssave[w2], ONES, a_S);
            }
        }
        nw[wi] = v;
    }
}

// coarse draws of half `half` (0/1) of chunk C.j in row y
__device__ __forceinline__ void coarse_draw(const RowCtx& C, const uint32_t y, const uint32_t half, uint32_t (&q)[4]) {
    philox4x32_10(2u * C.j + half, y, C.ctr2, C.ctr3, C.k0, C.k1, q);
}

template <int COL, bool MEASURE, bool PAIR = false>
__device__ __forceinline__ uint4 update_chunk16(const RowCtx& C, const uint32_t y, const int par, const uint4& upv,
                                                const uint4& midv, const uint4& dnv, const uint32_t side,
                                                const uint4& ownv, uint32_t& a_c, uint32_t& a_c2, uint32_t& a_cS,
                                                uint32_t& a_S, uint32_t& n_ties) {
    uint32_t ones_r;  // ONES held in a register the compiler cannot fold back into an immediate
    asm volatile("mov.u32 %0, 0x01010101;" : "=r"(ones_r));
    const uint32_t o[4] = {ownv.x, ownv.y, ownv.z, ownv.w};
    const uint32_t u[4] = {upv.x, upv.y, upv.z, upv.w};
    const uint32_t m[4] = {midv.x, midv.y, midv.z, midv.w};
    const uint32_t d[4] = {dnv.x, dnv.y, dnv.z, dnv.w};
    uint32_t nw[4], q[4];
    coarse_draw(C, y, 0u, q);
    update_half8<COL, MEASURE, 0, PAIR>(C, y, par, o, u, m, d, side, ones_r, q, nw, a_c, a_c2, a_cS, a_S, n_ties);
    coarse_draw(C, y, 1u, q);
    update_half8<COL, MEASURE, 1, PAIR>(C, y, par, o, u, m, d, side, ones_r, q, nw, a_c, a_c2, a_cS, a_S, n_ties);
    return make_uint4(nw[0], nw[1], nw[2], nw[3]);
}

// ---------------------------------------------------------------------------------------------------------
// Branch-free forms for the streaming kernel's pair-table instantiations (sweep_fast_body, tie repair out of line).
// ---------------------------------------------------------------------------------------------------------
// Word wi of a chunk (4 sites): neighbour sums, table entry numbers and proposal XORs on packed bytes.
__device__ __forceinline__ void site_word(const int wi, const int par, const uint32_t (&o)[4], const uint32_t (&u)[4],
                                          const uint32_t (&m)[4], const uint32_t (&d)[4], const uint32_t side,
                                          const uint32_t ones_r, const uint32_t ra, const uint32_t rb, uint32_t& S4,
                                          uint32_t& idx4, uint32_t& xr) {
    // horizontal neighbour that is not at the same k: k+1 for par = 1, k-1 for par = 0
    const uint32_t sr = __funnelshift_r(m[wi], wi < 3 ? m[wi + 1] : side, 8);
    const uint32_t sl = __funnelshift_l(wi > 0 ? m[wi - 1] : side, m[wi], 8);
    const uint32_t sh = par ? sr : sl;
    S4 = u[wi] + d[wi] + m[wi] + sh;  // bytes 0..8, no carries (main.cpp:97-103)
    const uint32_t c4 = o[wi];
    // flips (bit 15 of each halfword) as byte masks via PRMT sign replication; fO = f + 1
    const uint32_t fO = and_xor(prmt(ra, rb, 0xFDB9u), 0x03030303u, ones_r);  // one LOP3
    // proposal xor x = c ^ p = ((2c + f) mod 3) + 1   (main.cpp:94 in code space)
    const uint32_t q1 = imad(c4, 2u, fO);            // 2c + f + 1 in 1..6
    // table entry numbers (biased by +9 through f+1): 18c + 9(f+1) + S = 9 q1 + S <= 62 -- one IMAD on top of q1; the
    // factor 2 of the u16 byte offset is applied by the IDP.4A that extracts each byte
    idx4 = imad(q1, 9u, S4);
    // (q1 + [q1 >= 4]) & 3: the shift drags the low bits of the next byte into bits 6..7, the byte sums stay
    // below 256 (no carry) and the mask drops them.  Shift + add fuse into one LEA.HI: two ALU-pipe
    // instructions, none on the FMA-heavy pipe that Philox saturates.
    xr = (q1 + (q1 >> 2)) & 0x03030303u;
}

// Branch-free form of update_half8 (streaming pair-table kernels): sites whose coarse compare ties are REJECTED here
// and only lower the running minimum `tiemin`; the caller tests it once per block of rows (has_tie + warp vote) and
// repairs the block out of line (fix_tied_block in bc_sweep_stream.cuh).
template <int COL, bool MEASURE, int HALF, bool PAIR>
__device__ __forceinline__ void update_half8_bf(const RowCtx& C, const int par, const uint32_t (&o)[4],
                                             const uint32_t (&u)[4], const uint32_t (&m)[4], const uint32_t (&d)[4],
                                             const uint32_t side, const uint32_t ones_r, const uint32_t (&q)[4],
                                             uint32_t (&nw)[4], uint32_t& a_c, uint32_t& a_c2, uint32_t& a_cS,
                                             uint32_t& a_S, uint32_t& tiemin) {
    const char* tabm = C.tabm;
    constexpr int half = HALF;
    // q = the coarse randoms of this half: one Philox call per 8 sites, 16 bits per site (coarse_draw)
#pragma unroll
    for (int w2 = 0; w2 < 2; ++w2) {
        const int wi = 2 * half + w2;
        const uint32_t ra = q[2 * w2], rb = q[2 * w2 + 1];
        uint32_t S4, idx4, xr;
        site_word(wi, par, o, u, m, d, side, ones_r, ra, rb, S4, idx4, xr);
        uint32_t Z01, Z23;
        if (PAIR) {
            // Pair table: ONE lookup per two sites.  Byte offset 4*(b0-9) + 216*(b1-9) by one IDP.4A (the
            // bias and the shared-space base are its accumulator), LDS.32 returns the packed word T with
            // ra - T = lanes 0x8000 + u15 - E15 exactly: the flip bits of ra (bits 15, 31) are known to the
            // table through the entry numbers and cancel, so no guard-bit OR and no PRMT to pack two
            // thresholds (see pair_table_kernel).
            const uint32_t a01 = __dp4a(idx4, 0x0000D804u, C.tab2s);
            const uint32_t a23 = __dp4a(idx4, 0xD8040000u, C.tab2s);
            uint32_t T01, T23;
            asm("ld.shared.u32 %0, [%1];" : "=r"(T01) : "r"(a01));
            asm("ld.shared.u32 %0, [%1];" : "=r"(T23) : "r"(a23));
            Z01 = ra - T01;
            Z23 = rb - T23;
        } else {
            // per-site table offsets: byte i of idx4, extracted on the FMA pipe (IDP.4A)
            const uint32_t a0 = __dp4a(idx4, 0x00000002u, 0u);
            const uint32_t a3 = __dp4a(idx4, 0x02000000u, 0u);
            const uint32_t a1 = __dp4a(idx4, 0x00000200u, 0u);
            const uint32_t a2 = __dp4a(idx4, 0x00020000u, 0u);
            const uint32_t E0 = *reinterpret_cast<const uint16_t*>(tabm + a0);
            const uint32_t E1 = *reinterpret_cast<const uint16_t*>(tabm + a1);
            const uint32_t E2 = *reinterpret_cast<const uint16_t*>(tabm + a2);
            const uint32_t E3 = *reinterpret_cast<const uint16_t*>(tabm + a3);
            // Packed 2-site compare: both 15-bit draws of a word get a guard bit, Z lane = 0x8000 + u15 - E15
            // in [0, 0xFFFF] (no cross-lane borrow): bit 15 clear <=> u15 < E15 (accept), lane == 0x8000 <=>
            // tie.  E15 = thr31 >> 16 (32768 = always accept: lane = u15, never a tie).
            Z01 = (ra | 0x80008000u) - __byte_perm(E0, E1, 0x5410);
            Z23 = (rb | 0x80008000u) - __byte_perm(E2, E3, 0x5410);
        }
        const uint32_t msk = prmt(Z01, Z23, 0xFDB9u);       // 0xFF where u15 >= E15 (REJECT mask)
        // tie <=> some lane == 0x8000, the most negative signed 16-bit value: one packed min per word pair
        tiemin = min_s16x2(tiemin, min_s16x2(Z01, Z23));
        const uint32_t v = o[wi] ^ (xr & ~msk);
        if (MEASURE) {
            a_c = __dp4a(v, ONES, a_c);
            a_c2 = __dp4a(v, v, a_c2);
            if (COL == 1) {
                a_cS = __dp4a(v, S4, a_cS);
                a_S = __dp4a(S4, ONES, a_S);
            }
        }
        nw[wi] = v;
    }
}

// tiemin = packed signed minimum of all compare lanes since it was initialised: a lane == 0x8000 <=> a tie.  High lane: the word, as a
// signed number, is below 0x80010000; low lane: shifted to the top.
constexpr uint32_t TIEMIN_INIT = 0x7FFF7FFFu;  // running minimum: the caller initialises it and tests it when it likes
__device__ __forceinline__ bool has_tie(const uint32_t tiemin) {
    return (int32_t)tiemin < (int32_t)0x80010000u || (tiemin << 16) == 0x80000000u;
}

template <int COL, bool MEASURE, bool PAIR = false>
__device__ __forceinline__ uint4 update_chunk16_bf(const RowCtx& C, const uint32_t y, const int par, const uint4& upv,
                                                const uint4& midv, const uint4& dnv, const uint32_t side,
                                                const uint4& ownv, uint32_t& a_c, uint32_t& a_c2, uint32_t& a_cS,
                                                uint32_t& a_S, uint32_t& tiemin) {
    uint32_t ones_r;  // ONES held in a register the compiler cannot fold back into an immediate
    asm volatile("mov.u32 %0, 0x01010101;" : "=r"(ones_r));
    const uint32_t o[4] = {ownv.x, ownv.y, ownv.z, ownv.w};
    const uint32_t u[4] = {upv.x, upv.y, upv.z, upv.w};
    const uint32_t m[4] = {midv.x, midv.y, midv.z, midv.w};
    const uint32_t d[4] = {dnv.x, dnv.y, dnv.z, dnv.w};
    uint32_t nw[4], q[4];
    coarse_draw(C, y, 0u, q);
    update_half8_bf<COL, MEASURE, 0, PAIR>(C, par, o, u, m, d, side, ones_r, q, nw, a_c, a_c2, a_cS, a_S, tiemin);
    coarse_draw(C, y, 1u, q);
    update_half8_bf<COL, MEASURE, 1, PAIR>(C, par, o, u, m, d, side, ones_r, q, nw, a_c, a_c2, a_cS, a_S, tiemin);
    return make_uint4(nw[0], nw[1], nw[2], nw[3]);
}

}  // namespace bc
