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

// Software-pipelined form for the streaming kernel: on entry qn holds the coarse draws of (row y, half 0); on exit
// those of (row y_next, half 0).  The draws of the NEXT half are issued in the same basic block as the site
// arithmetic of the current half (the rare-tie branch ends a block), so the scheduler can interleave the
// IMAD.WIDE chain of Philox (FMA-heavy pipe) with the LOP3/PRMT work (ALU pipe) of independent data instead of
// running them back to back.
template <int COL, bool MEASURE, bool PAIR>
__device__ __forceinline__ uint4 update_chunk16_sp(const RowCtx& C, const uint32_t y, const uint32_t y_next, const int par,
                                                   const uint4& upv, const uint4& midv, const uint4& dnv,
                                                   const uint32_t side, const uint4& ownv, uint32_t (&qn)[4],
                                                   uint32_t& a_c, uint32_t& a_c2, uint32_t& a_cS, uint32_t& a_S,
                                                   uint32_t& n_ties) {
    uint32_t ones_r;
    asm volatile("mov.u32 %0, 0x01010101;" : "=r"(ones_r));
    const uint32_t o[4] = {ownv.x, ownv.y, ownv.z, ownv.w};
    const uint32_t u[4] = {upv.x, upv.y, upv.z, upv.w};
    const uint32_t m[4] = {midv.x, midv.y, midv.z, midv.w};
    const uint32_t d[4] = {dnv.x, dnv.y, dnv.z, dnv.w};
    uint32_t nw[4], q1[4];
    const uint32_t q0[4] = {qn[0], qn[1], qn[2], qn[3]};
    coarse_draw(C, y, 1u, q1);
    update_half8<COL, MEASURE, 0, PAIR>(C, y, par, o, u, m, d, side, ones_r, q0, nw, a_c, a_c2, a_cS, a_S, n_ties);
    coarse_draw(C, y_next, 0u, qn);
    update_half8<COL, MEASURE, 1, PAIR>(C, y, par, o, u, m, d, side, ones_r, q1, nw, a_c, a_c2, a_cS, a_S, n_ties);
    return make_uint4(nw[0], nw[1], nw[2], nw[3]);
}

// ------------------------------------------------------------------------------------------
// Fast kernel: L % 32 == 0.  One thread = one 16-byte chunk column j, marching down the
// L/strips rows of its strip with the three rows of the other colour held in registers and the
// loads of the NEXT row (other-colour row y+2, side word, own chunk) issued before the current
// row is computed (software prefetch: the kernel is instruction-bound, so exposed load latency
// is pure loss).
// Work item id = blockIdx.x*128 + threadIdx.x -> (replica, strip, j), j fastest, so a warp
// reads 512 contiguous bytes per row; items/replica is a multiple of 32 (host guarantees),
// so a warp never spans replicas.
//   SMALL = false: rows per strip is even (row parity static per unrolled row) and
//                  items/replica is a multiple of 128 (one replica, one table per CTA).
//   SMALL = true : any strip height, up to 4 replicas per CTA (small lattices).
// Per-site compare: table offsets extracted with IDP.4A (FMA-heavy pipe), LDS.U16
// from a 27-word conflict-free table, two sites per packed 16-bit subtract, reject masks by PRMT sign
// replication, ties (coarse draw == coarse threshold) by a packed signed minimum (VIMNMX3.S16x2).  The
// two half-rate integer pipes (ALU: LOP3/PRMT/SHF, FMA-heavy: IMAD/IDP, quarter-rate IMAD.WIDE) bound
// this kernel, not HBM; the instruction mix is balanced between them (tools/microbench/pipes.cu).
// ------------------------------------------------------------------------------------------
constexpr int FAST_THREADS = 128;
constexpr int FAST_MAX_REPL_PER_CTA = FAST_THREADS / 32;
#ifndef BC_DEEP_PREFETCH
#define BC_DEEP_PREFETCH 1  // 0: one row ahead everywhere, 1: two rows ahead where it pays (see DEEP), 2: everywhere
#endif
#ifndef BC_PAIR_TMA
#define BC_PAIR_TMA 1  // pair table staged by one cp.async.bulk + mbarrier (0: by cp.async from all threads)
#endif
#ifndef BC_FAST_MIN_BLOCKS
#define BC_FAST_MIN_BLOCKS 6
#endif

//   HALO  = true : slab boundary-strip launch in peer-memory mode: the thread that produces a chunk of the
//                  slab's first / last owned row also stores it into the bottom / top ghost row of the rank
//                  above / below (peer pointers over NVLink): compute and halo transfer are one kernel.
//   PAIR >= 1    : (needs !SMALL) the CTA also stages its replica's 54 x 54 pair table (11.4 KB) and looks the
//                  thresholds of two sites up with one IDP.4A + one LDS.32; pays when a CTA marches over enough
//                  rows to amortise the staging (host: LaunchPlan::pair).
//   PAIR == 2    : additionally the software-pipelined Philox draws (plain periodic kernel only; long strips).
// Software pipelining of the Philox draws (update_chunk16_sp) needs ~8 more registers: it pays for the plain
// periodic pair-table kernel at 6 CTAs/SM (80 registers: 1778 -> 1813 updates/ns) and loses wherever it spills
// (72 registers) or drops the measuring / open variants to 5 CTAs/SM.
template <bool MEASURE, bool OPEN, int PAIR>
struct FastSwp { static constexpr bool value = PAIR == 2 && !MEASURE && !OPEN; };
template <bool SMALL, bool MEASURE, bool OPEN, int PAIR>
struct FastMinBlocks {
    static constexpr int value =
        (MEASURE || SMALL || FastSwp<MEASURE, OPEN, PAIR>::value) ? BC_FAST_MIN_BLOCKS : BC_FAST_MIN_BLOCKS + 1;
};

// The kernel body; `bid` is the CTA's index among the CTAs that work on P's lattices (blockIdx.x for an ordinary
// launch, blockIdx.x minus the group's first CTA for a group launch, see sweep_group_kernel).
template <int COL, bool SMALL, bool MEASURE, bool OPEN, bool HALO, int PAIR>
__device__ __forceinline__ void sweep_fast_body(const SweepParams& P, const unsigned bid, const int64_t t_add = 0) {
    static_assert(!(PAIR && SMALL) && !(PAIR && HALO), "the pair table is one replica per CTA; HALO kernels use the per-site table");
    __shared__ __align__(16) uint32_t s_tab[(SMALL ? FAST_MAX_REPL_PER_CTA : 1) * TAB16_WORDS];
    __shared__ __align__(16) uint32_t s_tab2[PAIR ? TAB2_WORDS : 4];
    __shared__ uint4 s_rows[4][FAST_THREADS];    // thread-private ring of other-colour rows
    // DEEP: inputs are requested two rows ahead instead of one (the row above is then carried in registers and
    // the ring holds rows y, y+1 and the two rows in flight).  Measured (tools/tune_libs2.sh): +1..3 % for the
    // periodic pair-table kernels at 8..16 rows per thread, -1 % for the pipelined kernel, -2 % with open edges.
    constexpr bool DEEP = BC_DEEP_PREFETCH == 2 || (BC_DEEP_PREFETCH == 1 && PAIR == 1 && !OPEN);
    __shared__ uint4 s_own[DEEP ? 4 : 2][FAST_THREADS];     // thread-private own chunks (current / next rows)
    __shared__ uint32_t s_side[DEEP ? 4 : 2][FAST_THREADS];  // thread-private side words

    // Programmatic dependent launch: let the next half-sweep's grid be scheduled as soon as SMs free up; its
    // CTAs run their prologue (table staging, index arithmetic) and then block in griddepcontrol.wait until
    // this grid has completed and its writes are visible.  Both are no-ops for an ordinary launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const uint32_t ipr = (uint32_t)(P.strips * P.cpr);  // items per replica
    const unsigned long long gid0 = (unsigned long long)bid * FAST_THREADS;
    const uint32_t rep0 = (uint32_t)(gid0 / ipr);
    for (int i = threadIdx.x; i < (SMALL ? FAST_MAX_REPL_PER_CTA : 1) * TAB16_WORDS; i += FAST_THREADS) {
        const uint32_t r = rep0 + i / TAB16_WORDS;
        s_tab[i] = (r < (uint32_t)P.n_replicas)
                       ? reinterpret_cast<const uint32_t*>(P.tab16)[(size_t)r * TAB16_WORDS + i % TAB16_WORDS]
                       : 0u;
    }
#if BC_PAIR_TMA
    // Pair table by ONE bulk copy (TMA engine, UBLKCP): thread 0 arms an mbarrier with the byte count and issues the
    // copy; it completes in the background while the CTA computes its indices, waits for the previous grid and
    // requests its first rows -- every thread waits on the mbarrier just before the first lookup.  (The tables are
    // written by bc_set_params, never by a sweep: safe before griddepcontrol.wait.)
    __shared__ __align__(8) unsigned long long s_mbar;
    if (PAIR && threadIdx.x == 0) {
        const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&s_mbar), dst = (uint32_t)__cvta_generic_to_shared(s_tab2);
        const uint32_t* src = P.tab2 + (size_t)rep0 * TAB2_WORDS;
        constexpr uint32_t bytes = TAB2_WORDS * 4;
        static_assert(bytes % 16 == 0, "bulk copies move multiples of 16 bytes");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
    }
#else
    if (PAIR) {  // the tables are written by bc_set_params, never by a sweep: safe before griddepcontrol.wait
        const uint4* src = reinterpret_cast<const uint4*>(P.tab2 + (size_t)rep0 * TAB2_WORDS);
        for (int i = threadIdx.x; i < TAB2_WORDS / 4; i += FAST_THREADS) cp_async16(&s_tab2[4 * i], src + i);
        cp_async_commit();
        cp_async_wait<0>();
    }
#endif
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
    const uint32_t rows = (uint32_t)P.rows_per_strip;
    const uint32_t y0 = (uint32_t)P.row_first + ((uint32_t)P.strip_base + strip * (uint32_t)P.strip_stride) * rows;  // array row
    const uint32_t yg0 = y0 + (uint32_t)P.y_rng_off;
    const uint32_t Wc = (uint32_t)P.Wc;
    constexpr bool open = OPEN;

    int64_t t = P.t + t_add;
    if (P.t_ptr) t += *P.t_ptr;
    const uint32_t k0 = P.key0, k1 = P.key1 ^ rep;
    const uint32_t ctr2 = (uint32_t)((uint64_t)t & 0xFFFFFFFFu);
    const uint32_t ctr3 = ctr3_of(t, STREAM_COARSE, COL);

    // shared-memory table of this thread's replica: uniform for !SMALL
    const char* tab = reinterpret_cast<const char*>(s_tab) + (SMALL ? (rep - rep0) * (TAB16_WORDS * 4) : 0u);
    const size_t rep_base = (size_t)rep * (size_t)P.rows_arr * (size_t)Wc;
    uint8_t* own_col = P.own + rep_base + (size_t)j * 16;            // + y*Wc
    const uint8_t* oth_col = P.other + rep_base + (size_t)j * 16;    // + y*Wc
    const uint8_t* oth_rep = P.other + rep_base;
    const uint32_t jr = (j + 1 == (uint32_t)P.cpr) ? 0u : j + 1;     // chunk holding k+1 of our last site
    const uint32_t jl = (j == 0) ? (uint32_t)P.cpr - 1 : j - 1;      // chunk holding k-1 of our first site
    const uint32_t side_r = jr * 16, side_l = jl * 16 + 12;          // byte offsets inside a row
    const bool edge_r = open && (j + 1 == (uint32_t)P.cpr), edge_l = open && (j == 0);

    const uint4 ones4 = make_uint4(ONES, ONES, ONES, ONES);
    const uint32_t total = (uint32_t)P.rows_arr * Wc;  // bytes of one replica's colour array (< 2^32)

    uint32_t a_c = 0, a_c2 = 0, a_cS = 0, a_S = 0, n_ties = 0;

    // The 18*(f+1) term of the table offset is computed with f+1 instead of f: bias the base by -18.
    const char* tabm = tab - 18;

    RowCtx ctx;
    ctx.j = j; ctx.ctr2 = ctr2; ctx.ctr3 = ctr3; ctx.k0 = k0; ctx.k1 = k1; ctx.t = t; ctx.tabm = tabm;
    ctx.t31_rep = P.tab31 + (size_t)rep * TABLE_ENTRIES;
    ctx.tab2s = PAIR ? (uint32_t)__cvta_generic_to_shared(s_tab2) - TAB2_BIAS : 0u;
    constexpr bool SWP = FastSwp<MEASURE, OPEN, PAIR>::value;
    uint32_t qn[4];
    if (SWP) coarse_draw(ctx, yg0, 0u, qn);
    auto update_row = [&](const uint32_t y, const uint32_t off, const int par, const uint4& upv, const uint4& midv,
                          const uint4& dnv, const uint32_t side, const uint4& ownv) {
        if (SWP) {
            *reinterpret_cast<uint4*>(own_col + off) = update_chunk16_sp<COL, MEASURE, (PAIR > 0)>(
                ctx, y, y + 1, par, upv, midv, dnv, side, ownv, qn, a_c, a_c2, a_cS, a_S, n_ties);
        } else if (!HALO) {
            *reinterpret_cast<uint4*>(own_col + off) =
                update_chunk16<COL, MEASURE, (PAIR > 0)>(ctx, y, par, upv, midv, dnv, side, ownv, a_c, a_c2, a_cS, a_S, n_ties);
        } else {
            const uint4 nv = update_chunk16<COL, MEASURE, (PAIR > 0)>(ctx, y, par, upv, midv, dnv, side, ownv, a_c, a_c2, a_cS, a_S, n_ties);
            *reinterpret_cast<uint4*>(own_col + off) = nv;
            // fused halo transfer: first owned row -> bottom ghost row of the rank above, last owned row -> top
            // ghost row of the rank below (same array geometry on every rank)
            if (P.peer_prev && off == (uint32_t)P.row_first * Wc)
                *reinterpret_cast<uint4*>(P.peer_prev + rep_base + (size_t)(P.Ly + 1) * Wc + (size_t)j * 16) = nv;
            if (P.peer_next && off == (uint32_t)(P.row_first + P.Ly - 1) * Wc)
                *reinterpret_cast<uint4*>(P.peer_next + rep_base + (size_t)j * 16) = nv;
        }
    };

    // ---- row marching.  Thread-private staging slots in shared memory (a ring of four other-colour
    // ---- rows, two own chunks, two side words) filled by cp.async one row ahead: the inputs of row
    // ---- y+1 are in flight while row y is computed, and they cost no registers while in flight.
    const uint32_t tid = threadIdx.x;
    auto stage_other = [&](const int slot, uint32_t o) {  // other-colour row at byte offset o (== total: wrap/open)
        if (o == total) {
            if (open) { s_rows[slot][tid] = ones4; return; }
            o = 0;
        }
        cp_async16(&s_rows[slot][tid], oth_col + o);
    };
    auto stage_row_inputs = [&](const int slot2, const uint32_t off_row, const int par) {  // own chunk + side word
        cp_async16(&s_own[slot2][tid], own_col + off_row);
        if (par) {
            if (edge_r) s_side[slot2][tid] = 0x00000001u; else cp_async4(&s_side[slot2][tid], oth_rep + off_row + side_r);
        } else {
            if (edge_l) s_side[slot2][tid] = 0x01000000u; else cp_async4(&s_side[slot2][tid], oth_rep + off_row + side_l);
        }
    };

    asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous half-sweep (other colour) is complete
    uint32_t off = y0 * Wc;  // byte offset of row y inside the replica's colour array
#if BC_PAIR_TMA
    auto wait_pair_table = [&]() {  // phase 0 of the mbarrier: the body runs once per CTA in the PAIR kernels
        if (PAIR) {
            const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&s_mbar);
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "PAIR_TABLE_WAIT:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
                "@p bra PAIR_TABLE_READY;\n"
                "bra PAIR_TABLE_WAIT;\n"
                "PAIR_TABLE_READY:\n"
                "}\n" ::"r"(mb) : "memory");
        }
    };
#else
    auto wait_pair_table = [&]() {};
#endif
    if (!DEEP) {
        if (y0 == 0) { if (open) s_rows[0][tid] = ones4; else cp_async16(&s_rows[0][tid], oth_col + (total - Wc)); }
        else cp_async16(&s_rows[0][tid], oth_col + (off - Wc));
        cp_async16(&s_rows[1][tid], oth_col + off);
        stage_other(2, off + Wc);
        stage_row_inputs(0, off, (int)((yg0 + COL) & 1u));
        cp_async_commit();
        wait_pair_table();

        // row r of the strip: ring slots up = r&3, mid = (r+1)&3, dn = (r+2)&3, next = (r+3)&3
        auto march = [&](const uint32_t r, const int par, const bool last, const int sl) {
            if (!last) {
                stage_other((sl + 3) & 3, off + 2 * Wc);
                stage_row_inputs((sl + 1) & 1, off + Wc, 1 - par);
            }
            cp_async_commit();
            cp_async_wait<1>();  // everything but the group just committed has landed: row r is ready
            const uint4 upv = s_rows[sl & 3][tid], midv = s_rows[(sl + 1) & 3][tid], dnv = s_rows[(sl + 2) & 3][tid];
            const uint4 ownv = s_own[sl & 1][tid];
            const uint32_t side = s_side[sl & 1][tid];
            update_row(yg0 + r, off, par, upv, midv, dnv, side, ownv);
            off += Wc;
        };

        if (SMALL || (rows & 3u)) {
            for (uint32_t r = 0; r < rows; ++r) march(r, (int)((yg0 + r + COL) & 1u), r + 1 == rows, (int)(r & 3u));
        } else {
            // yg0 is even and rows is a multiple of 4: parities alternate COL, 1-COL statically; static ring slots
            for (uint32_t r = 0; r < rows; r += 4) {
                march(r, COL, false, 0);
                march(r + 1, 1 - COL, false, 1);
                march(r + 2, COL, false, 2);
                march(r + 3, 1 - COL, r + 4 == rows, 3);
            }
        }
    } else {
        // Strip row r lives in ring slot r&3 (other colour) and r&3 (own chunk, side word).  Group G_r, committed at
        // the top of row r, holds other row r+3 and the own chunk / side word of row r+2; row r needs G_{r-2}, so
        // two groups stay in flight (wait_group 2).  Row r-1 of the other colour is carried in registers.
        const int par0 = (int)((yg0 + COL) & 1u);
        uint4 upv;
        if (y0 == 0) upv = open ? ones4 : __ldcg(reinterpret_cast<const uint4*>(oth_col + (total - Wc)));
        else upv = __ldcg(reinterpret_cast<const uint4*>(oth_col + (off - Wc)));
        cp_async16(&s_rows[0][tid], oth_col + off);
        stage_other(1, off + Wc);
        stage_row_inputs(0, off, par0);
        cp_async_commit();  // G_{-2}
        if (rows > 1) {
            stage_other(2, off + 2 * Wc);
            stage_row_inputs(1, off + Wc, 1 - par0);
        }
        cp_async_commit();  // G_{-1}
        wait_pair_table();

        auto march = [&](const uint32_t r, const int par, const bool more, const int sl) {
            if (more) {  // row r+2 exists (same parity as row r)
                stage_other((sl + 3) & 3, off + 3 * Wc);
                stage_row_inputs((sl + 2) & 3, off + 2 * Wc, par);
            }
            cp_async_commit();
            cp_async_wait<2>();
            const uint4 midv = s_rows[sl & 3][tid], dnv = s_rows[(sl + 1) & 3][tid];
            const uint4 ownv = s_own[sl & 3][tid];
            const uint32_t side = s_side[sl & 3][tid];
            update_row(yg0 + r, off, par, upv, midv, dnv, side, ownv);
            upv = midv;
            off += Wc;
        };

        if (SMALL || (rows & 3u)) {
            for (uint32_t r = 0; r < rows; ++r) march(r, (int)((yg0 + r + COL) & 1u), r + 3 <= rows, (int)(r & 3u));
        } else {
            for (uint32_t r = 0; r < rows; r += 4) {
                const bool more = r + 4 != rows;
                march(r, COL, true, 0);
                march(r + 1, 1 - COL, true, 1);
                march(r + 2, COL, more, 2);
                march(r + 3, 1 - COL, more, 3);
            }
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

template <int COL, bool SMALL, bool MEASURE, bool OPEN, bool HALO = false, int PAIR = 0>
__global__ void __launch_bounds__(FAST_THREADS, FastMinBlocks<SMALL, MEASURE, OPEN, PAIR>::value)
sweep_fast_kernel(const __grid_constant__ SweepParams P) {
    sweep_fast_body<COL, SMALL, MEASURE, OPEN, HALO, PAIR>(P, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// Group launch: ONE grid updates the lattices of up to GROUP_MAX engine handles of different sizes (a
// finite-size-scaling ladder, config 3; the open-boundary size series, config 5).  Each handle keeps its own
// geometry, tables, Philox key, sweep counter and observable slots (its SweepParams); the CTAs are dealt out
// handle after handle and a CTA finds its handle by comparing blockIdx.x with the prefix sums.  Per-handle
// launches of small lattices are launch-latency-bound and, on separate streams, queue behind the big
// lattice's CTAs; one grid has one tail.  Same body, same results as sweep_fast_kernel.
// ------------------------------------------------------------------------------------------
constexpr int GROUP_MAX = 8;
struct GroupParams {
    SweepParams p[GROUP_MAX];
    unsigned cta_end[GROUP_MAX];  // exclusive prefix end of each handle's CTAs
    int n;
};

template <int COL, bool MEASURE, bool OPEN, int PAIR>
__global__ void __launch_bounds__(FAST_THREADS, BC_FAST_MIN_BLOCKS)
sweep_group_kernel(const __grid_constant__ GroupParams G) {
    int g = 0;
#pragma unroll
    for (int i = 0; i < GROUP_MAX - 1; ++i)
        if (i + 1 < G.n && blockIdx.x >= G.cta_end[i]) g = i + 1;
    const unsigned first = g ? G.cta_end[g - 1] : 0u;
    sweep_fast_body<COL, false, MEASURE, OPEN, false, PAIR>(G.p[g], blockIdx.x - first);
}

// ------------------------------------------------------------------------------------------
// Persistent kernel -- experimental, off by default (a measured negative result): ALL half-sweeps of a
// run of unmeasured sweeps in ONE cooperative launch, with a grid-wide barrier between half-sweeps instead of a
// kernel boundary.  The idea: for lattices that do not fill the chip for long (one L = 4096 lattice is 4.3 us of
// work per half-sweep) the floor of a dependent launch -- about 3 us even from a CUDA graph with programmatic
// dependent launch -- is most of the time.  The measurement: a software barrier is four serialised L2 round trips
// (the stores' fence, the arrival atomic, the poll that sees the release, the first loads after it), and nothing
// overlaps them, while programmatic dependent launch runs the next grid's prologue under the previous grid's tail:
// one L = 4096 lattice 761 vs 1090 updates/ns, L = 2048 358 vs 588 (profiles/r01_tune_persist.log).
// The grid is exactly the number of CTAs that are resident at once (cooperative launch: co-residency is
// guaranteed or the launch fails); a CTA takes the 128-item work blocks b = blockIdx.x, blockIdx.x + gridDim.x, ...
// of every half-sweep and runs the streaming kernel's body on them (same results).  Handles several lattice sizes
// like sweep_group_kernel (n = 1: a single handle).  The barrier: a monotonically increasing arrival counter (the
// g-th barrier is complete at count g * gridDim.x), the last arriver publishes the generation with a release
// store, the others poll it with acquire loads (which also drop stale L1 lines before the next half-sweep reads
// its neighbours); the wait is bounded (~2 s) and raises bar[2] instead of hanging the GPU.
// ------------------------------------------------------------------------------------------
struct PersistParams {
    GroupParams g[2];   // members' parameters for the colour-0 and the colour-1 half-sweep (own / other swapped)
    unsigned* bar;      // [0] arrivals, [1] generation, [2] abort; zeroed by the host before the launch
    int n_half;         // half-sweeps to run, starting with colour 0 of the members' sweep counters
};

__device__ __forceinline__ bool persist_barrier(unsigned* bar, const unsigned gen) {
    __shared__ int s_abort;
    __syncthreads();
    if (threadIdx.x == 0) {
        int give_up = 0;
        __threadfence();
        const unsigned arrived = atomicAdd(&bar[0], 1u) + 1u;
        if (arrived == gen * gridDim.x) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen) : "memory");
        } else {
            const long long c0 = clock64();
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar + 1) : "memory");
                if (v < gen && clock64() - c0 > 4000000000LL) { atomicExch(&bar[2], 1u); give_up = 1; break; }
            } while (v < gen);
        }
        if (*reinterpret_cast<volatile unsigned*>(bar + 2)) give_up = 1;
        __threadfence();
        s_abort = give_up;
    }
    __syncthreads();
    return s_abort == 0;
}

#ifndef BC_PERSIST_MIN_BLOCKS
#define BC_PERSIST_MIN_BLOCKS 7
#endif
template <bool OPEN>
__global__ void __launch_bounds__(FAST_THREADS, BC_PERSIST_MIN_BLOCKS)
sweep_persist_kernel(const __grid_constant__ PersistParams P) {
    const int n = P.g[0].n;
    const unsigned n_blocks = P.g[0].cta_end[n - 1];
    for (int h = 0; h < P.n_half; ++h) {
        const GroupParams& G = P.g[h & 1];
        for (unsigned b = blockIdx.x; b < n_blocks; b += gridDim.x) {
            int g = 0;
#pragma unroll
            for (int i = 0; i < GROUP_MAX - 1; ++i)
                if (i + 1 < n && b >= G.cta_end[i]) g = i + 1;
            const unsigned first = g ? G.cta_end[g - 1] : 0u;
            if (h & 1) sweep_fast_body<1, false, false, OPEN, false, 0>(G.p[g], b - first, (int64_t)(h >> 1));
            else sweep_fast_body<0, false, false, OPEN, false, 0>(G.p[g], b - first, (int64_t)(h >> 1));
            __syncthreads();  // the table and the staging slots are reused by the next block
        }
        if (h + 1 < P.n_half && !persist_barrier(P.bar, (unsigned)(h + 1))) return;
    }
}

// ------------------------------------------------------------------------------------------
// Wavefront kernel: ALL half-sweeps of a call in one launch for lattices that live in L2 (a single
// L = 4096 lattice is 16 MiB), without any grid-wide barrier.  The work of one half-sweep is cut into the
// same 128-item blocks as the streaming kernel; a task is (half-sweep h, replica, block b) and tasks are
// handed out in increasing id by an atomic counter.  Task (h, b) only needs the blocks of half-sweep h-1
// that own its neighbouring rows / chunks (they wrote the other colour it reads, and they were the last
// readers of the rows it overwrites), so a CTA polls their per-block progress counters (ld.acquire.gpu) and
// publishes its own (st.release.gpu) -- half-sweeps overlap like a wavefront and there is no launch per
// half-sweep.  Every dependency has a smaller task id, hence was handed out earlier to a CTA that is
// running: the smallest unfinished task can always proceed, whatever the grid size (no co-residency
// assumption).  Waits are bounded (~2 s) and raise *abort instead of hanging the GPU.
// All global reads of lattice data go through cp.async.cg (L2), never L1, so other CTAs' writes are seen.
// ------------------------------------------------------------------------------------------
struct WaveParams {
    uint8_t* c0;
    uint8_t* c1;
    const uint16_t* tab16;
    const uint32_t* tab31;
    unsigned long long* ties;
    int* next_task;  // [1] task counter (zeroed by the host)
    int* progress;   // [n_replicas * nblk] half-sweeps completed by each block (zeroed by the host)
    int* abort;      // [1] set when a wait timed out
    int64_t t0;      // first sweep index
    int n_half;      // half-sweeps to run, starting with colour 0 of sweep t0
    int L, Wc, cpr, rows_per_strip, strips, n_replicas, rows_arr, nblk;
    uint32_t key0, key1;
};

template <bool OPEN>
__global__ void __launch_bounds__(FAST_THREADS, 5)
sweep_wave_kernel(const __grid_constant__ WaveParams P) {
    __shared__ __align__(16) uint32_t s_tab[TAB16_WORDS];
    __shared__ uint4 s_rows[4][FAST_THREADS];
    __shared__ uint4 s_own[2][FAST_THREADS];
    __shared__ uint4 s_side[2][FAST_THREADS];  // the 16-byte vector holding the side word (cp.async.cg needs 16 B)
    __shared__ int s_task;
    const uint32_t tid = threadIdx.x;
    const uint32_t Wc = (uint32_t)P.Wc, rows = (uint32_t)P.rows_per_strip, cpr = (uint32_t)P.cpr;
    const uint32_t total = (uint32_t)P.rows_arr * Wc;
    const int nblk = P.nblk, per_half = P.n_replicas * nblk;
    const long long n_tasks = (long long)P.n_half * per_half;
    const uint4 ones4 = make_uint4(ONES, ONES, ONES, ONES);
    int cur_rep = -1;
    uint32_t n_ties = 0;

    for (;;) {
        __syncthreads();  // everyone is done with s_task / s_tab of the previous task
        if (tid == 0) s_task = (*reinterpret_cast<volatile int*>(P.abort)) ? -1 : atomicAdd(P.next_task, 1);
        __syncthreads();
        const int task = s_task;
        if (task < 0 || task >= n_tasks) break;
        const int h = task / per_half, rem_t = task - h * per_half;
        const int rep = rem_t / nblk, b = rem_t - rep * nblk;
        const int col = h & 1;
        const int64_t t = P.t0 + (h >> 1);

        // ---- wait for the blocks of half-sweep h-1 that touch our neighbourhood
        int give_up = 0;
        if (h > 0 && tid < 5) {
            int dep;
            if (cpr >= (uint32_t)FAST_THREADS) {
                const int nbs = (int)cpr / FAST_THREADS, s = b / nbs, cb = b - s * nbs, S = P.strips;
                const int su = s == 0 ? S - 1 : s - 1, sd = s + 1 == S ? 0 : s + 1;
                const int cl = cb == 0 ? nbs - 1 : cb - 1, cr = cb + 1 == nbs ? 0 : cb + 1;
                dep = tid == 0 ? su * nbs + cb : tid == 1 ? sd * nbs + cb : tid == 2 ? s * nbs + cl : tid == 3 ? b : s * nbs + cr;
            } else {
                dep = tid == 0 ? (b == 0 ? nblk - 1 : b - 1) : tid == 1 ? (b + 1 == nblk ? 0 : b + 1) : b;
            }
            const int* flag = P.progress + (size_t)rep * nblk + dep;
            const long long c0 = clock64();
            int v;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                if (v < h && clock64() - c0 > 4000000000LL) { *reinterpret_cast<volatile int*>(P.abort) = 1; give_up = 1; break; }
            } while (v < h);
        }
        if (rep != cur_rep) {  // (re)stage the replica's coarse table; the barrier below also orders it
            for (int i = tid; i < TAB16_WORDS; i += FAST_THREADS)
                s_tab[i] = reinterpret_cast<const uint32_t*>(P.tab16)[(size_t)rep * TAB16_WORDS + i];
            cur_rep = rep;
        }
        if (__syncthreads_or(give_up)) break;  // uniform decision for the whole CTA

        // ---- this thread's chunk column and strip, exactly as in the streaming kernel
        const uint32_t item = (uint32_t)b * FAST_THREADS + tid;
        const uint32_t strip = item / cpr, j = item - strip * cpr;
        const uint32_t y0 = strip * rows;
        const size_t rep_base = (size_t)rep * (size_t)P.rows_arr * (size_t)Wc;
        uint8_t* own_arr = col ? P.c1 : P.c0;
        const uint8_t* oth_arr = col ? P.c0 : P.c1;
        uint8_t* own_col = own_arr + rep_base + (size_t)j * 16;
        const uint8_t* oth_col = oth_arr + rep_base + (size_t)j * 16;
        const uint8_t* oth_rep = oth_arr + rep_base;
        const uint32_t jr = (j + 1 == cpr) ? 0u : j + 1, jl = (j == 0) ? cpr - 1 : j - 1;
        const bool edge_r = OPEN && (j + 1 == cpr), edge_l = OPEN && (j == 0);
        RowCtx ctx;
        ctx.j = j; ctx.k0 = P.key0; ctx.k1 = P.key1 ^ (uint32_t)rep; ctx.t = t;
        ctx.ctr2 = (uint32_t)((uint64_t)t & 0xFFFFFFFFu);
        ctx.ctr3 = ctr3_of(t, STREAM_COARSE, col);
        ctx.tabm = reinterpret_cast<const char*>(s_tab) - 18;
        ctx.t31_rep = P.tab31 + (size_t)rep * TABLE_ENTRIES;
        uint32_t dummy0 = 0, dummy1 = 0, dummy2 = 0, dummy3 = 0;

        auto stage_other = [&](const int slot, uint32_t o) {
            if (o == total) {
                if (OPEN) { s_rows[slot][tid] = ones4; return; }
                o = 0;
            }
            cp_async16(&s_rows[slot][tid], oth_col + o);
        };
        auto stage_row_inputs = [&](const int slot2, const uint32_t off_row, const int par) {
            cp_async16(&s_own[slot2][tid], own_col + off_row);
            if (par) {
                if (edge_r) s_side[slot2][tid] = make_uint4(0x00000001u, 0, 0, 0);
                else cp_async16(&s_side[slot2][tid], oth_rep + off_row + (size_t)jr * 16);
            } else {
                if (edge_l) s_side[slot2][tid] = make_uint4(0, 0, 0, 0x01000000u);
                else cp_async16(&s_side[slot2][tid], oth_rep + off_row + (size_t)jl * 16);
            }
        };
        uint32_t off = y0 * Wc;
        if (y0 == 0) { if (OPEN) s_rows[0][tid] = ones4; else cp_async16(&s_rows[0][tid], oth_col + (total - Wc)); }
        else cp_async16(&s_rows[0][tid], oth_col + (off - Wc));
        cp_async16(&s_rows[1][tid], oth_col + off);
        stage_other(2, off + Wc);
        stage_row_inputs(0, off, (int)((y0 + col) & 1u));
        cp_async_commit();

        auto run_strip = [&](auto colc) {
            constexpr int COL = decltype(colc)::value;
            auto march = [&](const uint32_t r, const int par, const bool last, const int sl) {
                if (!last) {
                    stage_other((sl + 3) & 3, off + 2 * Wc);
                    stage_row_inputs((sl + 1) & 1, off + Wc, 1 - par);
                }
                cp_async_commit();
                cp_async_wait<1>();
                const uint4 upv = s_rows[sl & 3][tid], midv = s_rows[(sl + 1) & 3][tid], dnv = s_rows[(sl + 2) & 3][tid];
                const uint4 ownv = s_own[sl & 1][tid];
                const uint4 sv = s_side[sl & 1][tid];
                const uint32_t side = par ? sv.x : sv.w;
                *reinterpret_cast<uint4*>(own_col + off) =
                    update_chunk16<COL, false>(ctx, y0 + r, par, upv, midv, dnv, side, ownv, dummy0, dummy1, dummy2, dummy3, n_ties);
                off += Wc;
            };
            for (uint32_t r = 0; r < rows; r += 4) {  // y0 and rows are multiples of 4 (host guarantees)
                march(r, COL, false, 0);
                march(r + 1, 1 - COL, false, 1);
                march(r + 2, COL, false, 2);
                march(r + 3, 1 - COL, r + 4 == rows, 3);
            }
        };
        if (col) run_strip(std::integral_constant<int, 1>{}); else run_strip(std::integral_constant<int, 0>{});
        cp_async_wait<0>();

        // ---- publish: all stores of the block, then the progress counter
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            int* flag = P.progress + (size_t)rep * nblk + b;
            const int v = h + 1;
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
        }
    }
    if (n_ties && P.ties) atomicAdd(P.ties, (unsigned long long)n_ties);
}

// ------------------------------------------------------------------------------------------
// Resident kernel: lattices small enough (L % 32 == 0, L <= 448) to stay in shared memory for a whole
// launch.  One CTA holds both colour arrays of its replica(s) and runs ALL n_sweeps sweeps of the call
// with __syncthreads() between half-sweeps -- no launch per half-sweep, no HBM traffic inside the loop.
// This is what the reference's own sizes (L = 8 ... 128, gen_files.py:11) and finite-size-scaling ladders
// need: at L = 64 a launch-per-half-sweep engine is launch-bound (~3 us per 2048 updates).
// Same spec, same update_chunk16 -> bit-identical to the streaming kernel and to the CPU restatement of the spec.
//   threads_per_rep T (32 / 128 / 256 / 512) threads share one replica; reps_per_cta = 128 / T for T < 128.
// ------------------------------------------------------------------------------------------
struct ResidentParams {
    uint8_t* c0;             // colour-0 array [replica][L][Wc]
    uint8_t* c1;             // colour-1 array
    const uint16_t* tab16;   // [replica][54]
    const uint32_t* tab31;   // [replica][54]
    unsigned long long* obs; // raw accumulators [slot][replica][OBS_RAW]
    unsigned long long* ties;
    int64_t t0;              // first sweep index
    int64_t n_sweeps;
    int64_t measure_every;   // 0 = never
    int L, Wc, cpr, n_replicas, threads_per_rep, reps_per_cta;
    uint32_t key0, key1;
};

//   ANY = true : any lattice size (the reference's 8, 12, 16, 24, 48, 96: gen_files.py:11, runner.sh:4; odd L with
//                open edges).  Rows are padded to 16-byte chunks with spin-0 bytes that must never change: the
//                chunk is updated as a whole and only its real sites are stored; on a torus the wrap-around
//                neighbours of a row's first / last site are patched in; the observables are summed here over
//                the real sites only (update_chunk16 would count the pad sites).
template <bool OPEN, bool ANY = false>
__global__ void __launch_bounds__(512)
sweep_resident_kernel(const __grid_constant__ ResidentParams P) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const int L = P.L, Wc = P.Wc, cpr = P.cpr, T = P.threads_per_rep;
    const uint32_t arr = (uint32_t)L * (uint32_t)Wc;            // bytes of one colour array
    const uint32_t rep_stride = 2u * arr + 128u;                 // two colours + coarse table (27 words, padded)
    const int lr = threadIdx.x / T, ltid = threadIdx.x - lr * T;
    const int rep = blockIdx.x * P.reps_per_cta + lr;
    const bool live = rep < P.n_replicas;
    uint8_t* s_c0 = s_dyn + (size_t)lr * rep_stride;
    uint8_t* s_c1 = s_c0 + arr;
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_c1 + arr);
    uint32_t* s_red = reinterpret_cast<uint32_t*>(s_dyn + (size_t)P.reps_per_cta * rep_stride) + lr * 8;

    // ---- load: global -> shared (16-byte vectors), table, zero the reduction scratch
    if (live) {
        const uint4* g0 = reinterpret_cast<const uint4*>(P.c0 + (size_t)rep * arr);
        const uint4* g1 = reinterpret_cast<const uint4*>(P.c1 + (size_t)rep * arr);
        for (uint32_t i = ltid; i < arr / 16; i += T) {
            reinterpret_cast<uint4*>(s_c0)[i] = g0[i];
            reinterpret_cast<uint4*>(s_c1)[i] = g1[i];
        }
        for (int i = ltid; i < TAB16_WORDS; i += T)
            s_tab[i] = reinterpret_cast<const uint32_t*>(P.tab16)[(size_t)rep * TAB16_WORDS + i];
        if (ltid < 8) s_red[ltid] = 0u;
    }
    __syncthreads();

    RowCtx ctx;
    ctx.k0 = P.key0; ctx.k1 = P.key1 ^ (uint32_t)rep;
    ctx.tabm = reinterpret_cast<const char*>(s_tab) - 18;
    ctx.t31_rep = P.tab31 + (size_t)(live ? rep : 0) * TABLE_ENTRIES;
    const uint4 ones4 = make_uint4(ONES, ONES, ONES, ONES);
    const int items = L * cpr;
    uint32_t n_ties = 0;
    int64_t slot = 0;

    const int y_first = ltid / cpr, j_first = ltid - y_first * cpr;
    // one half-sweep of colour COL over this replica's rows
    auto half_sweep = [&](auto colc, auto measc, uint32_t& a_c, uint32_t& a_c2, uint32_t& a_cS, uint32_t& a_S) {
        constexpr int COL = decltype(colc)::value;
        constexpr bool MEAS = decltype(measc)::value;
        uint8_t* own = COL ? s_c1 : s_c0;
        const uint8_t* oth = COL ? s_c0 : s_c1;
        for (int item = ltid; item < items; item += T) {
            // a thread's first item (its only one when T >= items, the latency-bound case) never changes: its row and
            // chunk are computed once per launch, not with an integer division per half-sweep
            int y = y_first, j = j_first;
            if (item != ltid) { y = item / cpr; j = item - y * cpr; }
            const int par = (y + COL) & 1;
            const uint32_t rowo = (uint32_t)y * Wc, co = (uint32_t)j * 16;
            // same-colour sites of this row: x = 2k + par < L (L/2 on a torus; odd L: one more where par == 0)
            const int nk = ANY ? (L - par + 1) >> 1 : 0;
            const int nreal = ANY ? min(16, nk - 16 * j) : 16;  // real sites in this chunk
            if (ANY && nreal <= 0) continue;                    // a chunk of pad bytes only (odd L)
            uint4 upv, dnv;
            if (y == 0) upv = OPEN ? ones4 : *reinterpret_cast<const uint4*>(oth + (arr - Wc) + co);
            else upv = *reinterpret_cast<const uint4*>(oth + rowo - Wc + co);
            if (y == L - 1) dnv = OPEN ? ones4 : *reinterpret_cast<const uint4*>(oth + co);
            else dnv = *reinterpret_cast<const uint4*>(oth + rowo + Wc + co);
            uint4 midv = *reinterpret_cast<const uint4*>(oth + rowo + co);
            uint32_t side;
            if (par) {
                if (j + 1 == cpr) side = OPEN ? 0x00000001u : *reinterpret_cast<const uint32_t*>(oth + rowo);
                else side = *reinterpret_cast<const uint32_t*>(oth + rowo + co + 16);
                if (ANY && !OPEN && nreal < 16) {
                    // torus: the right neighbour of the row's last site is x = 0, not the pad byte behind it
                    const uint32_t sh = 8u * ((uint32_t)nreal & 3u), w0 = (uint32_t)oth[rowo] << sh, keep = ~(0xFFu << sh);
                    const int wi = nreal >> 2;
                    if (wi == 0) midv.x = (midv.x & keep) | w0;
                    if (wi == 1) midv.y = (midv.y & keep) | w0;
                    if (wi == 2) midv.z = (midv.z & keep) | w0;
                    if (wi == 3) midv.w = (midv.w & keep) | w0;
                }
            } else {
                if (j == 0) {
                    if (OPEN) side = 0x01000000u;
                    else if (ANY) side = (uint32_t)oth[rowo + (L >> 1) - 1] << 24;  // torus: left neighbour of x = 0 is x = L-1
                    else side = *reinterpret_cast<const uint32_t*>(oth + rowo + (cpr - 1) * 16 + 12);
                } else {
                    side = *reinterpret_cast<const uint32_t*>(oth + rowo + co - 4);
                }
            }
            const uint4 ownv = *reinterpret_cast<const uint4*>(own + rowo + co);
            ctx.j = (uint32_t)j;
            if (!ANY) {
                *reinterpret_cast<uint4*>(own + rowo + co) =
                    update_chunk16<COL, MEAS>(ctx, (uint32_t)y, par, upv, midv, dnv, side, ownv, a_c, a_c2, a_cS, a_S, n_ties);
            } else {
                uint32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;
                const uint4 nv = update_chunk16<COL, false>(ctx, (uint32_t)y, par, upv, midv, dnv, side, ownv, d0, d1, d2, d3, n_ties);
                // byte masks of the real sites, word by word
                uint32_t mk[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int n = nreal - 4 * w;
                    mk[w] = n >= 4 ? 0xFFFFFFFFu : (n <= 0 ? 0u : (1u << (8 * n)) - 1u);
                }
                const uint4 st = make_uint4((nv.x & mk[0]) | (ownv.x & ~mk[0]), (nv.y & mk[1]) | (ownv.y & ~mk[1]),
                                            (nv.z & mk[2]) | (ownv.z & ~mk[2]), (nv.w & mk[3]) | (ownv.w & ~mk[3]));
                *reinterpret_cast<uint4*>(own + rowo + co) = st;
                if (MEAS) {
                    const uint32_t v[4] = {nv.x & mk[0], nv.y & mk[1], nv.z & mk[2], nv.w & mk[3]};
                    const uint32_t u[4] = {upv.x, upv.y, upv.z, upv.w}, d[4] = {dnv.x, dnv.y, dnv.z, dnv.w};
                    const uint32_t m[4] = {midv.x, midv.y, midv.z, midv.w};
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        a_c = __dp4a(v[w], ONES, a_c);
                        a_c2 = __dp4a(v[w], v[w], a_c2);
                        if (COL == 1) {  // neighbour sums as in update_half8 (main.cpp:97-103)
                            const uint32_t sr = __funnelshift_r(m[w], w < 3 ? m[w + 1] : side, 8);
                            const uint32_t sl = __funnelshift_l(w > 0 ? m[w - 1] : side, m[w], 8);
                            const uint32_t S4 = u[w] + d[w] + m[w] + (par ? sr : sl);
                            a_cS = __dp4a(v[w], S4, a_cS);
                            a_S = __dp4a(S4 & mk[w], ONES, a_S);
                        }
                    }
                }
            }
        }
    };

    // sweeps left until the next measured one (a 64-bit modulo per sweep would cost more than a small half-sweep)
    int64_t to_meas = P.measure_every > 0 ? P.measure_every - 1 - (P.t0 % P.measure_every) : -1;
    for (int64_t s = 0; s < P.n_sweeps; ++s) {
        const int64_t t = P.t0 + s;
        const bool measure = to_meas == 0;
        to_meas = measure ? P.measure_every - 1 : to_meas - 1;
        ctx.t = t;
        ctx.ctr2 = (uint32_t)((uint64_t)t & 0xFFFFFFFFu);
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
        ctx.ctr3 = ctr3_of(t, STREAM_COARSE, 0);
        if (live) {
            if (measure) half_sweep(std::integral_constant<int, 0>{}, std::true_type{}, a0, a1, a4, a5);
            else half_sweep(std::integral_constant<int, 0>{}, std::false_type{}, a0, a1, a4, a5);
        }
        __syncthreads();
        ctx.ctr3 = ctr3_of(t, STREAM_COARSE, 1);
        if (live) {
            if (measure) half_sweep(std::integral_constant<int, 1>{}, std::true_type{}, a2, a3, a4, a5);
            else half_sweep(std::integral_constant<int, 1>{}, std::false_type{}, a2, a3, a4, a5);
        }
        if (measure) {
            // block reduction of the raw sums of this replica, then one plain store per quantity
            if (live) {
                const uint32_t v[OBS_RAW] = {warp_sum(a0), warp_sum(a1), warp_sum(a2), warp_sum(a3), warp_sum(a4), warp_sum(a5)};
                if ((threadIdx.x & 31) == 0)
                    for (int k = 0; k < OBS_RAW; ++k) atomicAdd(&s_red[k], v[k]);
            }
            __syncthreads();
            if (live && ltid < OBS_RAW) {
                P.obs[((size_t)slot * P.n_replicas + rep) * OBS_RAW + ltid] = (unsigned long long)s_red[ltid];
                s_red[ltid] = 0u;
            }
            ++slot;
        }
        __syncthreads();
    }

    // ---- store back
    if (live) {
        uint4* g0 = reinterpret_cast<uint4*>(P.c0 + (size_t)rep * arr);
        uint4* g1 = reinterpret_cast<uint4*>(P.c1 + (size_t)rep * arr);
        for (uint32_t i = ltid; i < arr / 16; i += T) {
            g0[i] = reinterpret_cast<const uint4*>(s_c0)[i];
            g1[i] = reinterpret_cast<const uint4*>(s_c1)[i];
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

// ------------------------------------------------------------------------------------------
// Layout conversion, init, halo, standalone measurement.  Geometry of one replica's colour array:
// `rows_arr` rows of Wc bytes; the owned rows are array rows row_first .. row_first+Ly-1 and hold
// global rows y_glob0 .. y_glob0+Ly-1 (a slab owns a row range; ghost rows, if any, sit at array
// row 0 and rows_arr-1).
// ------------------------------------------------------------------------------------------
struct Geom {
    int L;          // columns (global linear size)
    int Wc;         // compact row stride
    int Ly;         // owned rows
    int rows_arr;   // array rows (Ly, or Ly + 2 with ghost rows)
    int row_first;  // array row of the first owned row
    int y_glob0;    // global row of the first owned row (even)
    int open;       // open boundary
    int ghost;      // ghost rows present (vertical neighbours never wrap inside the array)
};

// row-major int8 spins [rep][Ly][L]  ->  compact codes (both colours); pad bytes = code 1 (spin 0)
__global__ void pack_kernel(const int8_t* __restrict__ spins, uint8_t* __restrict__ c0, uint8_t* __restrict__ c1,
                            const Geom g, int rep_first, int n_rep) {
    const long long per_rep = (long long)g.Ly * g.Wc;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const long long rem = gid - (long long)r * per_rep;
    const int y = (int)(rem / g.Wc), k = (int)(rem - (long long)y * g.Wc);
    const size_t dst = ((size_t)(rep_first + r) * g.rows_arr + g.row_first + y) * g.Wc + k;
    const int8_t* src = spins + (size_t)r * g.Ly * g.L + (size_t)y * g.L;
    const int yg = y + g.y_glob0;
    const int xa = 2 * k + (yg & 1), xb = 2 * k + ((yg + 1) & 1);
    c0[dst] = (xa < g.L) ? (uint8_t)(src[xa] + 1) : (uint8_t)1;
    c1[dst] = (xb < g.L) ? (uint8_t)(src[xb] + 1) : (uint8_t)1;
}

__global__ void unpack_kernel(int8_t* __restrict__ spins, const uint8_t* __restrict__ c0,
                              const uint8_t* __restrict__ c1, const Geom g, int rep_first, int n_rep) {
    const long long per_rep = (long long)g.Ly * g.L;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const long long rem = gid - (long long)r * per_rep;
    const int y = (int)(rem / g.L), x = (int)(rem - (long long)y * g.L);
    const size_t src = ((size_t)(rep_first + r) * g.rows_arr + g.row_first + y) * g.Wc + (x >> 1);
    const uint8_t code = ((x + y + g.y_glob0) & 1) ? c1[src] : c0[src];
    spins[gid] = (int8_t)((int)code - 1);
}

// Vector forms for L % 32 == 0 (rows are 16-byte aligned, no pad bytes): one thread converts 32 consecutive sites
// of a row = one 16-byte chunk of each colour array.  grid = (ceil(Ly * L/32 / 256), replicas).
__device__ __forceinline__ uint32_t spins_to_codes4(uint32_t w) {  // bytes -1,0,1 -> 0,1,2 without inter-byte carries
    return ((w & 0x03030303u) + ONES) & 0x03030303u;
}
__device__ __forceinline__ uint32_t codes_to_spins4(uint32_t c) {  // bytes 0,1,2 -> 0xFF,0,1
    const uint32_t t = (c + 0x03030303u) & 0x03030303u;  // 3, 0, 1
    const uint32_t neg = (t >> 1) & ONES;                // 1 where the spin is -1
    return t | ((neg << 8) - neg);                        // neg * 0xFF per byte (mod 2^32)
}

__global__ void __launch_bounds__(256)
pack16_kernel(const int8_t* __restrict__ spins, uint8_t* __restrict__ c0, uint8_t* __restrict__ c1, const Geom g, int rep_first) {
    const uint32_t cpr = (uint32_t)g.L >> 5;
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= (uint32_t)g.Ly * cpr) return;
    const uint32_t r = blockIdx.y, y = idx / cpr, j = idx - y * cpr;
    const uint4* src = reinterpret_cast<const uint4*>(spins + ((size_t)r * g.Ly + y) * g.L + (size_t)j * 32);
    const uint4 a = __ldcs(src), b = __ldcs(src + 1);
    const uint4 ev = make_uint4(spins_to_codes4(__byte_perm(a.x, a.y, 0x6420)), spins_to_codes4(__byte_perm(a.z, a.w, 0x6420)),
                                spins_to_codes4(__byte_perm(b.x, b.y, 0x6420)), spins_to_codes4(__byte_perm(b.z, b.w, 0x6420)));
    const uint4 od = make_uint4(spins_to_codes4(__byte_perm(a.x, a.y, 0x7531)), spins_to_codes4(__byte_perm(a.z, a.w, 0x7531)),
                                spins_to_codes4(__byte_perm(b.x, b.y, 0x7531)), spins_to_codes4(__byte_perm(b.z, b.w, 0x7531)));
    const size_t dst = ((size_t)(rep_first + r) * g.rows_arr + g.row_first + y) * g.Wc + (size_t)j * 16;
    const bool odd_row = ((y + (uint32_t)g.y_glob0) & 1u) != 0;  // colour 0 holds x = 2k + (row & 1)
    *reinterpret_cast<uint4*>(c0 + dst) = odd_row ? od : ev;
    *reinterpret_cast<uint4*>(c1 + dst) = odd_row ? ev : od;
}

__global__ void __launch_bounds__(256)
unpack16_kernel(int8_t* __restrict__ spins, const uint8_t* __restrict__ c0, const uint8_t* __restrict__ c1, const Geom g,
                int rep_first) {
    const uint32_t cpr = (uint32_t)g.L >> 5;
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= (uint32_t)g.Ly * cpr) return;
    const uint32_t r = blockIdx.y, y = idx / cpr, j = idx - y * cpr;
    const size_t src = ((size_t)(rep_first + r) * g.rows_arr + g.row_first + y) * g.Wc + (size_t)j * 16;
    const bool odd_row = ((y + (uint32_t)g.y_glob0) & 1u) != 0;
    const uint4 p = *reinterpret_cast<const uint4*>(c0 + src), q = *reinterpret_cast<const uint4*>(c1 + src);
    const uint4 ev = odd_row ? q : p, od = odd_row ? p : q;
    uint4 a, b;
    a.x = codes_to_spins4(__byte_perm(ev.x, od.x, 0x5140)); a.y = codes_to_spins4(__byte_perm(ev.x, od.x, 0x7362));
    a.z = codes_to_spins4(__byte_perm(ev.y, od.y, 0x5140)); a.w = codes_to_spins4(__byte_perm(ev.y, od.y, 0x7362));
    b.x = codes_to_spins4(__byte_perm(ev.z, od.z, 0x5140)); b.y = codes_to_spins4(__byte_perm(ev.z, od.z, 0x7362));
    b.z = codes_to_spins4(__byte_perm(ev.w, od.w, 0x5140)); b.w = codes_to_spins4(__byte_perm(ev.w, od.w, 0x7362));
    uint4* dst = reinterpret_cast<uint4*>(spins + ((size_t)r * g.Ly + y) * g.L + (size_t)j * 32);
    __stcs(dst, a);
    __stcs(dst + 1, b);
}

// iid uniform {-1,0,1} (main.cpp:166-171) from Philox stream INIT, indexed by the GLOBAL site index x + y*L
__global__ void init_kernel(uint8_t* __restrict__ c0, uint8_t* __restrict__ c1, const Geom g, int rep_first,
                            int n_rep, uint32_t key0, uint32_t key1) {
    const long long per_rep = (long long)g.Ly * g.L;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const int rep = rep_first + r;
    const long long li = gid - (long long)r * per_rep;
    const int y = (int)(li / g.L), x = (int)(li - (long long)y * g.L);
    const long long i = (long long)(y + g.y_glob0) * g.L + x;
    uint32_t w[4];
    philox4x32_10((uint32_t)(i >> 2), (uint32_t)((unsigned long long)i >> 34), 0u, STREAM_INIT << 1, key0,
                  key1 ^ (uint32_t)rep, w);
    const uint32_t code = (uint32_t)(((unsigned long long)w[i & 3] * 3ull) >> 32);
    const size_t dst = ((size_t)rep * g.rows_arr + g.row_first + y) * g.Wc + (x >> 1);
    if ((x + y + g.y_glob0) & 1) c1[dst] = (uint8_t)code; else c0[dst] = (uint8_t)code;
}

// device-resident sweep counter used by CUDA-graph replay: kernels read t = *clock + offset
// Pair table of replicas [rep_first, rep_first + n): word a + 54*b = the 32-bit T with
//   (h_a | h_b << 16) - T  ==  (0x8000 + u15_a - E15_a) | (0x8000 + u15_b - E15_b) << 16   (mod 2^32)
// for coarse draws h = flip << 15 | u15 whose flip bits agree with the entry numbers a, b (flip = (i % 18) >= 9):
// lane = flip*0x8000 - 0x8000 + E15.  Both result lanes lie in [0, 0xFFFF], so they never borrow from each other.
__global__ void pair_table_kernel(const uint16_t* __restrict__ tab16, uint32_t* __restrict__ tab2, int rep_first, int n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * TAB2_WORDS) return;
    const int r = rep_first + (int)(i / TAB2_WORDS), w = (int)(i % TAB2_WORDS);
    const int a = w % TABLE_ENTRIES, b = w / TABLE_ENTRIES;
    const uint16_t* t = tab16 + (size_t)r * TABLE_ENTRIES;
    const uint32_t la = ((a % 18) >= 9 ? 0u : 0xFFFF8000u) + t[a];
    const uint32_t lb = ((b % 18) >= 9 ? 0u : 0xFFFF8000u) + t[b];
    tab2[(size_t)r * TAB2_WORDS + w] = la + (lb << 16);
}

__global__ void clock_set_kernel(int64_t* clock, int64_t value) { *clock = value; }
__global__ void clock_add_kernel(int64_t* clock, int64_t delta) { *clock += delta; }

__global__ void fill_kernel(uint8_t* __restrict__ p, size_t n, uint8_t v) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) p[gid] = v;
}

// single-rank periodic slab: ghost rows <- opposite owned rows of the same array (all replicas)
__global__ void halo_self_kernel(uint8_t* __restrict__ c, const Geom g, int n_rep) {
    const long long per_rep = 2LL * g.Wc;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const int rem = (int)(gid - (long long)r * per_rep);
    const int which = rem / g.Wc, k = rem - which * g.Wc;
    uint8_t* base = c + (size_t)r * g.rows_arr * g.Wc;
    if (which == 0) base[k] = base[(size_t)g.Ly * g.Wc + k];                              // top ghost <- last owned row
    else base[(size_t)(g.Ly + 1) * g.Wc + k] = base[(size_t)g.Wc + k];                    // bottom ghost <- first owned row
}

// ---- slab halo over peer memory (NVLink P2P, one process per GPU, pointers from cudaIpcOpenMemHandle) --------
// Copies this rank's first / last owned row of one colour array into the bottom / top ghost row of the
// ring neighbour above / below, for every replica, with plain 16-byte stores to the peer's memory.
__global__ void halo_push_kernel(const uint8_t* __restrict__ c, uint8_t* __restrict__ peer_prev,
                                 uint8_t* __restrict__ peer_next, const Geom g, int n_rep) {
    const int vec_per_row = g.Wc / 16;
    const long long per_rep = 2LL * vec_per_row;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const int rem = (int)(gid - (long long)r * per_rep);
    const int which = rem / vec_per_row, v = rem - which * vec_per_row;
    const size_t rep_off = (size_t)r * g.rows_arr * g.Wc;
    if (which == 0) {
        if (peer_prev)  // my first owned row -> bottom ghost row of the rank above
            reinterpret_cast<uint4*>(peer_prev + rep_off + (size_t)(g.Ly + 1) * g.Wc)[v] =
                reinterpret_cast<const uint4*>(c + rep_off + (size_t)g.Wc)[v];
    } else {
        if (peer_next)  // my last owned row -> top ghost row of the rank below
            reinterpret_cast<uint4*>(peer_next + rep_off)[v] =
                reinterpret_cast<const uint4*>(c + rep_off + (size_t)g.Ly * g.Wc)[v];
    }
}

// After the pushes of an exchange have completed (stream order): publish the epoch in the neighbours' flags.
// flags[0] of a rank = "top ghost rows valid through epoch" (written by the rank above),
// flags[1]           = "bottom ghost rows valid through epoch" (written by the rank below).
__global__ void halo_signal_kernel(long long* peer_prev_flags, long long* peer_next_flags, long long epoch) {
    __threadfence_system();
    if (peer_prev_flags) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(peer_prev_flags + 1), "l"(epoch) : "memory");
    if (peer_next_flags) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(peer_next_flags + 0), "l"(epoch) : "memory");
}

// Before ghost rows are read: wait until both neighbours have published `epoch`.  One thread, one GPU per
// rank (never two ranks on one GPU), bounded: gives up after ~4 s and raises *err instead of hanging the GPU.
__global__ void halo_wait_kernel(const long long* my_flags, int need_top, int need_bottom, long long epoch, int* err) {
    const long long t0 = clock64();
    for (int k = 0; k < 2; ++k) {
        if (!(k == 0 ? need_top : need_bottom)) continue;
        long long v;
        do {
            asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(my_flags + k) : "memory");
            if (v < epoch && clock64() - t0 > 8000000000LL) { *err = 1; return; }
        } while (v < epoch);
    }
}

// code of the neighbour (nx, ny) [ny = owned-row index, may be -1 or Ly] in the other colour's array
__device__ __forceinline__ uint32_t geom_other_code(const uint8_t* oth_rep, const Geom& g, int nx, int ny) {
    if (g.open) { if (nx < 0 || nx >= g.L) return 1u; }
    else nx = (nx < 0) ? nx + g.L : (nx >= g.L ? nx - g.L : nx);
    if (!g.ghost) {
        if (g.open) { if (ny < 0 || ny >= g.Ly) return 1u; }
        else ny = (ny < 0) ? ny + g.Ly : (ny >= g.Ly ? ny - g.Ly : ny);
    }
    return oth_rep[(size_t)(g.row_first + ny) * g.Wc + (nx >> 1)];
}

// standalone observables of the current configuration into raw accumulators [rep][OBS_RAW]
__global__ void measure_kernel(const uint8_t* __restrict__ c0, const uint8_t* __restrict__ c1, const Geom g,
                               int n_rep, unsigned long long* __restrict__ obs) {
    const long long per_rep = (long long)g.Ly * g.L;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int rep = (int)(gid / per_rep);
    const long long i = gid - (long long)rep * per_rep;
    const int y = (int)(i / g.L), x = (int)(i - (long long)y * g.L);
    const size_t rb = (size_t)rep * g.rows_arr * g.Wc;
    const int col = (x + y + g.y_glob0) & 1;
    const uint32_t c = (col ? c1 : c0)[rb + (size_t)(g.row_first + y) * g.Wc + (x >> 1)];
    uint32_t v[OBS_RAW] = {0, 0, 0, 0, 0, 0};
    v[2 * col + 0] = c;
    v[2 * col + 1] = c * c;
    if (col == 1) {
        const uint8_t* oth = c0 + rb;
        const uint32_t S = geom_other_code(oth, g, x + 1, y) + geom_other_code(oth, g, x - 1, y) +
                           geom_other_code(oth, g, x, y + 1) + geom_other_code(oth, g, x, y - 1);
        v[4] = c * S;
        v[5] = S;
    }
    const unsigned mask = __activemask();
    const int leader = __ffs(mask) - 1;
    const int rep_lo = __shfl_sync(mask, rep, leader);
    const bool uniform = __all_sync(mask, rep == rep_lo);
    unsigned long long* dst = obs + (size_t)rep * OBS_RAW;
#pragma unroll
    for (int k = 0; k < OBS_RAW; ++k) {
        if (uniform) {
            const uint32_t sum = __reduce_add_sync(mask, v[k]);
            if ((int)(threadIdx.x & 31) == leader && sum) atomicAdd(dst + k, (unsigned long long)sum);
        } else if (v[k]) {
            atomicAdd(dst + k, (unsigned long long)v[k]);
        }
    }
}

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
