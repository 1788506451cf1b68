// bc_sweep_resident.cuh -- shared-memory-resident kernel for small lattices of any size (all sweeps of a call in one launch).
// Part of the sm_100a Blume-Capel engine; included through bc_kernels.cuh.
#pragma once
#include "bc_update.cuh"
#include <type_traits>

namespace bc {

// Register cap of the resident kernel.  Left to __launch_bounds__(512) alone the kernel takes 112 registers: four CTAs
// of 128 threads per SM; this kernel is one work item per thread and a barrier per half-sweep, i.e. latency-bound,
// and lives on resident warps.  96 registers (five CTAs per SM, no spills): L=64 x 4096 replicas 1114 -> 1256
// updates/ns, measured every sweep 994 -> 1119, L=32 +13 %, L=128 +7 %, one L=64 replica 3.0 -> 3.2; 80 registers
// (six CTAs, a few spills): 1181; 64 registers (eight CTAs, 60..80 spill instructions): 1116
// (profiles/r02_tune_resident.jsonl).
#ifndef BC_RESIDENT_MAXNREG
#define BC_RESIDENT_MAXNREG 96
#endif
// The instantiation for calls that measure nothing (CANMEAS = false) fits 80 registers without spills: six CTAs per SM,
// 1256 -> 1293 at L=64, 1142 -> 1230 at L=128 (72 registers: 1251 / 1165; 64: 1178 / 1112).
#ifndef BC_RESIDENT_MAXNREG_PLAIN
#define BC_RESIDENT_MAXNREG_PLAIN 80
#endif

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
    const uint32_t* tab2;    // [replica][54*54] pair tables (PAIR instantiations) or nullptr
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
//   CANMEAS = false: a call that measures nothing (burn-in, unmeasured stretches): the measuring half-sweeps and the
//                reduction are not compiled in, which frees registers for a sixth CTA per SM.
//   PAIR = true  : (one replica per CTA, calls of many sweeps) the replica's 54 x 54 pair table (11.4 KB) is staged
//                once per launch behind the lattice and two thresholds are looked up with one IDP.4A + LDS.32, as in
//                the streaming kernel.
template <bool OPEN, bool ANY = false, bool CANMEAS = true, bool PAIR = false>
__global__ void __maxnreg__(CANMEAS ? BC_RESIDENT_MAXNREG : BC_RESIDENT_MAXNREG_PLAIN)
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
    uint32_t* s_tab2 = reinterpret_cast<uint32_t*>(s_dyn + (size_t)P.reps_per_cta * (rep_stride + 32));  // PAIR (reps_per_cta == 1)

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
        if (PAIR) {
            const uint4* g2 = reinterpret_cast<const uint4*>(P.tab2 + (size_t)rep * TAB2_WORDS);
            for (int i = ltid; i < TAB2_WORDS / 4; i += T) reinterpret_cast<uint4*>(s_tab2)[i] = g2[i];
        }
    }
    __syncthreads();

    RowCtx ctx;
    ctx.k0 = P.key0; ctx.k1 = P.key1 ^ (uint32_t)rep;
    ctx.tabm = reinterpret_cast<const char*>(s_tab) - 18;
    ctx.t31_rep = P.tab31 + (size_t)(live ? rep : 0) * TABLE_ENTRIES;
    ctx.tab2s = PAIR ? (uint32_t)__cvta_generic_to_shared(s_tab2) - TAB2_BIAS : 0u;
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
                    update_chunk16<COL, MEAS, PAIR>(ctx, (uint32_t)y, par, upv, midv, dnv, side, ownv, a_c, a_c2, a_cS, a_S, n_ties);
            } else {
                uint32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;
                const uint4 nv = update_chunk16<COL, false, PAIR>(ctx, (uint32_t)y, par, upv, midv, dnv, side, ownv, d0, d1, d2, d3, n_ties);
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
        const bool measure = CANMEAS && to_meas == 0;
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

}  // namespace bc
