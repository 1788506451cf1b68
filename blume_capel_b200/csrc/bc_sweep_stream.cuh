// bc_sweep_stream.cuh -- the streaming kernel (L % 32 == 0), its group launch for several lattice sizes and the experimental persistent variant.
// Part of the sm_100a Blume-Capel engine; included through bc_kernels.cuh.
#pragma once
#include "bc_update.cuh"

namespace bc {

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
// from a 27-word conflict-free table (PAIR: one LDS.32 from the 54 x 54 pair table per two sites), two sites per
// packed 16-bit subtract, reject masks by PRMT sign replication, ties (coarse draw == coarse threshold) found by a
// packed signed minimum (VIMNMX3.S16x2) and resolved with the fine stream -- inline, one test per 8 sites, in the
// per-site-table kernels; out of the per-site path, one warp vote per two rows, in the pair-table kernels
// (fix_tied_block).  The two half-rate integer pipes (ALU: LOP3/PRMT/SHF, FMA-heavy: IMAD/IDP, quarter-rate
// IMAD.WIDE) bound this kernel, not HBM; the instruction mix is balanced between them (tools/microbench/pipes.cu).
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

//   HALO  = true : half-sweep of a slab in peer-memory mode, ONE launch for the whole slab: the first CTAs of every
//                  replica hold the two boundary strips (first and last strip of the slab); they wait for the
//                  neighbours' ghost rows (epoch flags, ld.acquire.sys), the thread that produces a chunk of the
//                  slab's first / last owned row also stores it into the bottom / top ghost row of the rank above /
//                  below (peer pointers over NVLink), and the last boundary CTA to finish publishes the next epoch
//                  in the neighbours' flags (st.release.sys).  The interior strips follow in the same grid and
//                  overlap the transfer.  Compute, halo transfer and synchronisation are one kernel, so a slab
//                  half-sweep is graph-replayed and programmatically chained exactly like a plain one.
//   PAIR >= 1    : (needs !SMALL) the CTA also stages its replica's 54 x 54 pair table (11.4 KB) and looks the
//                  thresholds of two sites up with one IDP.4A + one LDS.32; pays when a CTA marches over enough
//                  rows to amortise the staging (host: LaunchPlan::pair).
// (Round 1 and the first half of round 2 also had PAIR == 2, software-pipelined Philox draws at 80 registers / 6 CTAs per
// SM; with the branch-free row loop below the plain pair-table kernel at 72 registers / 7 CTAs per SM is faster at
// every strip height -- 1917 against 1840 updates/ns on the bench batch -- and the variant was removed.)
// DEFER: tie repair out of the per-site path (fix_tied_block) in the pair-table kernels.  The per-site-table kernels
// (strips of fewer than 8 rows: small, launch-bound batches) keep the inline repair: with the out-of-line one they
// gain 3..5 % where they are throughput-bound (64 x L=4096 at 4 rows per thread: 1427 -> 1500 updates/ns) but a
// single L=2048 lattice, 3.5 us per launch, loses 11 % to the larger kernel image and prologue.
template <int PAIR>
struct TieDefer { static constexpr bool value = PAIR > 0; };
template <bool SMALL, bool MEASURE, bool OPEN, int PAIR, bool HALO = false>
struct FastMinBlocks {
    static constexpr int value =
        (MEASURE || SMALL) ? BC_FAST_MIN_BLOCKS : BC_FAST_MIN_BLOCKS + 1;
};

// HALO launches: CTA order inside a replica.  A replica has C = items / 128 CTAs in item order (strip-major); the
// nb = max(1, cpr / 128) CTAs that hold its first strip and the nb CTAs that hold its last strip are moved to the
// front of the replica's CTA range (they are scheduled first, so their halo rows travel while the interior is
// computed).  Returns the item-order CTA index for launch-order index `bid`; *boundary = CTA holds a boundary strip.
// All uniform-datapath arithmetic on the block index (permuting the strip index per thread instead cost the row loop
// of the periodic kernels 64 bytes of spills).
__device__ __forceinline__ unsigned halo_cta_order(const SweepParams& P, const unsigned bid, bool* boundary) {
    const unsigned cpr = (unsigned)P.cpr, C = (unsigned)P.strips * cpr / FAST_THREADS;
    const unsigned nb = cpr >= FAST_THREADS ? cpr / FAST_THREADS : 1u;
    const unsigned nh = 2u * nb < C ? 2u * nb : C;
    const unsigned rep = bid / C, b = bid - rep * C;
    *boundary = b < nh;
    const unsigned item = b < nb ? b : (b < nh ? C - (nh - nb) + (b - nb) : b - (nh - nb));
    return rep * C + item;
}

// Epilogue of a HALO launch (see sweep_fast_body): halo transfer and epoch signal of the boundary CTAs.  Everything
// is recomputed from laundered thread / block ids so that nothing has to stay live across the row loop.
__device__ __forceinline__ void halo_epilogue(const SweepParams& P) {
    uint32_t tid, bid0;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bid0));
    bool boundary;
    const unsigned bid = halo_cta_order(P, bid0, &boundary);
    if (!boundary) return;  // CTA-uniform
    const uint32_t cpr = (uint32_t)P.cpr, ipr = (uint32_t)P.strips * cpr;
    const unsigned long long gid0 = (unsigned long long)bid * FAST_THREADS;
    const uint32_t rep = (uint32_t)(gid0 / ipr);
    const uint32_t rem = (uint32_t)(gid0 - (unsigned long long)rep * ipr) + tid;
    const uint32_t s = rem / cpr, j = rem - s * cpr;
    const uint32_t Wc = (uint32_t)P.Wc;
    const size_t rep_base = (size_t)rep * (size_t)P.rows_arr * Wc;
    // Fused halo transfer: the thread that produced a chunk of the slab's first owned row copies it into the bottom
    // ghost row of the rank above, the one that produced a chunk of the last owned row into the top ghost row of
    // the rank below (same array geometry on every rank; a thread reads back its own stores).
    if (P.peer_prev && s == 0u)
        *reinterpret_cast<uint4*>(P.peer_prev + rep_base + (size_t)(P.Ly + 1) * Wc + (size_t)j * 16) =
            *reinterpret_cast<const uint4*>(P.own + rep_base + (size_t)P.row_first * Wc + (size_t)j * 16);
    if (P.peer_next && s == (uint32_t)P.strips - 1u)
        *reinterpret_cast<uint4*>(P.peer_next + rep_base + (size_t)j * 16) =
            *reinterpret_cast<const uint4*>(P.own + rep_base + (size_t)(P.row_first + P.Ly - 1) * Wc + (size_t)j * 16);
    // every boundary CTA: its halo stores are issued; the last one to get here publishes the next epoch
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        const unsigned total = (unsigned)P.halo_ctas_per_rep * (unsigned)P.n_replicas;
        if (atomicAdd(P.halo_done, 1u) + 1u == total) {
            __threadfence_system();
            *P.halo_done = 0u;
            const long long epoch = *reinterpret_cast<volatile long long*>(P.halo_clock) + 1;
            *reinterpret_cast<volatile long long*>(P.halo_clock) = epoch;
            if (P.peer_prev_flags) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(P.peer_prev_flags + 1), "l"(epoch) : "memory");
            if (P.peer_next_flags) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(P.peer_next_flags + 0), "l"(epoch) : "memory");
        }
    }
}

// Tie repair (cold, warp-cooperative).  The row loop of the pair-table kernels is branch-free per site: a site whose
// coarse compare ties (coarse draw == coarse threshold, about 2^-15 per uphill attempt) is REJECTED there and only
// lowers a running minimum; once per block of B rows (B = 2 in the unrolled loops, else 1) the warp votes on that
// minimum and, if any lane saw a tie, runs this function.  The warp takes its flagged threads one by one and
// re-evaluates the 16 B sites of the flagged thread's block, one site per lane, with the scalar form of the rule --
// exactly the per-site function of the spec (coarse compare, tie -> u31 = (u15 << 16 | f16) < thr31 with
// f16 from the fine stream) -- on the site's ORIGINAL value (the flagged thread's own-chunk staging slots, which
// still hold the block's rows; the stored row cannot tell a rejected tie from an accepted move) and the unchanged
// other colour (global memory, read a microsecond ago: L2 hits), and rewrites the byte if the exact rule accepts.
// Sites without a tie are left untouched.  Geometry and Philox context are recomputed from the launch parameters,
// so nothing stays live in the row loop for this path.  r0 = first strip row of the block, slot0 = own-ring slot of
// that row, own_slots = the CTA's own ring [4][FAST_THREADS] x 16 bytes; tab = coarse table + exact thresholds
// (shared memory); acc = {sum c, sum c^2, sum c*S, ties} of the calling lane.
template <int COL, bool MEASURE, bool OPEN, bool HALO>
__device__ __forceinline__ void fix_tied_block(const SweepParams& P, const unsigned bid_launch, const int64_t t_add,
                                               const bool tied, const uint32_t r0, const uint32_t B, const uint32_t slot0,
                                               const uint8_t* own_slots, const uint32_t* tab, uint32_t* acc) {
    const unsigned flagged = __ballot_sync(0xFFFFFFFFu, tied);
    __syncwarp();  // the helper lanes' byte stores come after the flagged threads' row stores
    bool halo_cta;
    const unsigned bid = HALO ? halo_cta_order(P, bid_launch, &halo_cta) : bid_launch;
    const uint32_t ipr = (uint32_t)(P.strips * P.cpr), cpr = (uint32_t)P.cpr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t rows = (uint32_t)P.rows_per_strip, Wc = (uint32_t)P.Wc, total = (uint32_t)P.rows_arr * Wc;
    const uint16_t* tab16 = reinterpret_cast<const uint16_t*>(tab);  // shared memory: coarse thresholds ...
    const uint32_t* tab31 = tab + TAB16_WORDS;                       // ... and the exact ones behind them
    int64_t t = P.t + t_add;
    if (P.t_ptr) t += *P.t_ptr;
    const uint32_t ctr2 = (uint32_t)((uint64_t)t & 0xFFFFFFFFu);
    // the CTA lies inside one replica (pair-table kernels: items per replica % 128 == 0)
    const unsigned long long gid0 = (unsigned long long)bid * FAST_THREADS;
    const uint32_t rep = (uint32_t)(gid0 / ipr);
    const uint32_t rem0 = (uint32_t)(gid0 - (unsigned long long)rep * ipr) + (threadIdx.x & ~31u);
    const size_t rep_base = (size_t)rep * (size_t)P.rows_arr * (size_t)Wc;
    const uint8_t* oth_rep = P.other + rep_base;
    const uint32_t k0 = P.key0, k1 = P.key1 ^ rep;
    const uint32_t rr = lane >> 4, site = lane & 15u;                // lane -> (row of the block, site of the chunk)
    if (rr < B) {
        for (unsigned todo = flagged; todo; todo &= todo - 1u) {
            const uint32_t src = (uint32_t)__ffs(todo) - 1u;
            // the flagged thread's work item -> (strip, chunk), as in sweep_fast_body
            const uint32_t rem = rem0 + src, strip = rem / cpr, j = rem - strip * cpr;
            const uint32_t y = (uint32_t)P.row_first + ((uint32_t)P.strip_base + strip * (uint32_t)P.strip_stride) * rows + r0 + rr;
            const uint32_t yg = y + (uint32_t)P.y_rng_off, o = y * Wc;
            const int par = (int)((yg + COL) & 1u);
            const uint8_t* oth_col = oth_rep + (size_t)j * 16;
            const uint32_t c = own_slots[((((slot0 + rr) & 3u) * FAST_THREADS) + (threadIdx.x & ~31u) + src) * 16u + site];
            const uint32_t up = o == 0 ? (OPEN ? 1u : oth_col[total - Wc + site]) : oth_col[o - Wc + site];
            const uint32_t dn = o + Wc == total ? (OPEN ? 1u : oth_col[site]) : oth_col[o + Wc + site];
            uint32_t hor;  // the horizontal neighbour that is not at the same k: k+1 for par = 1, k-1 for par = 0
            if (par) hor = site < 15u ? oth_col[o + site + 1u]
                                      : ((OPEN && j + 1 == cpr) ? 1u : oth_rep[o + ((j + 1 == cpr) ? 0u : j + 1) * 16]);
            else hor = site > 0u ? oth_col[o + site - 1u]
                                 : ((OPEN && j == 0) ? 1u : oth_rep[o + ((j == 0) ? cpr - 1 : j - 1) * 16 + 15]);
            const uint32_t S = up + dn + oth_col[o + site] + hor;      // main.cpp:97-103 in code space
            uint32_t q[4];
            philox4x32_10(2u * j + (site >> 3), yg, ctr2, ctr3_of(t, STREAM_COARSE, COL), k0, k1, q);
            const uint32_t s8 = site & 7u;
            const uint32_t rw = (s8 >> 1) == 0 ? q[0] : ((s8 >> 1) == 1 ? q[1] : ((s8 >> 1) == 2 ? q[2] : q[3]));
            const uint32_t h = (s8 & 1u) ? (rw >> 16) : (rw & 0xFFFFu);
            const uint32_t fl = h >> 15;
            const uint32_t e = 18u * c + 9u * fl + S;                  // table entry
            if ((h & 0x7FFFu) != tab16[e]) continue;                   // no tie at this site
            philox4x32_10(2u * j + (site >> 3), yg, ctr2, ctr3_of(t, STREAM_FINE, COL), k0, k1, q);
            const uint32_t fw = (s8 >> 1) == 0 ? q[0] : ((s8 >> 1) == 1 ? q[1] : ((s8 >> 1) == 2 ? q[2] : q[3]));
            const uint32_t f16 = (s8 & 1u) ? (fw >> 16) : (fw & 0xFFFFu);
            ++acc[3];
            if (exact_accept(h, f16, tab31[e])) {
                const uint32_t cn = c ^ (((2u * c + fl) % 3u) + 1u);  // proposal, main.cpp:94 in code space
                (P.own + rep_base + (size_t)j * 16)[o + site] = (uint8_t)cn;
                if (MEASURE) {
                    acc[0] += cn - c;
                    acc[1] += cn * cn - c * c;
                    if (COL == 1) acc[2] += (cn - c) * S;
                }
            }
        }
    }
    __syncwarp();  // the flagged threads read their rows back (halo transfer)
}

// The kernel body; `bid` is the CTA's index among the CTAs that work on P's lattices (blockIdx.x for an ordinary
// launch, blockIdx.x minus the group's first CTA for a group launch, see sweep_group_kernel).
template <int COL, bool SMALL, bool MEASURE, bool OPEN, bool HALO, int PAIR>
__device__ __forceinline__ void sweep_fast_body(const SweepParams& P, const unsigned bid_launch, const int64_t t_add = 0) {
    bool halo_cta = false;  // HALO: this CTA holds (part of) a boundary strip
    const unsigned bid = HALO ? halo_cta_order(P, bid_launch, &halo_cta) : bid_launch;
    static_assert(!(PAIR && SMALL) && !(HALO && SMALL), "the pair table and the slab halo logic are one replica per CTA");
    // coarse table(s); the pair-table kernels keep the replica's exact thresholds behind it for the tie repair
    __shared__ __align__(16) uint32_t s_tab[(SMALL ? FAST_MAX_REPL_PER_CTA : 1) * TAB16_WORDS + (PAIR ? TABLE_ENTRIES : 0)];
    __shared__ __align__(16) uint32_t s_tab2[PAIR ? TAB2_WORDS : 4];
    __shared__ uint4 s_rows[4][FAST_THREADS];    // thread-private ring of other-colour rows
    // DEEP: inputs are requested two rows ahead instead of one (the row above is then carried in registers and
    // the ring holds rows y, y+1 and the two rows in flight).  Measured (tools/tune_libs2.sh): +1..3 % for the
    // periodic pair-table kernels at 8..16 rows per thread, -1 % for the pipelined kernel, -2 % with open edges.
    constexpr bool DEEP = BC_DEEP_PREFETCH == 2 || (BC_DEEP_PREFETCH == 1 && PAIR == 1 && !OPEN);
    // DEFER (the pair-table kernels): ties are repaired out of the per-site path, once per block of two rows, from the
    // ORIGINAL own chunks of the block -- the own ring then keeps four rows in every variant (see end_block)
    constexpr bool DEFER = TieDefer<PAIR>::value;
    constexpr int OM = (DEEP || DEFER) ? 3 : 1;              // own / side ring: slot of strip row r = r & OM
    __shared__ uint4 s_own[OM + 1][FAST_THREADS];            // thread-private own chunks (current / next rows)
    __shared__ uint32_t s_side[OM + 1][FAST_THREADS];        // thread-private side words

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
    if (PAIR && threadIdx.x >= FAST_THREADS - TABLE_ENTRIES)
        s_tab[TAB16_WORDS + threadIdx.x - (FAST_THREADS - TABLE_ENTRIES)] =
            P.tab31[(size_t)rep0 * TABLE_ENTRIES + threadIdx.x - (FAST_THREADS - TABLE_ENTRIES)];
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
    // One row.  DEFER (the pair-table kernels): branch-free update; a tie (rare) only lowers the running minimum
    // `tiemin`; once per block of rows (4 rows in the unrolled loops, 1 otherwise) the warp votes on it and repairs
    // the block out of line (fix_tied_block).  Otherwise (short strips, SMALL): ties are resolved inline, one test
    // per 8 sites.
    uint32_t tiemin = TIEMIN_INIT;
    auto update_row = [&](const uint32_t y, const uint32_t off, const int par, const uint4& upv, const uint4& midv,
                          const uint4& dnv, const uint32_t side, const uint4& ownv) {
        if (DEFER) {
            *reinterpret_cast<uint4*>(own_col + off) =
                update_chunk16_bf<COL, MEASURE, (PAIR > 0)>(ctx, y, par, upv, midv, dnv, side, ownv, a_c, a_c2, a_cS, a_S, tiemin);
        } else {
            *reinterpret_cast<uint4*>(own_col + off) =
                update_chunk16<COL, MEASURE, (PAIR > 0)>(ctx, y, par, upv, midv, dnv, side, ownv, a_c, a_c2, a_cS, a_S, n_ties);
        }
    };
    // r0: first strip row of the block of `nrows` rows just stored; slot0: own-ring slot of row r0
    auto end_block = [&](const uint32_t r0, const uint32_t nrows, const uint32_t slot0) {
        if (!DEFER) return;
        const bool tied = has_tie(tiemin);
        if (__any_sync(0xFFFFFFFFu, tied)) {
            uint32_t acc[4] = {a_c, a_c2, a_cS, 0u};
            fix_tied_block<COL, MEASURE, OPEN, HALO>(P, bid_launch, t_add, tied, r0, nrows, slot0,
                                                     reinterpret_cast<const uint8_t*>(&s_own[0][0]), s_tab, acc);
            a_c = acc[0]; a_c2 = acc[1]; a_cS = acc[2]; n_ties += acc[3];
        }
        tiemin = TIEMIN_INIT;
    };

    // ---- row marching.  Thread-private staging slots in shared memory (a ring of four other-colour
    // ---- rows, two or four own chunks and side words) filled by cp.async one or two rows ahead: the inputs of
    // ---- the next rows are in flight while row y is computed, and they cost no registers while in flight.
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
    if (HALO && halo_cta) {
        // boundary strips read ghost rows: wait until both neighbours have published the epoch this rank is at
        // (bounded: raises *halo_err instead of hanging the GPU; the run is then reported as failed by the host)
        if (threadIdx.x == 0) {
            const long long epoch = *reinterpret_cast<volatile long long*>(P.halo_clock);
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (int k = 0; k < 2; ++k) {
                if (!((P.halo_need >> k) & 1)) continue;
                long long v;
                for (;;) {
                    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(P.halo_flags + k) : "memory");
                    if (v >= epoch) break;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if ((long long)(t1 - t0) > P.halo_timeout_ns) { *P.halo_err = 1; break; }
                }
            }
        }
        __syncthreads();
    }
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
                stage_row_inputs((sl + 1) & OM, off + Wc, 1 - par);
            }
            cp_async_commit();
            cp_async_wait<1>();  // everything but the group just committed has landed: row r is ready
            const uint4 upv = s_rows[sl & 3][tid], midv = s_rows[(sl + 1) & 3][tid], dnv = s_rows[(sl + 2) & 3][tid];
            const uint4 ownv = s_own[sl & OM][tid];
            const uint32_t side = s_side[sl & OM][tid];
            update_row(yg0 + r, off, par, upv, midv, dnv, side, ownv);
            off += Wc;
        };

        if (SMALL || (rows & 3u)) {
            for (uint32_t r = 0; r < rows; ++r) { march(r, (int)((yg0 + r + COL) & 1u), r + 1 == rows, (int)(r & 3u)); end_block(r, 1u, r & 3u); }
        } else {
            // yg0 is even and rows is a multiple of 4: parities alternate COL, 1-COL statically; static ring slots
            for (uint32_t r = 0; r < rows; r += 4) {
                march(r, COL, false, 0);
                march(r + 1, 1 - COL, false, 1);
                end_block(r, 2u, 0u);
                march(r + 2, COL, false, 2);
                march(r + 3, 1 - COL, r + 4 == rows, 3);
                end_block(r + 2, 2u, 2u);
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
            for (uint32_t r = 0; r < rows; ++r) { march(r, (int)((yg0 + r + COL) & 1u), r + 3 <= rows, (int)(r & 3u)); end_block(r, 1u, r & 3u); }
        } else {
            for (uint32_t r = 0; r < rows; r += 4) {
                const bool more = r + 4 != rows;
                march(r, COL, true, 0);
                march(r + 1, 1 - COL, true, 1);
                end_block(r, 2u, 0u);
                march(r + 2, COL, more, 2);
                march(r + 3, 1 - COL, more, 3);
                end_block(r + 2, 2u, 2u);
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
    if (HALO) halo_epilogue(P);
}

template <int COL, bool SMALL, bool MEASURE, bool OPEN, bool HALO = false, int PAIR = 0>
__global__ void __launch_bounds__(FAST_THREADS, FastMinBlocks<SMALL, MEASURE, OPEN, PAIR, HALO>::value)
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
__global__ void __launch_bounds__(FAST_THREADS, FastMinBlocks<false, MEASURE, OPEN, PAIR>::value)
sweep_group_kernel(const __grid_constant__ GroupParams G) {
    int g = 0;
#pragma unroll
    for (int i = 0; i < GROUP_MAX - 1; ++i)
        if (i + 1 < G.n && blockIdx.x >= G.cta_end[i]) g = i + 1;
    const unsigned first = g ? G.cta_end[g - 1] : 0u;
    sweep_fast_body<COL, false, MEASURE, OPEN, false, PAIR>(G.p[g], blockIdx.x - first);
}

#ifdef BC_EXPERIMENTAL
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
#endif  // BC_EXPERIMENTAL

}  // namespace bc
