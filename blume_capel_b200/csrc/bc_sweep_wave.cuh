// bc_sweep_wave.cuh -- experimental wavefront kernel (all half-sweeps in one launch via per-block progress counters).
// Part of the sm_100a Blume-Capel engine; included through bc_kernels.cuh.
#pragma once
#include "bc_sweep_stream.cuh"

namespace bc {

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

}  // namespace bc
