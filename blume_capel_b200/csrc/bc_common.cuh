// bc_common.cuh -- shared definitions for the sm_100a Blume-Capel engine.
//
// Bit-exact spec.  Reference rule being restated:
// /root/reference/main.cpp:94 (proposal), :97-103 (neighbour sum), :105-106 (dE), :109 (accept).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bc {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

constexpr uint32_t STREAM_COARSE = 0u;
constexpr uint32_t STREAM_FINE = 1u;
constexpr uint32_t STREAM_INIT = 2u;
constexpr uint32_t STREAM_BOND = 3u;
constexpr uint32_t STREAM_FLIP = 4u;

constexpr int TABLE_ENTRIES = 54;        // 18*c + 9*flip + S
constexpr int TAB16_WORDS = 27;          // 54 x u16 -> 27 words: every entry in its own bank
// Pair table of the streaming kernel (one lookup per two sites): word a + 54*b for the entry numbers a (low
// halfword of the random word) and b (high halfword); 11664 bytes per replica.  The kernel's entry numbers
// carry a bias of +9 (f+1 instead of f), folded into the base address.
constexpr int TAB2_WORDS = TABLE_ENTRIES * TABLE_ENTRIES;
constexpr uint32_t TAB2_BIAS = 9u * 4u + 9u * 4u * TABLE_ENTRIES;
constexpr uint32_t ONES = 0x01010101u;   // code 1 == spin 0 in every byte

// Philox4x32-10 with a runtime key.  The multiply pairs compile to IMAD.WIDE.U32.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ (k0 + (uint32_t)r * PHILOX_W0);
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ (k1 + (uint32_t)r * PHILOX_W1);
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ uint32_t ctr3_of(int64_t t, uint32_t stream, int colour) {
    return ((uint32_t)((uint64_t)t >> 32) << 4) | (stream << 1) | (uint32_t)(colour & 1);
}

// Raw per-(measurement, replica) accumulators written by the sweep kernels (all exact integers):
//   [0] sum c  over colour-0 sites   [1] sum c^2 over colour-0 sites
//   [2] sum c  over colour-1 sites   [3] sum c^2 over colour-1 sites
//   [4] sum c*S over colour-1 sites  [5] sum S  over colour-1 sites        (c = s+1, S = nb+4)
// M = A0+A2-N, Q = (A1+A3) - 2(A0+A2) + N, B = A4 - 4*A2 - A5 + 4*N1  (N1 = #colour-1 sites).
constexpr int OBS_RAW = 6;

struct SweepParams {
    uint8_t* own;            // colour array being updated   [replica][y][Wc]
    const uint8_t* other;    // the other colour's array      [replica][y][Wc]
    const uint16_t* tab16;   // [replica][54] coarse thresholds incl. flip in bit 15 (see engine)
    const uint32_t* tab31;   // [replica][54] full thresholds (tie resolution)
    const uint32_t* tab2;    // [replica][54*54] pair table (PAIR kernels) or nullptr
    unsigned long long* obs; // raw accumulators [slot][replica][OBS_RAW] or nullptr
    unsigned long long* ties;// tie counter (diagnostics)
    const int64_t* t_ptr;    // optional device-resident sweep counter (graph replay); t = *t_ptr + t
    const int64_t* slot_ptr; // optional device-resident obs slot base
    int64_t t;               // sweep index (or offset to *t_ptr)
    int64_t slot;            // obs slot (or offset to *slot_ptr)
    int L;                   // lattice size (columns; also rows unless this is a slab)
    int rows_arr;            // rows of one replica's colour array (L, or local rows + 2 ghost rows for a slab)
    int row_first;           // first array row that is updated (0, or 1 when ghost rows are present)
    int rows_per_strip;      // rows each thread of the fast kernel marches over
    int y_rng_off;           // global row = array row + y_rng_off (RNG counter and colour parity)
    int Wc;                  // compact row stride in bytes
    int cpr;                 // 16-byte chunks per compact row (fast kernel)
    int strips;              // row strips per replica processed by THIS launch
    int strip_base;          // actual strip = strip_base + local strip * strip_stride (a slab launches its two
    int strip_stride;        //   boundary strips first, then the interior, so the halo exchange can overlap)
    int n_replicas;
    int open;                // 1 = open boundary (off-lattice neighbours are spin 0)
    uint32_t key0, key1;     // Philox key of replica 0; replica r uses (key0, key1 ^ r)
    int Ly;                  // owned rows (== rows_arr unless ghost rows are present)
    uint8_t* peer_prev;      // HALO kernels: the rank above's array of the colour being updated (peer memory);
    uint8_t* peer_next;      //   the rank below's.  nullptr at an open edge.
    // HALO kernels (peer-memory slabs): the whole half-sweep of a slab is ONE launch whose first CTAs hold the two
    // boundary strips.  Those CTAs wait for the neighbours' ghost rows (epoch flags), store their fresh first / last
    // rows into the neighbours' ghost rows, and the last of them to finish publishes the next epoch.
    const long long* halo_flags;  // this rank's flags: [0] top ghost rows valid through epoch, [1] bottom
    long long* halo_clock;        // epochs this rank has published so far (== what the neighbours must have published)
    unsigned* halo_done;          // boundary CTAs of this launch that have finished (reset by the last one)
    int* halo_err;                // raised when a neighbour did not deliver in time
    long long* peer_prev_flags;   // the rank above's flags (its [1] is ours to write), nullptr at an open edge
    long long* peer_next_flags;   // the rank below's flags (its [0] is ours to write)
    long long halo_timeout_ns;
    int halo_ctas_per_rep;        // CTAs per replica that hold boundary-strip items (the first ones of the replica)
    int halo_need;                // bit 0: wait for the top ghost rows, bit 1: for the bottom ghost rows
};

}  // namespace bc
