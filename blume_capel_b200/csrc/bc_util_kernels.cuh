// bc_util_kernels.cuh -- layout conversion, initialisation, pair tables, slab halo exchange, stand-alone measurement.
// Part of the sm_100a Blume-Capel engine; included through bc_kernels.cuh.
#pragma once
#include "bc_common.cuh"

namespace bc {

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

// ---- 2-bit wire format ---------------------------------------------------------------------------------------
// Host <-> device transfers of whole lattices are PCIe-bound (one byte per site each way); the wire format packs four
// sites into a byte: row-major like the reference's lattice (main.cpp:98-101), byte b of a row holds the codes
// (spin + 1) of sites x = 4b .. 4b+3 in bits 0-1, 2-3, 4-5, 6-7; rows are padded to whole bytes.  Conversion to and
// from the compact colour arrays happens on the device.  Vector forms for L % 64 == 0: one thread converts 64
// consecutive sites of a row = two 16-byte chunks of each colour array <-> 16 packed bytes.
__device__ __forceinline__ uint32_t pack2_word(uint32_t ev, uint32_t od) {  // 4 even + 4 odd site codes -> 2 bytes (low half)
    const uint32_t t = ev | (od << 2);          // per byte: two sites (even, odd) in 4 bits
    const uint32_t u = t | (t >> 4);            // bytes 0 and 2: four sites each
    return __byte_perm(u, 0u, 0x4420);          // (byte 0, byte 2, 0, 0)
}
__device__ __forceinline__ void unpack2_half(uint32_t two_bytes, uint32_t* ev, uint32_t* od) {  // inverse of pack2_word
    const uint32_t u = __byte_perm(two_bytes, 0u, 0x4140);       // byte 0 -> byte 0, byte 1 -> byte 2
    const uint32_t t = (u | (u << 4)) & 0x0F0F0F0Fu;              // one (even, odd) pair per byte
    *ev = t & 0x03030303u;
    *od = (t >> 2) & 0x03030303u;
}

__global__ void __launch_bounds__(256)
pack2_kernel(uint8_t* __restrict__ packed, const uint8_t* __restrict__ c0, const uint8_t* __restrict__ c1, const Geom g, int rep_first) {
    const uint32_t per_row = (uint32_t)g.L >> 6;
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= (uint32_t)g.Ly * per_row) return;
    const uint32_t r = blockIdx.y, y = idx / per_row, m = idx - y * per_row;
    const size_t src = ((size_t)(rep_first + r) * g.rows_arr + g.row_first + y) * g.Wc + (size_t)m * 32;
    const bool odd_row = ((y + (uint32_t)g.y_glob0) & 1u) != 0;  // colour 0 holds x = 2k + (row & 1)
    const uint8_t* ev = (odd_row ? c1 : c0) + src;
    const uint8_t* od = (odd_row ? c0 : c1) + src;
    const uint4 e0 = *reinterpret_cast<const uint4*>(ev), e1 = *reinterpret_cast<const uint4*>(ev + 16);
    const uint4 o0 = *reinterpret_cast<const uint4*>(od), o1 = *reinterpret_cast<const uint4*>(od + 16);
    uint4 out;
    out.x = pack2_word(e0.x, o0.x) | (pack2_word(e0.y, o0.y) << 16);
    out.y = pack2_word(e0.z, o0.z) | (pack2_word(e0.w, o0.w) << 16);
    out.z = pack2_word(e1.x, o1.x) | (pack2_word(e1.y, o1.y) << 16);
    out.w = pack2_word(e1.z, o1.z) | (pack2_word(e1.w, o1.w) << 16);
    __stcs(reinterpret_cast<uint4*>(packed + ((size_t)r * g.Ly + y) * ((size_t)g.L >> 2) + (size_t)m * 16), out);
}

__global__ void __launch_bounds__(256)
unpack2_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ c0, uint8_t* __restrict__ c1, const Geom g, int rep_first) {
    const uint32_t per_row = (uint32_t)g.L >> 6;
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= (uint32_t)g.Ly * per_row) return;
    const uint32_t r = blockIdx.y, y = idx / per_row, m = idx - y * per_row;
    const uint4 in = __ldcs(reinterpret_cast<const uint4*>(packed + ((size_t)r * g.Ly + y) * ((size_t)g.L >> 2) + (size_t)m * 16));
    uint4 e0, e1, o0, o1;
    unpack2_half(in.x, &e0.x, &o0.x); unpack2_half(in.x >> 16, &e0.y, &o0.y);
    unpack2_half(in.y, &e0.z, &o0.z); unpack2_half(in.y >> 16, &e0.w, &o0.w);
    unpack2_half(in.z, &e1.x, &o1.x); unpack2_half(in.z >> 16, &e1.y, &o1.y);
    unpack2_half(in.w, &e1.z, &o1.z); unpack2_half(in.w >> 16, &e1.w, &o1.w);
    const size_t dst = ((size_t)(rep_first + r) * g.rows_arr + g.row_first + y) * g.Wc + (size_t)m * 32;
    const bool odd_row = ((y + (uint32_t)g.y_glob0) & 1u) != 0;
    uint8_t* ev = (odd_row ? c1 : c0) + dst;
    uint8_t* od = (odd_row ? c0 : c1) + dst;
    *reinterpret_cast<uint4*>(ev) = e0; *reinterpret_cast<uint4*>(ev + 16) = e1;
    *reinterpret_cast<uint4*>(od) = o0; *reinterpret_cast<uint4*>(od + 16) = o1;
}

// any L: one thread per packed byte (four sites; rows padded to whole bytes)
__global__ void pack2_any_kernel(uint8_t* __restrict__ packed, const uint8_t* __restrict__ c0, const uint8_t* __restrict__ c1, const Geom g,
                                 int rep_first, int n_rep) {
    const long long row_bytes = (g.L + 3) >> 2, per_rep = (long long)g.Ly * row_bytes;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const long long rem = gid - (long long)r * per_rep;
    const int y = (int)(rem / row_bytes), b = (int)(rem - (long long)y * row_bytes);
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) {
        const int x = 4 * b + i;
        uint32_t code = 1u;
        if (x < g.L) {
            const size_t src = ((size_t)(rep_first + r) * g.rows_arr + g.row_first + y) * g.Wc + (x >> 1);
            code = ((x + y + g.y_glob0) & 1) ? c1[src] : c0[src];
        }
        v |= code << (2 * i);
    }
    packed[gid] = (uint8_t)v;
}
__global__ void unpack2_any_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ c0, uint8_t* __restrict__ c1, const Geom g,
                                   int rep_first, int n_rep) {
    const long long row_bytes = (g.L + 3) >> 2, per_rep = (long long)g.Ly * g.L;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_rep * n_rep) return;
    const int r = (int)(gid / per_rep);
    const long long rem = gid - (long long)r * per_rep;
    const int y = (int)(rem / g.L), x = (int)(rem - (long long)y * g.L);
    const uint32_t code = (packed[((size_t)r * g.Ly + y) * row_bytes + (x >> 2)] >> (2 * (x & 3))) & 3u;
    const size_t dst = ((size_t)(rep_first + r) * g.rows_arr + g.row_first + y) * g.Wc + (x >> 1);
    if ((x + y + g.y_glob0) & 1) c1[dst] = (uint8_t)code; else c0[dst] = (uint8_t)code;
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

// After the pushes of an exchange have completed (stream order): advance this rank's epoch counter and publish the
// new epoch in the neighbours' flags.  (The HALO sweep kernel does the same from its last boundary CTA.)
// flags[0] of a rank = "top ghost rows valid through epoch" (written by the rank above),
// flags[1]           = "bottom ghost rows valid through epoch" (written by the rank below).
// The epoch lives in device memory so that graph-replayed launches need no per-launch argument; every rank issues
// the same sequence of exchanges, so "the epoch I am at" is also what both neighbours must have published.
__global__ void halo_signal_kernel(long long* clock, long long* peer_prev_flags, long long* peer_next_flags) {
    __threadfence_system();
    const long long epoch = *clock + 1;
    *clock = epoch;
    if (peer_prev_flags) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(peer_prev_flags + 1), "l"(epoch) : "memory");
    if (peer_next_flags) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(peer_next_flags + 0), "l"(epoch) : "memory");
}

// Before ghost rows are read: wait until both neighbours have published this rank's current epoch.  One thread, one
// GPU per rank (never two ranks on one GPU), bounded: gives up after timeout_ns and raises *err (sticky; every
// synchronising entry point of the library reports it) instead of hanging the GPU.
__global__ void halo_wait_kernel(const long long* my_flags, const long long* clock, int need_top, int need_bottom,
                                 long long timeout_ns, int* err) {
    const long long epoch = *clock;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int k = 0; k < 2; ++k) {
        if (!(k == 0 ? need_top : need_bottom)) continue;
        long long v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(my_flags + k) : "memory");
            if (v >= epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if ((long long)(t1 - t0) > timeout_ns) { *err = 1; break; }
        }
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

}  // namespace bc
