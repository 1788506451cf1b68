// bc_engine.cu -- host side of libbc_engine.so: the C ABI of include/bc_engine.h on top of
// the kernels in bc_kernels.cuh.  No CPU fallback: every entry point needs a CUDA device.
//
// Replaces the in-process hot path of the reference, `refill_random(); step(lattice);`
// (/root/reference/main.cpp:343-347, 363-367), for the Metropolis part of step()
// (main.cpp:159-161, metropolis() at main.cpp:89-112).
#include "../../include/bc_engine.h"

#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound lazily with dlopen, so libbc_engine.so has no NCCL dependency

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bc_kernels.cuh"

using namespace bc;

namespace {
thread_local std::string g_create_error;

// NCCL entry points used by the slab halo exchange, resolved at first use from the libnccl.so.2
// already in the process (torch's bundled copy under torchrun) or on the library path.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define BC_NCCL_SYM(field, name)                                         \
        field = reinterpret_cast<decltype(field)>(dlsym(handle, name));   \
        if (!field) { error = std::string("dlsym ") + name; handle = nullptr; return false; }
        BC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        BC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        BC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        BC_NCCL_SYM(Send, "ncclSend")
        BC_NCCL_SYM(Recv, "ncclRecv")
        BC_NCCL_SYM(AllReduce, "ncclAllReduce")
        BC_NCCL_SYM(GroupStart, "ncclGroupStart")
        BC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        BC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef BC_NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;
}  // namespace

struct bc_engine {
    int L = 0, boundary = 0, n_rep = 0, device = 0;
    uint64_t seed = 0;
    int Wc = 0;
    // geometry of the local piece (== the whole lattice unless this handle is one slab of a decomposition)
    int Ly = 0;         // owned rows
    int y_glob0 = 0;    // global row of the first owned row
    int rows_arr = 0;   // rows of one replica's colour array (Ly, or Ly + 2 ghost rows)
    int row_first = 0;  // array row of the first owned row
    bool ghost = false;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    // P2P halo exchange (peer pointers opened from the neighbours' CUDA IPC handles)
    bool p2p = false;
    long long* d_flags = nullptr;      // [0] top ghost epoch, [1] bottom ghost epoch (written by the neighbours)
    int* d_halo_err = nullptr;
    uint8_t* peer_prev_c[2] = {nullptr, nullptr};
    uint8_t* peer_next_c[2] = {nullptr, nullptr};
    long long* peer_prev_flags = nullptr;
    long long* peer_next_flags = nullptr;
    void* ipc_opened[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    long long* d_halo_clock = nullptr; // epochs published so far (device resident: graph-replayed kernels advance it)
    unsigned* d_halo_done = nullptr;   // boundary CTAs of the running HALO launch that have finished
    int opt_halo_timeout_ms = 60000;   // bounded wait for a neighbour's halo rows
    cudaStream_t halo_stream = nullptr;  // high-priority side stream for the slab halo exchange
    cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr;
    int opt_overlap = 1;
    int opt_fused_halo = 1;  // peer-memory slabs: halo stores issued by the boundary-strip kernel itself
    size_t colour_bytes = 0;  // bytes of one colour array (all replicas)
    uint8_t* d_c[2] = {nullptr, nullptr};
    uint16_t* d_tab16 = nullptr;
    uint32_t* d_tab31 = nullptr;
    uint32_t* d_tab2 = nullptr;   // [replica][54*54] pair tables of the streaming kernel (L % 32 == 0, L >= 128) and of the resident kernel
    std::vector<uint32_t> h_tab31;
    std::vector<uint8_t> params_set;
    unsigned long long* d_obs = nullptr;
    size_t obs_slots_cap = 0;
    unsigned long long* d_ties = nullptr;
    int8_t* d_stage = nullptr;
    size_t stage_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, tm0 = nullptr, tm1 = nullptr;
    bool have_timing = false;
    int64_t t = 0;
    int64_t launches = 0;
    // pending (not yet fetched) measurements of the last bc_sweep
    int64_t pending_meas = 0;
    std::vector<int64_t> pending_sweeps;
    // CUDA-graph replay of blocks of unmeasured sweeps (cuts per-launch CPU cost and launch gaps)
    int64_t* d_clock = nullptr;      // device-resident sweep counter read by graph-captured kernels
    cudaGraphExec_t graph_exec = nullptr;
    int graph_rows = -1, graph_block = 0, graph_halo = 0;  // plan signature the graph was captured for
    // bc_sweep_group with this handle as member 0: graph of a block of unmeasured group sweeps + what it was built for
    cudaGraphExec_t group_graph = nullptr;
    std::vector<uint64_t> group_sig;
    cudaEvent_t ev_group = nullptr;  // joins / forks the members' streams around a group sweep
    bool group_carveout_set = false; // shared-memory carve-out of the group kernels requested on this handle's device
    // persistent kernel (all half-sweeps of an unmeasured run in one cooperative launch)
    int opt_persist = 0;             // 0 never (default: measured slower than graph replay + PDL), 1 whenever possible
    unsigned* d_bar = nullptr;       // [0] arrivals, [1] generation, [2] abort
    int persist_cap[2] = {0, 0};     // resident CTAs of sweep_persist_kernel<OPEN> on this device
    unsigned graph_grid = 0;
    // cluster labelling (allocated on first use): what a labelled tile leaves between kernels (bc_cluster_tile.cuh)
    uint16_t* cl_par = nullptr;             // [tile][4096] labels
    uint64_t* cl_hb = nullptr;              // [tile][64] open bonds to the right
    uint64_t* cl_pm = nullptr;              // [tile][2][64] plus / minus masks of the rows
    uint64_t* cl_vbb = nullptr;             // [tile] open bonds into the tile below
    uint16_t* cl_proot = nullptr;           // [tile][256] roots of the perimeter sites
    uint32_t* cl_gparent = nullptr;         // [replica][L*L] global union-find over flagged tile roots (sparse)
    uint32_t* cl_labels = nullptr;          // [replica][L*L] smallest site of every site's cluster (measurements)
    // cluster move on a slab of a multi-rank decomposition: the other ranks' pieces of the global union-find and the
    // perimeter roots of the rank below, mapped through CUDA IPC (bc_slab_cluster_export / _connect)
    uint32_t* cl_peer_gparent[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const uint16_t* cl_peer_proot_below = nullptr;
    void* cl_ipc_opened[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool cl_connected = false;
    int* d_cl_barrier = nullptr;            // one int all-reduced over NCCL: a stream-ordered barrier between the phases
    uint32_t* d_bond_thr = nullptr;         // [replica] bond threshold p * 2^32, p = 1 - exp(-2 J / T)
    // cluster member lists (bc_cluster_sort.cuh): sort buffers, the sorted list, per-replica counts, dump-writer arrays
    uint2* cl_sort[2] = {nullptr, nullptr};
    uint2* cl_list = nullptr;               // == cl_sort[k] holding the final pass
    uint32_t* cl_sort_hist = nullptr;
    uint32_t* cl_count = nullptr;
    uint32_t* cl_sites = nullptr;
    uint32_t* cl_starts = nullptr;
    uint32_t* cl_ordinal = nullptr;
    int8_t* cl_signs = nullptr;
    uint32_t* cl_nclusters = nullptr;
    unsigned long long* d_gap_hist = nullptr;
    size_t gap_hist_cap = 0;
    int64_t state_version = 0;              // bumped by everything that changes a lattice; keys the list cache
    int64_t list_version = -1, list_mstep = -1, heads_version = -1, heads_mstep = -1;
    int list_kind = -1, heads_kind = -1;
    uint32_t* d_csize = nullptr;            // cluster sizes per root (bc_cluster_moments)
    unsigned long long* d_cmom = nullptr;   // [replica][3] moments
    int64_t cluster_step = 0;  // counter of the cluster-move RNG streams
    std::vector<uint32_t> h_bond_thr;
    // options
    int opt_use_graph = 1;
    int opt_graph_block = 16;  // sweeps per graph launch
    int opt_wave = 0;  // wavefront kernel for L2-resident lattices: 0 off (default: measured slower), 1 on
    int* d_wave = nullptr;        // [0] task counter, [1] abort flag, [2..] per-block progress
    size_t wave_cap = 0;
    int wave_grid[2] = {0, 0};
    int opt_pdl = 1;  // programmatic dependent launch between consecutive half-sweeps
    int opt_resident = -1;  // small-lattice shared-memory-resident kernel: -1 auto, 0 off, 1 force (if it fits)
    bool resident_attr_set[16] = {};
    int opt_rows = 0;  // rows per thread, 0 = auto
    int sm_count = 148;
    int opt_pair_min_rows = 8;  // pair-table kernels when a thread marches over at least this many rows (0 = never)
    int opt_force_generic = 0;
    std::string err;
};

#define BC_NCCL(e, call)                                                                         \
    do {                                                                                         \
        ncclResult_t _st = (call);                                                               \
        if (_st != ncclSuccess) {                                                                \
            (e)->err = std::string(#call) + ": " + g_nccl.GetErrorString(_st);                   \
            return BC_ERR_NCCL;                                                                  \
        }                                                                                        \
    } while (0)

typedef void (*fast_fn)(const SweepParams);
static void pair_kernels(fast_fn (&out)[16]);

static Geom geom_of(const bc_engine* e) {
    Geom g;
    g.L = e->L; g.Wc = e->Wc; g.Ly = e->Ly; g.rows_arr = e->rows_arr; g.row_first = e->row_first;
    g.y_glob0 = e->y_glob0; g.open = e->boundary == BC_OPEN; g.ghost = e->ghost;
    return g;
}

#define BC_CUDA(e, call)                                                                         \
    do {                                                                                         \
        cudaError_t _st = (call);                                                                \
        if (_st != cudaSuccess) {                                                                \
            (e)->err = std::string(#call) + ": " + cudaGetErrorString(_st);                      \
            return BC_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

static int fail(bc_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg; else g_create_error = msg;
    return code;
}

// ---------------------------------------------------------------------------------------------
// acceptance table: thr31[18c + 9f + S] for the reference rule main.cpp:94,105-106,109
// ---------------------------------------------------------------------------------------------
static void build_table(double T, double D, double J, uint32_t thr31[TABLE_ENTRIES], uint16_t tab16[TABLE_ENTRIES]) {
    const double B = 1 / T;  // main.cpp:321
    for (int c = 0; c < 3; ++c)
        for (int f = 0; f < 2; ++f)
            for (int S = 0; S < 9; ++S) {
                const int cur = c - 1, nb = S - 4;
                const int prop = (cur + 2 + f) % 3 - 1;                                          // main.cpp:94
                const double de = -J * (prop - cur) * nb + D * (prop * prop - cur * cur);        // main.cpp:105-106
                uint32_t thr;
                if (de <= 0) {                                                                   // main.cpp:109
                    thr = 0x80000000u;
                } else {
                    long long q = std::llround(std::exp(-B * de) * 2147483648.0);
                    if (q < 0) q = 0;
                    if (q > 2147483648LL) q = 2147483648LL;
                    thr = (uint32_t)q;
                }
                const int i = 18 * c + 9 * f + S;
                thr31[i] = thr;
                // coarse threshold compared against u15 (bits 0..14 of the draw): accept iff u15 < E, tie iff
                // u15 == E.  E = thr31 >> 16 in [0, 32768]; 32768 (always accept, thr31 = 2^31) never ties, exactly
                // like the oracle's lazy rule, and still fits the 16-bit lanes: 0x8000 + u15 - 32768 = u15 >= 0.
                const uint32_t E = thr >> 16;
                tab16[i] = (uint16_t)E;
            }
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
extern "C" int bc_abi_version(void) { return BC_ABI_VERSION; }

extern "C" const char* bc_last_error(const bc_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

extern "C" void bc_destroy(bc_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->halo_stream) cudaStreamSynchronize(e->halo_stream);
    for (void* p : e->ipc_opened) if (p) cudaIpcCloseMemHandle(p);
    for (void* p : e->cl_ipc_opened) if (p) cudaIpcCloseMemHandle(p);
    cudaFree(e->d_cl_barrier);
    cudaFree(e->d_flags); cudaFree(e->d_halo_err); cudaFree(e->d_halo_clock); cudaFree(e->d_halo_done);
    if (e->comm && g_nccl.handle) g_nccl.CommDestroy(e->comm);
    if (e->ev_boundary) cudaEventDestroy(e->ev_boundary);
    if (e->ev_halo) cudaEventDestroy(e->ev_halo);
    if (e->halo_stream) cudaStreamDestroy(e->halo_stream);
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->group_graph) cudaGraphExecDestroy(e->group_graph);
    cudaFree(e->d_bar);
    if (e->ev_group) cudaEventDestroy(e->ev_group);
    cudaFree(e->d_clock);
    cudaFree(e->cl_par); cudaFree(e->cl_hb); cudaFree(e->cl_pm); cudaFree(e->cl_vbb); cudaFree(e->cl_proot); cudaFree(e->cl_gparent);
    cudaFree(e->cl_labels); cudaFree(e->d_bond_thr); cudaFree(e->d_csize); cudaFree(e->d_cmom);
    cudaFree(e->cl_sort[0]); cudaFree(e->cl_sort[1]); cudaFree(e->cl_sort_hist); cudaFree(e->cl_count);
    cudaFree(e->cl_sites); cudaFree(e->cl_starts); cudaFree(e->cl_ordinal); cudaFree(e->cl_signs); cudaFree(e->cl_nclusters);
    cudaFree(e->d_gap_hist);
    cudaFree(e->d_wave);
    cudaFree(e->d_c[0]); cudaFree(e->d_c[1]);
    cudaFree(e->d_tab16); cudaFree(e->d_tab31); cudaFree(e->d_tab2);
    cudaFree(e->d_obs); cudaFree(e->d_ties); cudaFree(e->d_stage);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->tm0) cudaEventDestroy(e->tm0);
    if (e->tm1) cudaEventDestroy(e->tm1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int launch_fill(bc_engine* e, uint8_t* p, size_t n, uint8_t v) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    fill_kernel<<<blocks, 256, 0, e->stream>>>(p, n, v);
    ++e->launches;
    BC_CUDA(e, cudaGetLastError());
    return BC_OK;
}

static int create_common(bc_engine** out, int L, int boundary, int n_replicas, int storage, int device,
                         uint64_t seed, int nranks, int rank, const void* nccl_id) {
    if (!out) return fail(nullptr, BC_ERR_ARG, "bc_create: out is NULL");
    *out = nullptr;
    if (L < 2 || L > 65536) return fail(nullptr, BC_ERR_ARG, "bc_create: L out of range [2, 65536]");
    if (boundary != BC_PERIODIC && boundary != BC_OPEN) return fail(nullptr, BC_ERR_ARG, "bc_create: bad boundary");
    if (boundary == BC_PERIODIC && (L & 1))
        return fail(nullptr, BC_ERR_ARG, "bc_create: the periodic checkerboard needs even L");
    if (n_replicas < 1) return fail(nullptr, BC_ERR_ARG, "bc_create: n_replicas < 1");
    if (storage != BC_STORAGE_INT8) return fail(nullptr, BC_ERR_UNSUPPORTED, "bc_create: only BC_STORAGE_INT8");
    const bool slab = nranks > 0;
    if (slab) {
        if (rank < 0 || rank >= nranks) return fail(nullptr, BC_ERR_ARG, "bc_create_slab: rank out of range");
        if (L % 32 || L % nranks || ((L / nranks) & 1))
            return fail(nullptr, BC_ERR_ARG, "bc_create_slab: need L % 32 == 0 and an even number of rows per rank");
        if (nranks > 1 && !nccl_id) return fail(nullptr, BC_ERR_ARG, "bc_create_slab: nccl_id is NULL");
    }
    int ndev = 0;
    cudaError_t st = cudaGetDeviceCount(&ndev);
    if (st != cudaSuccess || ndev == 0)
        return fail(nullptr, BC_ERR_CUDA,
                    std::string("bc_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(st));
    if (device < 0 || device >= ndev) return fail(nullptr, BC_ERR_ARG, "bc_create: device index out of range");

    bc_engine* e = new bc_engine();
    e->L = L; e->boundary = boundary; e->n_rep = n_replicas; e->device = device; e->seed = seed;
    e->Wc = (((L + 1) / 2) + 15) / 16 * 16;
    if (slab) {
        e->nranks = nranks; e->rank = rank; e->ghost = true;
        e->Ly = L / nranks; e->y_glob0 = rank * e->Ly; e->rows_arr = e->Ly + 2; e->row_first = 1;
    } else {
        e->Ly = L; e->y_glob0 = 0; e->rows_arr = L; e->row_first = 0;
    }
    e->colour_bytes = (size_t)n_replicas * (size_t)e->rows_arr * (size_t)e->Wc;
    e->h_tab31.assign((size_t)n_replicas * TABLE_ENTRIES, 0u);
    e->params_set.assign((size_t)n_replicas, 0);
    e->h_bond_thr.assign((size_t)n_replicas, 0u);
    auto bail = [&](const char* what, cudaError_t s) {
        g_create_error = std::string("bc_create: ") + what + ": " + cudaGetErrorString(s);
        bc_destroy(e);
        return (int)BC_ERR_CUDA;
    };
    if ((st = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", st);
    if (cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || e->sm_count < 1) e->sm_count = 148;
    if ((st = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", st);
    if ((st = cudaEventCreate(&e->ev0)) != cudaSuccess) return bail("event", st);
    if ((st = cudaEventCreate(&e->ev1)) != cudaSuccess) return bail("event", st);
    if ((st = cudaEventCreate(&e->tm0)) != cudaSuccess) return bail("event", st);
    if ((st = cudaEventCreate(&e->tm1)) != cudaSuccess) return bail("event", st);
    if ((st = cudaMalloc(&e->d_c[0], e->colour_bytes)) != cudaSuccess) return bail("cudaMalloc lattice", st);
    if ((st = cudaMalloc(&e->d_c[1], e->colour_bytes)) != cudaSuccess) return bail("cudaMalloc lattice", st);
    if ((st = cudaMalloc(&e->d_tab16, (size_t)n_replicas * TABLE_ENTRIES * sizeof(uint16_t))) != cudaSuccess) return bail("cudaMalloc table", st);
    if ((st = cudaMalloc(&e->d_tab31, (size_t)n_replicas * TABLE_ENTRIES * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc table", st);
    if ((st = cudaMalloc(&e->d_bond_thr, (size_t)n_replicas * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc table", st);
    // pair tables: the streaming kernel's sizes, and the resident kernel's when a CTA holds one replica (more than 32
    // 16-byte chunks per colour: L >= 48)
    if (((L % 32 == 0 && L >= 128) || (L <= 448 && (long long)L * ((((L + 1) / 2) + 15) / 16) > 32)) &&
        (st = cudaMalloc(&e->d_tab2, (size_t)n_replicas * TAB2_WORDS * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc pair table", st);
    if (e->d_tab2) {  // 7 CTAs x 30 KB per SM need the large shared-memory carve-out
        fast_fn pf[16];
        pair_kernels(pf);
        for (fast_fn f : pf) cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if ((st = cudaMalloc(&e->d_ties, sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc ties", st);
    if ((st = cudaMalloc(&e->d_clock, sizeof(int64_t))) != cudaSuccess) return bail("cudaMalloc clock", st);
    if ((st = cudaMemsetAsync(e->d_ties, 0, sizeof(unsigned long long), e->stream)) != cudaSuccess) return bail("memset", st);
    // all spins 0 (code 1), including the pad bytes of each compact row and the ghost rows
    if (launch_fill(e, e->d_c[0], e->colour_bytes, 1) != BC_OK || launch_fill(e, e->d_c[1], e->colour_bytes, 1) != BC_OK) {
        g_create_error = e->err;
        bc_destroy(e);
        return BC_ERR_CUDA;
    }
    if ((st = cudaStreamSynchronize(e->stream)) != cudaSuccess) return bail("sync", st);
    if (slab) {
        if ((st = cudaMalloc(&e->d_flags, 2 * sizeof(long long))) != cudaSuccess) return bail("cudaMalloc flags", st);
        if ((st = cudaMemset(e->d_flags, 0, 2 * sizeof(long long))) != cudaSuccess) return bail("memset", st);
        if ((st = cudaMalloc(&e->d_halo_err, sizeof(int))) != cudaSuccess) return bail("cudaMalloc", st);
        if ((st = cudaMemset(e->d_halo_err, 0, sizeof(int))) != cudaSuccess) return bail("memset", st);
        if ((st = cudaMalloc(&e->d_halo_clock, sizeof(long long))) != cudaSuccess) return bail("cudaMalloc", st);
        if ((st = cudaMemset(e->d_halo_clock, 0, sizeof(long long))) != cudaSuccess) return bail("memset", st);
        if ((st = cudaMalloc(&e->d_halo_done, sizeof(unsigned))) != cudaSuccess) return bail("cudaMalloc", st);
        if ((st = cudaMemset(e->d_halo_done, 0, sizeof(unsigned))) != cudaSuccess) return bail("memset", st);
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if ((st = cudaStreamCreateWithPriority(&e->halo_stream, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail("halo stream", st);
        if ((st = cudaEventCreateWithFlags(&e->ev_boundary, cudaEventDisableTiming)) != cudaSuccess) return bail("event", st);
        if ((st = cudaEventCreateWithFlags(&e->ev_halo, cudaEventDisableTiming)) != cudaSuccess) return bail("event", st);
    }
    if (slab && nranks > 1) {
        if (!g_nccl.load()) {
            g_create_error = "bc_create_slab: " + g_nccl.error;
            bc_destroy(e);
            return BC_ERR_NCCL;
        }
        ncclUniqueId id;
        std::memcpy(&id, nccl_id, sizeof(id));
        ncclResult_t ns = g_nccl.CommInitRank(&e->comm, nranks, id, rank);
        if (ns != ncclSuccess) {
            g_create_error = std::string("bc_create_slab: ncclCommInitRank: ") + g_nccl.GetErrorString(ns);
            e->comm = nullptr;
            bc_destroy(e);
            return BC_ERR_NCCL;
        }
    }
    *out = e;
    return BC_OK;
}

extern "C" int bc_create(bc_engine** out, int L, int boundary, int n_replicas, int storage, int device,
                         uint64_t seed) {
    return create_common(out, L, boundary, n_replicas, storage, device, seed, 0, 0, nullptr);
}

extern "C" int bc_slab_unique_id(void* id128) {
    if (!id128) return BC_ERR_ARG;
    if (!g_nccl.load()) return fail(nullptr, BC_ERR_NCCL, "bc_slab_unique_id: " + g_nccl.error);
    ncclUniqueId id;
    ncclResult_t ns = g_nccl.GetUniqueId(&id);
    if (ns != ncclSuccess) return fail(nullptr, BC_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(ns));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, sizeof(id));
    return BC_OK;
}

extern "C" int bc_create_slab(bc_engine** out, int L, int boundary, int n_replicas, int storage, int device,
                              uint64_t seed, int n_ranks, int rank, const void* nccl_id128) {
    if (n_ranks < 1) return fail(nullptr, BC_ERR_ARG, "bc_create_slab: n_ranks < 1");
    return create_common(out, L, boundary, n_replicas, storage, device, seed, n_ranks, rank, nccl_id128);
}

extern "C" int bc_local_rows(const bc_engine* e, int* rows, int* first_global_row) {
    if (!e) return BC_ERR_ARG;
    if (rows) *rows = e->Ly;
    if (first_global_row) *first_global_row = e->y_glob0;
    return BC_OK;
}

// ---------------------------------------------------------------------------------------------
// parameters
// ---------------------------------------------------------------------------------------------
extern "C" int bc_set_params(bc_engine* e, int replica, double T, double D, double J) {
    if (!e) return BC_ERR_ARG;
    if (replica < -1 || replica >= e->n_rep) return fail(e, BC_ERR_ARG, "bc_set_params: replica out of range");
    if (!(T > 0) || !std::isfinite(T) || !std::isfinite(D) || !std::isfinite(J))
        return fail(e, BC_ERR_ARG, "bc_set_params: need finite D, J and T > 0");
    BC_CUDA(e, cudaSetDevice(e->device));
    uint32_t thr[TABLE_ENTRIES];
    uint16_t t16[TABLE_ENTRIES];
    build_table(T, D, J, thr, t16);
    const int r0 = replica < 0 ? 0 : replica, r1 = replica < 0 ? e->n_rep : replica + 1;
    // bond probability of the cluster move, p = 1 - exp(-2 J / T) (main.cpp:115), as a 32-bit threshold
    uint32_t bond_thr = 0;
    {
        const double p = 1 - std::exp(-2 * (1 / T) * J);
        if (p > 0) { const double q = std::floor(p * 4294967296.0 + 0.5); bond_thr = q > 4294967295.0 ? 4294967295u : (uint32_t)q; }
    }
    std::vector<uint16_t> h16((size_t)(r1 - r0) * TABLE_ENTRIES);
    for (int r = r0; r < r1; ++r) {
        std::memcpy(&e->h_tab31[(size_t)r * TABLE_ENTRIES], thr, sizeof(thr));
        std::memcpy(&h16[(size_t)(r - r0) * TABLE_ENTRIES], t16, sizeof(t16));
        e->params_set[r] = 1;
        e->h_bond_thr[r] = bond_thr;
    }
    // Everything on the engine's stream, in order: the two table copies, the pair-table kernel that reads the coarse
    // table, and (after the final synchronize) whatever sweep comes next.  h16 is a temporary, and a sweep launched
    // with programmatic stream serialization stages the pair table BEFORE griddepcontrol.wait: that is only safe if
    // pair_table_kernel has completed and flushed, hence the wait here (not a hot call).
    BC_CUDA(e, cudaMemcpyAsync(e->d_tab31 + (size_t)r0 * TABLE_ENTRIES, &e->h_tab31[(size_t)r0 * TABLE_ENTRIES],
                               (size_t)(r1 - r0) * TABLE_ENTRIES * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    BC_CUDA(e, cudaMemcpyAsync(e->d_tab16 + (size_t)r0 * TABLE_ENTRIES, h16.data(),
                               h16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream));
    BC_CUDA(e, cudaMemcpyAsync(e->d_bond_thr + r0, &e->h_bond_thr[(size_t)r0], (size_t)(r1 - r0) * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, e->stream));
    if (e->d_tab2) {
        const long long words = (long long)(r1 - r0) * TAB2_WORDS;
        pair_table_kernel<<<(unsigned)((words + 255) / 256), 256, 0, e->stream>>>(e->d_tab16, e->d_tab2, r0, r1 - r0);
        BC_CUDA(e, cudaGetLastError());
    }
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    return BC_OK;
}

extern "C" int bc_get_table(const bc_engine* e, int replica, uint32_t thr31[54]) {
    if (!e || !thr31) return BC_ERR_ARG;
    if (replica < 0 || replica >= e->n_rep) return BC_ERR_ARG;
    if (!e->params_set[replica]) return BC_ERR_STATE;
    std::memcpy(thr31, &e->h_tab31[(size_t)replica * TABLE_ENTRIES], TABLE_ENTRIES * sizeof(uint32_t));
    return BC_OK;
}

// One translation unit, split by subject (the parts are included exactly once, in this order, and share the
// static helpers above):
#include "engine_halo.inl"  // slab halo exchange over NCCL or peer memory (between half-sweeps of a slab handle)
#include "engine_state.inl"  // lattice state: initialisation, uploads / downloads in the reference layout, sweep counter
#include "engine_sweep.inl"  // the hot path: launch plans, half-sweep launches, CUDA-graph replay, wavefront / persistent / resident paths, bc_sweep
#include "engine_group.inl"  // bc_sweep_group (several lattice sizes in one launch per half-sweep), bc_read_obs, bc_measure
#include "engine_cluster.inl"  // cluster move and cluster-size moments, bc_sync
#include "engine_ipc.inl"  // peer-memory slab halos: export / connect CUDA IPC handles
#include "engine_introspection.inl"  // introspection, options, plan description
