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
    long long halo_epoch = 0;          // exchanges issued so far (identical on every rank)
    long long halo_waited = 0;         // last epoch a wait was enqueued for
    cudaStream_t halo_stream = nullptr;  // high-priority side stream for the slab halo exchange
    cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr;
    int opt_overlap = 1;
    int opt_fused_halo = 1;  // peer-memory slabs: halo stores issued by the boundary-strip kernel itself
    size_t colour_bytes = 0;  // bytes of one colour array (all replicas)
    uint8_t* d_c[2] = {nullptr, nullptr};
    uint16_t* d_tab16 = nullptr;
    uint32_t* d_tab31 = nullptr;
    uint32_t* d_tab2 = nullptr;   // [replica][54*54] pair tables of the streaming kernel (L % 32 == 0, L >= 128)
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
    int graph_rows = -1, graph_block = 0;  // plan signature the graph was captured for
    // bc_sweep_group with this handle as member 0: graph of a block of unmeasured group sweeps + what it was built for
    cudaGraphExec_t group_graph = nullptr;
    std::vector<uint64_t> group_sig;
    cudaEvent_t ev_group = nullptr;  // joins / forks the members' streams around a group sweep
    // persistent kernel (all half-sweeps of an unmeasured run in one cooperative launch)
    int opt_persist = 0;             // 0 never (default: measured slower than graph replay + PDL), 1 whenever possible
    unsigned* d_bar = nullptr;       // [0] arrivals, [1] generation, [2] abort
    int persist_cap[2] = {0, 0};     // resident CTAs of sweep_persist_kernel<OPEN> on this device
    unsigned graph_grid = 0;
    // cluster move scratch
    int32_t* d_parent = nullptr;
    uint8_t* d_bonds = nullptr;
    int32_t* d_csize = nullptr;             // cluster sizes per root (bc_cluster_moments), allocated on first use
    unsigned long long* d_cmom = nullptr;   // [replica][3] moments
    size_t cluster_cap = 0;
    int64_t cluster_step = 0;  // counter of the cluster-move RNG streams
    std::vector<double> h_TJ;  // (T, J) per replica, for the bond probability
    // options
    int opt_use_graph = 1;
    int opt_graph_block = 16;  // sweeps per graph launch
    int opt_wave = 0;  // wavefront kernel for L2-resident lattices: 0 off (default: measured slower), 1 on
    int* d_wave = nullptr;        // [0] task counter, [1] abort flag, [2..] per-block progress
    size_t wave_cap = 0;
    int wave_grid[2] = {0, 0};
    int opt_pdl = 1;  // programmatic dependent launch between consecutive half-sweeps
    int opt_resident = -1;  // small-lattice shared-memory-resident kernel: -1 auto, 0 off, 1 force (if it fits)
    bool resident_attr_set[4] = {false, false, false, false};
    int opt_rows = 0;  // rows per thread, 0 = auto
    int opt_pair_min_rows = 8;  // pair-table kernels when a thread marches over at least this many rows (0 = never)
    int opt_swp_min_rows = 32;  // ... with software-pipelined Philox draws from this many rows on (0 = never)
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
static void pair_kernels(fast_fn (&out)[10]);

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
    cudaFree(e->d_flags); cudaFree(e->d_halo_err);
    if (e->comm && g_nccl.handle) g_nccl.CommDestroy(e->comm);
    if (e->ev_boundary) cudaEventDestroy(e->ev_boundary);
    if (e->ev_halo) cudaEventDestroy(e->ev_halo);
    if (e->halo_stream) cudaStreamDestroy(e->halo_stream);
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->group_graph) cudaGraphExecDestroy(e->group_graph);
    cudaFree(e->d_bar);
    if (e->ev_group) cudaEventDestroy(e->ev_group);
    cudaFree(e->d_clock);
    cudaFree(e->d_parent); cudaFree(e->d_bonds); cudaFree(e->d_csize); cudaFree(e->d_cmom);
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
    e->h_TJ.assign((size_t)n_replicas * 2, 0.0);
    auto bail = [&](const char* what, cudaError_t s) {
        g_create_error = std::string("bc_create: ") + what + ": " + cudaGetErrorString(s);
        bc_destroy(e);
        return (int)BC_ERR_CUDA;
    };
    if ((st = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", st);
    if ((st = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", st);
    if ((st = cudaEventCreate(&e->ev0)) != cudaSuccess) return bail("event", st);
    if ((st = cudaEventCreate(&e->ev1)) != cudaSuccess) return bail("event", st);
    if ((st = cudaEventCreate(&e->tm0)) != cudaSuccess) return bail("event", st);
    if ((st = cudaEventCreate(&e->tm1)) != cudaSuccess) return bail("event", st);
    if ((st = cudaMalloc(&e->d_c[0], e->colour_bytes)) != cudaSuccess) return bail("cudaMalloc lattice", st);
    if ((st = cudaMalloc(&e->d_c[1], e->colour_bytes)) != cudaSuccess) return bail("cudaMalloc lattice", st);
    if ((st = cudaMalloc(&e->d_tab16, (size_t)n_replicas * TABLE_ENTRIES * sizeof(uint16_t))) != cudaSuccess) return bail("cudaMalloc table", st);
    if ((st = cudaMalloc(&e->d_tab31, (size_t)n_replicas * TABLE_ENTRIES * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc table", st);
    if (L % 32 == 0 && L >= 128 &&
        (st = cudaMalloc(&e->d_tab2, (size_t)n_replicas * TAB2_WORDS * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc pair table", st);
    if (e->d_tab2) {  // 7 CTAs x 26 KB per SM need the large shared-memory carve-out
        fast_fn pf[10];
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
    std::vector<uint16_t> h16((size_t)(r1 - r0) * TABLE_ENTRIES);
    for (int r = r0; r < r1; ++r) {
        std::memcpy(&e->h_tab31[(size_t)r * TABLE_ENTRIES], thr, sizeof(thr));
        std::memcpy(&h16[(size_t)(r - r0) * TABLE_ENTRIES], t16, sizeof(t16));
        e->params_set[r] = 1;
        e->h_TJ[(size_t)r * 2] = T; e->h_TJ[(size_t)r * 2 + 1] = J;
    }
    // synchronous copies: the tables are tiny and the host vectors are temporaries
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    BC_CUDA(e, cudaMemcpy(e->d_tab31 + (size_t)r0 * TABLE_ENTRIES, &e->h_tab31[(size_t)r0 * TABLE_ENTRIES],
                          (size_t)(r1 - r0) * TABLE_ENTRIES * sizeof(uint32_t), cudaMemcpyHostToDevice));
    BC_CUDA(e, cudaMemcpy(e->d_tab16 + (size_t)r0 * TABLE_ENTRIES, h16.data(),
                          h16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    if (e->d_tab2) {
        const long long words = (long long)(r1 - r0) * TAB2_WORDS;
        pair_table_kernel<<<(unsigned)((words + 255) / 256), 256, 0, e->stream>>>(e->d_tab16, e->d_tab2, r0, r1 - r0);
        BC_CUDA(e, cudaGetLastError());
    }
    return BC_OK;
}

extern "C" int bc_get_table(const bc_engine* e, int replica, uint32_t thr31[54]) {
    if (!e || !thr31) return BC_ERR_ARG;
    if (replica < 0 || replica >= e->n_rep) return BC_ERR_ARG;
    if (!e->params_set[replica]) return BC_ERR_STATE;
    std::memcpy(thr31, &e->h_tab31[(size_t)replica * TABLE_ENTRIES], TABLE_ENTRIES * sizeof(uint32_t));
    return BC_OK;
}

// ---------------------------------------------------------------------------------------------
// slab halo exchange: after colour `col` was updated, its first owned row goes to the bottom ghost row
// of the rank above and its last owned row to the top ghost row of the rank below (periodic wrap; an
// open lattice keeps code-1 ghost rows at the global edges).  32 KiB per message at L = 65536.
// ---------------------------------------------------------------------------------------------
// fused = true: the boundary-strip kernel already stored the halo rows into the neighbours (peer-memory mode)
static int exchange_halo(bc_engine* e, int col, cudaStream_t stream = nullptr, bool fused = false) {
    if (!e->ghost) return BC_OK;
    if (!stream) stream = e->stream;
    const bool periodic = e->boundary == BC_PERIODIC;
    const size_t Wc = (size_t)e->Wc, rep_stride = (size_t)e->rows_arr * Wc;
    if (e->nranks == 1) {
        if (!periodic) return BC_OK;
        const long long total = 2LL * e->Wc * e->n_rep;
        halo_self_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(e->d_c[col], geom_of(e), e->n_rep);
        ++e->launches;
        BC_CUDA(e, cudaGetLastError());
        return BC_OK;
    }
    const int up = e->rank - 1, dn = e->rank + 1;
    const int prev = up >= 0 ? up : (periodic ? e->nranks - 1 : -1);
    const int next = dn < e->nranks ? dn : (periodic ? 0 : -1);
    if (e->p2p) {
        // direct stores into the neighbours' ghost rows over NVLink, then publish the epoch in their flags
        if (!fused) {
            const long long total = 2LL * (e->Wc / 16) * e->n_rep;
            halo_push_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
                e->d_c[col], prev >= 0 ? e->peer_prev_c[col] : nullptr, next >= 0 ? e->peer_next_c[col] : nullptr,
                geom_of(e), e->n_rep);
            ++e->launches;
        }
        ++e->halo_epoch;
        halo_signal_kernel<<<1, 1, 0, stream>>>(prev >= 0 ? e->peer_prev_flags : nullptr,
                                                next >= 0 ? e->peer_next_flags : nullptr, e->halo_epoch);
        ++e->launches;
        BC_CUDA(e, cudaGetLastError());
        return BC_OK;
    }
    BC_NCCL(e, g_nccl.GroupStart());
    for (int r = 0; r < e->n_rep; ++r) {
        uint8_t* base = e->d_c[col] + (size_t)r * rep_stride;
        uint8_t* first_owned = base + Wc;
        uint8_t* last_owned = base + (size_t)e->Ly * Wc;
        uint8_t* ghost_top = base;
        uint8_t* ghost_bot = base + (size_t)(e->Ly + 1) * Wc;
        // order matters when prev == next (2 ranks): sends (first, last) pair with recvs (bottom, top)
        if (prev >= 0) BC_NCCL(e, g_nccl.Send(first_owned, Wc, ncclUint8, prev, e->comm, stream));
        if (next >= 0) BC_NCCL(e, g_nccl.Send(last_owned, Wc, ncclUint8, next, e->comm, stream));
        if (next >= 0) BC_NCCL(e, g_nccl.Recv(ghost_bot, Wc, ncclUint8, next, e->comm, stream));
        if (prev >= 0) BC_NCCL(e, g_nccl.Recv(ghost_top, Wc, ncclUint8, prev, e->comm, stream));
    }
    BC_NCCL(e, g_nccl.GroupEnd());
    return BC_OK;
}

// P2P mode: before a kernel reads ghost rows, wait (on the GPU) for the neighbours' latest epoch
static int halo_wait(bc_engine* e, cudaStream_t stream) {
    if (!e->p2p || e->halo_waited == e->halo_epoch) return BC_OK;
    const bool periodic = e->boundary == BC_PERIODIC;
    const int need_top = (e->rank > 0 || periodic) ? 1 : 0, need_bottom = (e->rank + 1 < e->nranks || periodic) ? 1 : 0;
    halo_wait_kernel<<<1, 1, 0, stream>>>(e->d_flags, need_top, need_bottom, e->halo_epoch, e->d_halo_err);
    ++e->launches;
    e->halo_waited = e->halo_epoch;
    BC_CUDA(e, cudaGetLastError());
    return BC_OK;
}

static int exchange_both(bc_engine* e) {
    // P2P: the neighbours must have finished reading their ghost rows (their last boundary kernel precedes
    // their last signal) before this rank overwrites them outside the half-sweep cadence (init / upload)
    if (e->p2p) { int rc0 = halo_wait(e, e->stream); if (rc0 != BC_OK) return rc0; }
    int rc = exchange_halo(e, 0);
    if (rc == BC_OK) rc = exchange_halo(e, 1);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// lattice state
// ---------------------------------------------------------------------------------------------
static int ensure_stage(bc_engine* e, size_t bytes) {
    if (e->stage_cap >= bytes) return BC_OK;
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    if (e->d_stage) BC_CUDA(e, cudaFree(e->d_stage));
    e->d_stage = nullptr; e->stage_cap = 0;
    BC_CUDA(e, cudaMalloc(&e->d_stage, bytes));
    e->stage_cap = bytes;
    return BC_OK;
}

extern "C" int bc_init_random(bc_engine* e, int replica) {
    if (!e) return BC_ERR_ARG;
    if (replica < -1 || replica >= e->n_rep) return fail(e, BC_ERR_ARG, "bc_init_random: replica out of range");
    BC_CUDA(e, cudaSetDevice(e->device));
    const int r0 = replica < 0 ? 0 : replica, n = replica < 0 ? e->n_rep : 1;
    const long long total = (long long)n * e->Ly * e->L;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    init_kernel<<<blocks, 256, 0, e->stream>>>(e->d_c[0], e->d_c[1], geom_of(e), r0, n, (uint32_t)e->seed,
                                               (uint32_t)(e->seed >> 32));
    ++e->launches;
    BC_CUDA(e, cudaGetLastError());
    return exchange_both(e);
}

// 16-byte conversion kernels: rows 16-byte aligned on both sides, one grid row per replica
static bool vector_layout(const bc_engine* e, int n) {
    return e->L % 32 == 0 && e->Wc == e->L / 2 && n <= 65535 && (long long)e->Ly * (e->L / 32) < (1LL << 31);
}

// reference layout (row-major int8 in d_stage, replicas r0 .. r0+n-1 at the start of the buffer) <-> compact colour arrays
static int stage_to_device(bc_engine* e, int r0, int n) {
    if (vector_layout(e, n)) {
        const dim3 grid((unsigned)(((long long)e->Ly * (e->L / 32) + 255) / 256), (unsigned)n);
        pack16_kernel<<<grid, 256, 0, e->stream>>>(e->d_stage, e->d_c[0], e->d_c[1], geom_of(e), r0);
    } else {
        const long long total = (long long)n * e->Ly * e->Wc;
        const unsigned blocks = (unsigned)((total + 255) / 256);
        pack_kernel<<<blocks, 256, 0, e->stream>>>(e->d_stage, e->d_c[0], e->d_c[1], geom_of(e), r0, n);
    }
    ++e->launches;
    BC_CUDA(e, cudaGetLastError());
    return BC_OK;
}

static int stage_from_device(bc_engine* e, int r0, int n) {
    if (vector_layout(e, n)) {
        const dim3 grid((unsigned)(((long long)e->Ly * (e->L / 32) + 255) / 256), (unsigned)n);
        unpack16_kernel<<<grid, 256, 0, e->stream>>>(e->d_stage, e->d_c[0], e->d_c[1], geom_of(e), r0);
    } else {
        const long long total = (long long)n * e->Ly * e->L;
        const unsigned blocks = (unsigned)((total + 255) / 256);
        unpack_kernel<<<blocks, 256, 0, e->stream>>>(e->d_stage, e->d_c[0], e->d_c[1], geom_of(e), r0, n);
    }
    ++e->launches;
    BC_CUDA(e, cudaGetLastError());
    return BC_OK;
}

static int upload_range(bc_engine* e, int r0, int n, const int8_t* spins, bool wait = true) {
    BC_CUDA(e, cudaSetDevice(e->device));
    const size_t bytes = (size_t)n * e->Ly * e->L;
    int rc = ensure_stage(e, bytes);
    if (rc != BC_OK) return rc;
    BC_CUDA(e, cudaMemcpyAsync(e->d_stage, spins, bytes, cudaMemcpyHostToDevice, e->stream));
    rc = stage_to_device(e, r0, n);
    if (rc != BC_OK) return rc;
    rc = exchange_both(e);
    if (rc != BC_OK) return rc;
    if (wait) BC_CUDA(e, cudaStreamSynchronize(e->stream));  // the caller may reuse `spins` on return
    return BC_OK;
}

static int download_range(bc_engine* e, int r0, int n, int8_t* spins, bool wait = true) {
    BC_CUDA(e, cudaSetDevice(e->device));
    const size_t bytes = (size_t)n * e->Ly * e->L;
    int rc = ensure_stage(e, bytes);
    if (rc != BC_OK) return rc;
    rc = stage_from_device(e, r0, n);
    if (rc != BC_OK) return rc;
    BC_CUDA(e, cudaMemcpyAsync(spins, e->d_stage, bytes, cudaMemcpyDeviceToHost, e->stream));
    if (wait) BC_CUDA(e, cudaStreamSynchronize(e->stream));
    return BC_OK;
}

extern "C" int bc_upload(bc_engine* e, int replica, const int8_t* spins) {
    if (!e || !spins) return BC_ERR_ARG;
    if (replica < 0 || replica >= e->n_rep) return fail(e, BC_ERR_ARG, "bc_upload: replica out of range");
    return upload_range(e, replica, 1, spins);
}
extern "C" int bc_download(bc_engine* e, int replica, int8_t* spins) {
    if (!e || !spins) return BC_ERR_ARG;
    if (replica < 0 || replica >= e->n_rep) return fail(e, BC_ERR_ARG, "bc_download: replica out of range");
    return download_range(e, replica, 1, spins);
}
extern "C" int bc_upload_all(bc_engine* e, const int8_t* spins) {
    if (!e || !spins) return BC_ERR_ARG;
    return upload_range(e, 0, e->n_rep, spins);
}
extern "C" int bc_download_all(bc_engine* e, int8_t* spins) {
    if (!e || !spins) return BC_ERR_ARG;
    return download_range(e, 0, e->n_rep, spins);
}

extern "C" int bc_upload_all_async(bc_engine* e, const int8_t* spins) {
    if (!e || !spins) return BC_ERR_ARG;
    return upload_range(e, 0, e->n_rep, spins, false);
}
extern "C" int bc_download_all_async(bc_engine* e, int8_t* spins) {
    if (!e || !spins) return BC_ERR_ARG;
    return download_range(e, 0, e->n_rep, spins, false);
}

extern "C" int bc_get_sweep_counter(const bc_engine* e, int64_t* t) {
    if (!e || !t) return BC_ERR_ARG;
    *t = e->t;
    return BC_OK;
}
extern "C" int bc_set_sweep_counter(bc_engine* e, int64_t t) {
    if (!e || t < 0) return BC_ERR_ARG;
    e->t = t;
    return BC_OK;
}

// ---------------------------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------------------------

template <int COL, bool SMALL, bool MEAS>
static fast_fn fast_open(bool open) {
    return open ? (fast_fn)sweep_fast_kernel<COL, SMALL, MEAS, true> : (fast_fn)sweep_fast_kernel<COL, SMALL, MEAS, false>;
}

template <int COL, bool MEAS>
static fast_fn fast_halo(bool open) {
    return open ? (fast_fn)sweep_fast_kernel<COL, false, MEAS, true, true> : (fast_fn)sweep_fast_kernel<COL, false, MEAS, false, true>;
}

template <int COL, bool MEAS>
static fast_fn fast_pair(bool open, bool swp = false) {
    if (swp && !open && !MEAS) return (fast_fn)sweep_fast_kernel<COL, false, false, false, false, 2>;
    return open ? (fast_fn)sweep_fast_kernel<COL, false, MEAS, true, false, 1>
                : (fast_fn)sweep_fast_kernel<COL, false, MEAS, false, false, 1>;
}

static void pair_kernels(fast_fn (&out)[10]) {
    int n = 0;
    for (int open = 0; open < 2; ++open) {
        out[n++] = fast_pair<0, false>(open); out[n++] = fast_pair<0, true>(open);
        out[n++] = fast_pair<1, false>(open); out[n++] = fast_pair<1, true>(open);
    }
    out[n++] = fast_pair<0, false>(false, true); out[n++] = fast_pair<1, false>(false, true);
}

static fast_fn pick_fast(int col, bool small, bool meas, bool open, bool halo = false, int pair = 0) {
    if (pair && !small && !halo) {
        if (col == 0) return meas ? fast_pair<0, true>(open) : fast_pair<0, false>(open, pair == 2);
        return meas ? fast_pair<1, true>(open) : fast_pair<1, false>(open, pair == 2);
    }
    if (halo && !small) {
        if (col == 0) return meas ? fast_halo<0, true>(open) : fast_halo<0, false>(open);
        return meas ? fast_halo<1, true>(open) : fast_halo<1, false>(open);
    }
    if (col == 0) {
        if (!small) return meas ? fast_open<0, false, true>(open) : fast_open<0, false, false>(open);
        return meas ? fast_open<0, true, true>(open) : fast_open<0, true, false>(open);
    }
    if (!small) return meas ? fast_open<1, false, true>(open) : fast_open<1, false, false>(open);
    return meas ? fast_open<1, true, true>(open) : fast_open<1, true, false>(open);
}

struct LaunchPlan {
    bool fast = false;
    bool small = false;
    int pair = 0;  // 1: pair-table kernels (one replica per CTA, enough rows per thread to amortise the 11.4 KB
                   // table); 2: with software-pipelined draws as well (long strips)
    int cpr = 0, strips = 0, rows = 0;
    unsigned grid = 0;
};

static bool rows_valid(int Ly, int cpr, int R) {
    if (R < 1 || Ly % R) return false;
    return (((long long)(Ly / R) * cpr) % 32) == 0;
}

static LaunchPlan make_plan(const bc_engine* e) {
    LaunchPlan p;
    const int L = e->L, Ly = e->Ly;
    if (L % 32 == 0 && (!e->opt_force_generic || e->ghost)) {
        p.fast = true;
        p.cpr = L / 32;
        int R = 0;
        if (e->opt_rows > 0 && rows_valid(Ly, p.cpr, e->opt_rows)) {
            R = e->opt_rows;
        } else {
            const long long want = 148LL * 768;  // enough threads to fill the chip (6-7 CTAs of 128 per SM)
            // 64 rows amortise the per-strip prologue (pair table, first draws) best, but only when the grid still
            // has four full waves of CTAs (1821 -> 1841 updates/ns on 64 x L=4096)
            if (rows_valid(Ly, p.cpr, 64) && (long long)e->n_rep * (Ly / 64) * p.cpr >= 4 * want) R = 64;
            const int cand[5] = {32, 16, 8, 4, 2};
            for (int i = 0; i < 5 && !R; ++i)
                if (rows_valid(Ly, p.cpr, cand[i]) && (long long)e->n_rep * (Ly / cand[i]) * p.cpr >= want) R = cand[i];
            if (!R) R = rows_valid(Ly, p.cpr, 2) ? 2 : 1;
        }
        p.rows = R;
        p.strips = Ly / R;
        p.small = (R & 1) != 0 || (((long long)p.strips * p.cpr) % FAST_THREADS) != 0;
        if (!p.small && e->d_tab2 && e->opt_pair_min_rows > 0 && R >= e->opt_pair_min_rows)
            p.pair = (e->opt_swp_min_rows > 0 && R >= e->opt_swp_min_rows) ? 2 : 1;
        const long long items = (long long)e->n_rep * p.strips * p.cpr;
        p.grid = (unsigned)((items + FAST_THREADS - 1) / FAST_THREADS);
    } else {
        const long long items = (long long)e->n_rep * L * ((L + 1) / 2);
        p.grid = (unsigned)((items + 255) / 256);
    }
    return p;
}

// kernel arguments of one half-sweep (colour `col`) of all strips of handle e
static SweepParams sweep_params(const bc_engine* e, const LaunchPlan& plan, int col, int64_t t, int64_t slot,
                                const int64_t* t_ptr) {
    SweepParams P;
    P.own = e->d_c[col];
    P.other = e->d_c[1 - col];
    P.tab16 = e->d_tab16;
    P.tab31 = e->d_tab31;
    P.tab2 = e->d_tab2;
    P.obs = e->d_obs;
    P.ties = e->d_ties;
    P.t_ptr = t_ptr;  // graph replay: t = *t_ptr + t
    P.slot_ptr = nullptr;
    P.t = t;
    P.slot = slot;
    P.L = e->L;
    P.rows_arr = e->rows_arr;
    P.row_first = e->row_first;
    P.rows_per_strip = plan.rows;
    P.y_rng_off = e->y_glob0 - e->row_first;
    P.Ly = e->Ly;
    P.peer_prev = nullptr;
    P.peer_next = nullptr;
    P.Wc = e->Wc;
    P.cpr = plan.cpr;
    P.strips = plan.strips;
    P.strip_base = 0;
    P.strip_stride = 1;
    P.n_replicas = e->n_rep;
    P.open = e->boundary == BC_OPEN;
    P.key0 = (uint32_t)e->seed;
    P.key1 = (uint32_t)(e->seed >> 32);
    return P;
}

// part: 0 = all strips, 1 = the two boundary strips of the slab, 2 = the interior strips
static int launch_half(bc_engine* e, const LaunchPlan& plan, int col, bool measure, int64_t t, int64_t slot,
                       const int64_t* t_ptr = nullptr, int part = 0, cudaStream_t stream = nullptr) {
    if (!stream) stream = e->stream;
    SweepParams P = sweep_params(e, plan, col, t, slot, t_ptr);
    // boundary-strip launch of a peer-memory slab: the kernel itself stores the halo rows into the neighbours
    const bool fused_halo = part == 1 && e->p2p && e->opt_fused_halo && plan.fast && !plan.small;
    if (fused_halo) {
        const bool periodic = e->boundary == BC_PERIODIC;
        if (e->rank > 0 || periodic) P.peer_prev = e->peer_prev_c[col];
        if (e->rank + 1 < e->nranks || periodic) P.peer_next = e->peer_next_c[col];
    }
    unsigned grid = plan.grid;
    if (part == 1) { P.strips = 2; P.strip_stride = plan.strips - 1; }
    if (part == 2) { P.strips = plan.strips - 2; P.strip_base = 1; }
    if (part != 0) grid = (unsigned)(((long long)e->n_rep * P.strips * plan.cpr + FAST_THREADS - 1) / FAST_THREADS);
    if (plan.fast) {
        fast_fn fn = pick_fast(col, plan.small, measure, e->boundary == BC_OPEN, fused_halo, plan.pair);
        if (e->opt_pdl) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FAST_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, fn, P);
        } else {
            fn<<<grid, FAST_THREADS, 0, stream>>>(P);
        }
    } else {
        if (measure) sweep_generic_kernel<true><<<plan.grid, 256, 0, e->stream>>>(P, col, e->boundary == BC_OPEN);
        else sweep_generic_kernel<false><<<plan.grid, 256, 0, e->stream>>>(P, col, e->boundary == BC_OPEN);
    }
    ++e->launches;
    return BC_OK;
}

static void raw_to_obs(const bc_engine* e, const unsigned long long* raw, int64_t sweep, bc_obs* o) {
    const int64_t N = (int64_t)e->L * e->L;  // global lattice (slab sums are all-reduced before this)
    const int64_t N1 = N / 2;  // colour-1 sites ((0,0) is colour 0)
    const int64_t A0 = (int64_t)raw[0], A1 = (int64_t)raw[1], A2 = (int64_t)raw[2], A3 = (int64_t)raw[3];
    const int64_t A4 = (int64_t)raw[4], A5 = (int64_t)raw[5];
    o->sweep = sweep;
    o->M = A0 + A2 - N;
    o->Q = (A1 + A3) - 2 * (A0 + A2) + N;
    o->B = A4 - 4 * A2 - A5 + 4 * N1;
}

static int fetch_pending(bc_engine* e, bc_obs* out) {
    const int64_t n = e->pending_meas;
    if (n == 0) return BC_OK;
    std::vector<unsigned long long> raw((size_t)n * e->n_rep * OBS_RAW);
    if (e->comm)  // slab: sum the per-rank partial sums (exact integers, order independent)
        BC_NCCL(e, g_nccl.AllReduce(e->d_obs, e->d_obs, raw.size(), ncclUint64, ncclSum, e->comm, e->stream));
    BC_CUDA(e, cudaMemcpyAsync(raw.data(), e->d_obs, raw.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    for (int64_t m = 0; m < n; ++m)
        for (int r = 0; r < e->n_rep; ++r)
            raw_to_obs(e, &raw[((size_t)m * e->n_rep + r) * OBS_RAW], e->pending_sweeps[m], &out[(size_t)m * e->n_rep + r]);
    return BC_OK;
}

// One half-sweep of a slab handle.  With overlap: the two boundary strips first, then their halo rows travel
// on the high-priority side stream while the interior strips are updated on the main stream; the next
// half-sweep waits for both.  (The boundary strips are whole strips of the ordinary decomposition, so the
// result is identical to the single-launch order: sites of one colour never read each other.)
static int slab_half(bc_engine* e, const LaunchPlan& plan, int col, bool measure, int64_t t, int64_t slot) {
    const long long ipr_part = (long long)2 * plan.cpr, ipr_int = (long long)(plan.strips - 2) * plan.cpr;
    const bool can_split = plan.fast && !plan.small && plan.strips >= 4 && (ipr_part % FAST_THREADS) == 0 &&
                           (ipr_int % FAST_THREADS) == 0 && e->opt_overlap;
    if (!can_split) {
        int rc = halo_wait(e, e->stream);
        if (rc != BC_OK) return rc;
        launch_half(e, plan, col, measure, t, slot);
        return exchange_halo(e, col);
    }
    // Split half-sweep.  Side stream (high priority): wait for the neighbours' ghost rows (P2P mode), update
    // the two boundary strips, send their fresh rows (peer stores + flag, or NCCL).  Main stream: the
    // interior strips, which read no ghost row, run concurrently.  Both start after the previous half-sweep
    // is complete on both streams; the next half-sweep starts after both parts.
    BC_CUDA(e, cudaEventRecord(e->ev_boundary, e->stream));
    BC_CUDA(e, cudaStreamWaitEvent(e->halo_stream, e->ev_boundary, 0));
    int rc = halo_wait(e, e->halo_stream);
    if (rc != BC_OK) return rc;
    launch_half(e, plan, col, measure, t, slot, nullptr, 1, e->halo_stream);
    rc = exchange_halo(e, col, e->halo_stream, e->p2p && e->opt_fused_halo);
    if (rc != BC_OK) return rc;
    BC_CUDA(e, cudaEventRecord(e->ev_halo, e->halo_stream));
    launch_half(e, plan, col, measure, t, slot, nullptr, 2, e->stream);
    BC_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_halo, 0));
    return BC_OK;
}

// (Re)capture the graph of `block` unmeasured sweeps for the current plan: 2*block kernels reading the
// device clock plus one node that advances it.
static int ensure_graph(bc_engine* e, const LaunchPlan& plan, int block) {
    if (e->graph_exec && e->graph_rows == plan.rows && e->graph_grid == plan.grid && e->graph_block == block) return BC_OK;
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    BC_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    const int64_t launches_before = e->launches;
    for (int s = 0; s < block; ++s) {
        launch_half(e, plan, 0, false, s, 0, e->d_clock);
        launch_half(e, plan, 1, false, s, 0, e->d_clock);
    }
    clock_add_kernel<<<1, 1, 0, e->stream>>>(e->d_clock, (int64_t)block);
    e->launches = launches_before;  // capturing is not launching
    cudaError_t st = cudaStreamEndCapture(e->stream, &graph);
    if (st != cudaSuccess) return fail(e, BC_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(st));
    st = cudaGraphInstantiate(&e->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (st != cudaSuccess) { e->graph_exec = nullptr; return fail(e, BC_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(st)); }
    e->graph_rows = plan.rows; e->graph_grid = plan.grid; e->graph_block = block;
    return BC_OK;
}

// ---- wavefront kernel: all half-sweeps of a run of unmeasured sweeps in one launch (L2-resident lattices) ----
static bool wave_eligible(const bc_engine* e, const LaunchPlan& plan, int64_t n) {
    if (e->opt_wave == 0 || e->ghost || !plan.fast || plan.small || (plan.rows & 3) || n < 2) return false;
    if (plan.cpr > FAST_THREADS && (plan.cpr % FAST_THREADS)) return false;  // blocks must not straddle strips
    return e->opt_wave > 0;
}

static int run_wave(bc_engine* e, const LaunchPlan& plan, int64_t t0, int64_t n) {
    const bool open = e->boundary == BC_OPEN;
    auto fn = open ? sweep_wave_kernel<true> : sweep_wave_kernel<false>;
    const int nblk = (int)(((long long)plan.strips * plan.cpr) / FAST_THREADS);
    const size_t need = 2 + (size_t)e->n_rep * nblk;
    if (e->wave_cap < need) {
        BC_CUDA(e, cudaStreamSynchronize(e->stream));
        cudaFree(e->d_wave);
        e->d_wave = nullptr; e->wave_cap = 0;
        BC_CUDA(e, cudaMalloc(&e->d_wave, need * sizeof(int)));
        e->wave_cap = need;
    }
    if (!e->wave_grid[open]) {
        int per_sm = 0, sms = 0;
        BC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, FAST_THREADS, 0));
        BC_CUDA(e, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
        e->wave_grid[open] = per_sm * sms;
    }
    // at most 2^30 tasks per launch
    const int64_t per_half = (int64_t)e->n_rep * nblk;
    int64_t done = 0;
    while (done < n) {
        int64_t chunk = n - done;
        if (2 * chunk * per_half > (1LL << 30)) chunk = std::max<int64_t>(1, (1LL << 29) / per_half);
        BC_CUDA(e, cudaMemsetAsync(e->d_wave, 0, need * sizeof(int), e->stream));
        WaveParams P;
        P.c0 = e->d_c[0]; P.c1 = e->d_c[1]; P.tab16 = e->d_tab16; P.tab31 = e->d_tab31; P.ties = e->d_ties;
        P.next_task = e->d_wave; P.abort = e->d_wave + 1; P.progress = e->d_wave + 2;
        P.t0 = t0 + done; P.n_half = (int)(2 * chunk);
        P.L = e->L; P.Wc = e->Wc; P.cpr = plan.cpr; P.rows_per_strip = plan.rows; P.strips = plan.strips;
        P.n_replicas = e->n_rep; P.rows_arr = e->rows_arr; P.nblk = nblk;
        P.key0 = (uint32_t)e->seed; P.key1 = (uint32_t)(e->seed >> 32);
        const long long tasks = 2 * chunk * per_half;
        const unsigned grid = (unsigned)std::min<long long>(tasks, e->wave_grid[open]);
        fn<<<grid, FAST_THREADS, 0, e->stream>>>(P);
        ++e->launches;
        BC_CUDA(e, cudaGetLastError());
        done += chunk;
    }
    return BC_OK;
}

// ---- persistent kernel: a whole run of unmeasured sweeps in one cooperative launch -------------------------
// rows per thread R for a member of a multi-handle launch: even, whole strips, whole CTAs per replica (!SMALL body)
static bool group_rows_valid(const bc_engine* e, int R) {
    const int cpr = e->L / 32;
    return e->L % 32 == 0 && R >= 2 && !(R & 1) && rows_valid(e->Ly, cpr, R) &&
           (((long long)(e->Ly / R) * cpr) % FAST_THREADS) == 0;
}

static int group_best_rows(const bc_engine* e, int target) {
    for (int R = target; R >= 2; R >>= 1)
        if (group_rows_valid(e, R)) return R;
    return 0;
}

struct GroupPlan {
    int n = 0;
    LaunchPlan plan[GROUP_MAX];
    unsigned cta_end[GROUP_MAX];
    bool pair = false, open = false;
};

static void fill_group_plan(bc_engine* const* es, int n, int target, GroupPlan* gp) {
    gp->n = n;
    gp->open = es[0]->boundary == BC_OPEN;
    unsigned long long ctas = 0;
    for (int g = 0; g < n; ++g) {
        const bc_engine* e = es[g];
        LaunchPlan& p = gp->plan[g];
        p.fast = true; p.small = false; p.pair = 0;
        p.cpr = e->L / 32;
        p.rows = group_best_rows(e, target);
        if (!p.rows) p.rows = group_best_rows(e, 32);
        p.strips = e->Ly / p.rows;
        p.grid = (unsigned)(((long long)e->n_rep * p.strips * p.cpr) / FAST_THREADS);
        ctas += p.grid;
        gp->cta_end[g] = (unsigned)std::min<unsigned long long>(ctas, 0xFFFFFFFFull);
    }
}

static int persist_capacity(bc_engine* e0) {
    const int open = e0->boundary == BC_OPEN;
    if (!e0->persist_cap[open]) {
        auto fn = open ? sweep_persist_kernel<true> : sweep_persist_kernel<false>;
        cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int per_sm = 0, sms = 0, coop = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, FAST_THREADS, 0) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e0->device) != cudaSuccess ||
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e0->device) != cudaSuccess || !coop)
            e0->persist_cap[open] = -1;
        else
            e0->persist_cap[open] = per_sm * sms;
    }
    return e0->persist_cap[open];
}

// Decide whether a run of unmeasured sweeps of these handles goes to the persistent kernel, and with which strips:
// the strip height that needs the fewest rounds of resident CTAs per half-sweep (a round costs its rows plus a
// prologue worth about two rows).  Only on request ("persist" 1): measured slower than graph replay + PDL.
static bool persist_plan(bc_engine* const* es, int n, int64_t n_sweeps, GroupPlan* gp) {
    bc_engine* e0 = es[0];
    if (e0->opt_persist <= 0 || n_sweeps < 2) return false;
    for (int g = 0; g < n; ++g)
        if (es[g]->ghost || es[g]->opt_force_generic || !group_best_rows(es[g], 2)) return false;
    const int cap = persist_capacity(e0);
    if (cap <= 0) return false;
    long long best_cost = -1;
    int best = 0;
    const int cand[5] = {32, 16, 8, 4, 2};
    for (int target : cand) {
        if (e0->opt_rows > 0 && target != e0->opt_rows) continue;
        fill_group_plan(es, n, target, gp);
        const long long blocks = gp->cta_end[n - 1];
        const long long cost = ((blocks + cap - 1) / cap) * (target + 2);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = target; }
    }
    if (!best) return false;
    fill_group_plan(es, n, best, gp);
    return true;
}

// n_sweeps unmeasured sweeps of every member, starting `s0` sweeps after the members' counters; on member 0's stream
static int run_persist(bc_engine* const* es, const GroupPlan& gp, int64_t s0, int64_t n_sweeps) {
    bc_engine* e0 = es[0];
    if (!e0->d_bar) BC_CUDA(e0, cudaMalloc(&e0->d_bar, 4 * sizeof(unsigned)));
    auto fn = gp.open ? sweep_persist_kernel<true> : sweep_persist_kernel<false>;
    const unsigned blocks = gp.cta_end[gp.n - 1];
    const unsigned grid = std::min<unsigned>(blocks, (unsigned)persist_capacity(e0));
    int64_t done = 0;
    while (done < n_sweeps) {
        const int64_t chunk = std::min<int64_t>(n_sweeps - done, 1 << 19);  // arrivals stay below 2^32
        PersistParams P;
        memset(&P, 0, sizeof(P));
        for (int col = 0; col < 2; ++col) {
            P.g[col].n = gp.n;
            for (int g = 0; g < gp.n; ++g) {
                P.g[col].p[g] = sweep_params(es[g], gp.plan[g], col, es[g]->t + s0 + done, 0, nullptr);
                P.g[col].cta_end[g] = gp.cta_end[g];
            }
        }
        P.bar = e0->d_bar;
        P.n_half = (int)(2 * chunk);
        BC_CUDA(e0, cudaMemsetAsync(e0->d_bar, 0, 4 * sizeof(unsigned), e0->stream));
        void* args[1] = {&P};
        BC_CUDA(e0, cudaLaunchCooperativeKernel((const void*)fn, dim3(grid), dim3(FAST_THREADS), args, 0, e0->stream));
        ++e0->launches;
        done += chunk;
    }
    return BC_OK;
}

// n unmeasured sweeps starting at sweep index t0
static int run_plain_sweeps(bc_engine* e, const LaunchPlan& plan, int64_t t0, int64_t n) {
    if (wave_eligible(e, plan, n)) return run_wave(e, plan, t0, n);
    if (plan.fast) {
        GroupPlan gp;
        bc_engine* one[1] = {e};
        if (persist_plan(one, 1, n, &gp)) return run_persist(one, gp, t0 - e->t, n);
    }
    int64_t done = 0;
    const int block = e->opt_graph_block;
    if (e->opt_use_graph && !e->ghost && block > 0 && n >= block) {
        int rc = ensure_graph(e, plan, block);
        if (rc != BC_OK) return rc;
        clock_set_kernel<<<1, 1, 0, e->stream>>>(e->d_clock, t0);
        ++e->launches;
        for (; done + block <= n; done += block) {
            BC_CUDA(e, cudaGraphLaunch(e->graph_exec, e->stream));
            e->launches += 2 * block + 1;
        }
    }
    for (; done < n; ++done) {
        if (e->ghost) {
            int rc = slab_half(e, plan, 0, false, t0 + done, 0);
            if (rc == BC_OK) rc = slab_half(e, plan, 1, false, t0 + done, 0);
            if (rc != BC_OK) return rc;
        } else {
            launch_half(e, plan, 0, false, t0 + done, 0);
            launch_half(e, plan, 1, false, t0 + done, 0);
        }
    }
    return BC_OK;
}

// ---- resident (shared-memory) path for small lattices -----------------------------------------------
struct ResidentPlan {
    bool ok = false;
    int T = 0, reps_per_cta = 0;
    size_t smem = 0;
    unsigned grid = 0;
};

static ResidentPlan resident_plan(const bc_engine* e) {
    ResidentPlan r;
    const int L = e->L;
    if (e->ghost || e->opt_resident == 0 || e->opt_force_generic) return r;
    const int items = L * (e->Wc / 16);  // 16-byte chunks of the padded rows (L % 32 != 0: the ANY instantiation)
    // 512 threads per replica shorten a half-sweep when there are too few replicas to fill the chip
    r.T = items <= 32 ? 32 : (items <= 128 ? 128 : ((items <= 256 || e->n_rep >= 2 * 148) ? 256 : 512));
    r.reps_per_cta = r.T < 128 ? 128 / r.T : 1;
    r.smem = (size_t)r.reps_per_cta * (2 * (size_t)L * e->Wc + 128) + (size_t)r.reps_per_cta * 32;
    if (r.smem > 227 * 1024) return r;
    // auto: always for the reference's sizes (L <= 128); for larger resident sizes only when there are
    // enough replicas to keep every SM busy with one CTA per replica
    if (e->opt_resident < 0 && L > 128 && e->n_rep < 2 * 148) return r;
    r.grid = (unsigned)((e->n_rep + r.reps_per_cta - 1) / r.reps_per_cta);
    r.ok = true;
    return r;
}

static int launch_resident(bc_engine* e, const ResidentPlan& rp, int64_t t0, int64_t n, int64_t measure_every) {
    const bool open = e->boundary == BC_OPEN;
    const bool any = e->L % 32 != 0;  // padded rows, wrap patches, masked stores
    auto fn = any ? (open ? sweep_resident_kernel<true, true> : sweep_resident_kernel<false, true>)
                  : (open ? sweep_resident_kernel<true, false> : sweep_resident_kernel<false, false>);
    if (!e->resident_attr_set[2 * any + open]) {
        BC_CUDA(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        e->resident_attr_set[2 * any + open] = true;
    }
    ResidentParams P;
    P.c0 = e->d_c[0]; P.c1 = e->d_c[1];
    P.tab16 = e->d_tab16; P.tab31 = e->d_tab31;
    P.obs = e->d_obs; P.ties = e->d_ties;
    P.t0 = t0; P.n_sweeps = n; P.measure_every = measure_every;
    P.L = e->L; P.Wc = e->Wc; P.cpr = e->Wc / 16; P.n_replicas = e->n_rep;
    P.threads_per_rep = rp.T; P.reps_per_cta = rp.reps_per_cta;
    P.key0 = (uint32_t)e->seed; P.key1 = (uint32_t)(e->seed >> 32);
    fn<<<rp.grid, rp.T * rp.reps_per_cta, rp.smem, e->stream>>>(P);
    ++e->launches;
    BC_CUDA(e, cudaGetLastError());
    return BC_OK;
}

// The measured sweeps of a call (t with (t+1) % measure_every == 0) and zeroed device slots for their sums.
static int prepare_obs(bc_engine* e, int64_t n_sweeps, int64_t measure_every) {
    e->pending_meas = 0;
    e->pending_sweeps.clear();
    if (measure_every > 0)
        for (int64_t t = e->t; t < e->t + n_sweeps; ++t)
            if ((t + 1) % measure_every == 0) e->pending_sweeps.push_back(t);
    const int64_t n_meas = (int64_t)e->pending_sweeps.size();
    if (n_meas > 0) {
        if ((size_t)n_meas > e->obs_slots_cap) {
            BC_CUDA(e, cudaStreamSynchronize(e->stream));
            if (e->d_obs) BC_CUDA(e, cudaFree(e->d_obs));
            e->d_obs = nullptr; e->obs_slots_cap = 0;
            BC_CUDA(e, cudaMalloc(&e->d_obs, (size_t)n_meas * e->n_rep * OBS_RAW * sizeof(unsigned long long)));
            e->obs_slots_cap = (size_t)n_meas;
        }
        BC_CUDA(e, cudaMemsetAsync(e->d_obs, 0, (size_t)n_meas * e->n_rep * OBS_RAW * sizeof(unsigned long long), e->stream));
    }
    return BC_OK;
}

extern "C" int64_t bc_sweep(bc_engine* e, int64_t n_sweeps, int64_t measure_every, bc_obs* out) {
    if (!e) return BC_ERR_ARG;
    if (n_sweeps < 0 || measure_every < 0) return fail(e, BC_ERR_ARG, "bc_sweep: negative count");
    for (int r = 0; r < e->n_rep; ++r)
        if (!e->params_set[r]) return fail(e, BC_ERR_STATE, "bc_sweep: bc_set_params not called for every replica");
    BC_CUDA(e, cudaSetDevice(e->device));
    const LaunchPlan plan = make_plan(e);

    int rc0 = prepare_obs(e, n_sweeps, measure_every);
    if (rc0 != BC_OK) return rc0;
    const int64_t n_meas = (int64_t)e->pending_sweeps.size();

    BC_CUDA(e, cudaEventRecord(e->ev0, e->stream));
    const ResidentPlan rp = resident_plan(e);
    if (rp.ok && n_sweeps > 0) {
        // whole call in one launch: lattices stay in shared memory, observables written per measured sweep
        int rc = launch_resident(e, rp, e->t, n_sweeps, measure_every);
        if (rc != BC_OK) return rc;
    }
    // runs of unmeasured sweeps (graph replay where possible), each followed by one measured sweep
    int64_t t = rp.ok ? e->t + n_sweeps : e->t;
    const int64_t t_end = e->t + n_sweeps;
    for (int64_t slot = 0; slot <= n_meas && !rp.ok; ++slot) {
        const int64_t t_meas = slot < n_meas ? e->pending_sweeps[(size_t)slot] : t_end;
        if (t_meas > t) {
            int rc = run_plain_sweeps(e, plan, t, t_meas - t);
            if (rc != BC_OK) return rc;
            t = t_meas;
        }
        if (slot < n_meas) {
            if (e->ghost) {
                int rc = slab_half(e, plan, 0, true, t, slot);
                if (rc == BC_OK) rc = slab_half(e, plan, 1, true, t, slot);
                if (rc != BC_OK) return rc;
            } else {
                launch_half(e, plan, 0, true, t, slot);
                launch_half(e, plan, 1, true, t, slot);
            }
            ++t;
        }
    }
    BC_CUDA(e, cudaEventRecord(e->ev1, e->stream));
    e->have_timing = true;
    BC_CUDA(e, cudaGetLastError());
    e->t += n_sweeps;
    e->pending_meas = n_meas;
    if (out && n_meas > 0) {
        int rc = fetch_pending(e, out);
        if (rc != BC_OK) return rc;
        e->pending_meas = 0;  // fetched (and, for slabs, already all-reduced in place)
        return n_meas * e->n_rep;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// group sweep: several handles (different L), one launch per half-sweep (sweep_group_kernel)
// ---------------------------------------------------------------------------------------------
typedef void (*group_fn)(const GroupParams);

template <int COL>
static group_fn pick_group(bool meas, bool open, bool pair) {
    if (pair) {
        if (meas) return open ? (group_fn)sweep_group_kernel<COL, true, true, 1> : (group_fn)sweep_group_kernel<COL, true, false, 1>;
        return open ? (group_fn)sweep_group_kernel<COL, false, true, 1> : (group_fn)sweep_group_kernel<COL, false, false, 1>;
    }
    if (meas) return open ? (group_fn)sweep_group_kernel<COL, true, true, 0> : (group_fn)sweep_group_kernel<COL, true, false, 0>;
    return open ? (group_fn)sweep_group_kernel<COL, false, true, 0> : (group_fn)sweep_group_kernel<COL, false, false, 0>;
}

// One half-sweep of every member: t = members' own counters + `s` (or + *t_ptr, the clock of member 0, in a graph)
static void launch_group_half(bc_engine* const* es, const GroupPlan& gp, int col, bool measure, int64_t s, int64_t slot,
                              const int64_t* t_ptr) {
    bc_engine* e0 = es[0];
    GroupParams G;
    memset(&G, 0, sizeof(G));
    G.n = gp.n;
    for (int g = 0; g < gp.n; ++g) {
        // with a clock the kernel adds member 0's absolute counter: pass this member's offset from it
        const int64_t t = t_ptr ? s + (es[g]->t - e0->t) : es[g]->t + s;
        G.p[g] = sweep_params(es[g], gp.plan[g], col, t, slot, t_ptr);
        G.cta_end[g] = gp.cta_end[g];
    }
    group_fn fn = col == 0 ? pick_group<0>(measure, gp.open, gp.pair) : pick_group<1>(measure, gp.open, gp.pair);
    const unsigned grid = gp.cta_end[gp.n - 1];
    if (e0->opt_pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FAST_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = e0->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, fn, G);
    } else {
        fn<<<grid, FAST_THREADS, 0, e0->stream>>>(G);
    }
    ++e0->launches;
}

// graph of `block` unmeasured group sweeps, cached on member 0 and keyed by everything baked into its kernels
static int ensure_group_graph(bc_engine* const* es, const GroupPlan& gp, int block) {
    bc_engine* e0 = es[0];
    std::vector<uint64_t> sig;
    sig.push_back((uint64_t)gp.n); sig.push_back((uint64_t)block); sig.push_back((uint64_t)gp.pair); sig.push_back((uint64_t)e0->opt_pdl);
    for (int g = 0; g < gp.n; ++g) {
        sig.push_back((uint64_t)(uintptr_t)es[g]->d_c[0]); sig.push_back((uint64_t)(uintptr_t)es[g]->d_tab16);
        sig.push_back((uint64_t)es[g]->seed); sig.push_back((uint64_t)gp.plan[g].rows);
        sig.push_back((uint64_t)(es[g]->t - e0->t)); sig.push_back((uint64_t)es[g]->n_rep);
    }
    if (e0->group_graph && sig == e0->group_sig) return BC_OK;
    if (e0->group_graph) { cudaGraphExecDestroy(e0->group_graph); e0->group_graph = nullptr; }
    cudaGraph_t graph = nullptr;
    BC_CUDA(e0, cudaStreamBeginCapture(e0->stream, cudaStreamCaptureModeThreadLocal));
    const int64_t launches_before = e0->launches;
    for (int s = 0; s < block; ++s) {
        launch_group_half(es, gp, 0, false, s, 0, e0->d_clock);
        launch_group_half(es, gp, 1, false, s, 0, e0->d_clock);
    }
    clock_add_kernel<<<1, 1, 0, e0->stream>>>(e0->d_clock, (int64_t)block);
    e0->launches = launches_before;  // capturing is not launching
    cudaError_t st = cudaStreamEndCapture(e0->stream, &graph);
    if (st != cudaSuccess) return fail(e0, BC_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(st));
    st = cudaGraphInstantiate(&e0->group_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (st != cudaSuccess) { e0->group_graph = nullptr; return fail(e0, BC_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(st)); }
    e0->group_sig = sig;
    return BC_OK;
}

extern "C" int64_t bc_sweep_group(bc_engine* const* engines, int n_engines, int64_t n_sweeps, int64_t measure_every) {
    if (!engines || n_engines < 1 || !engines[0]) return BC_ERR_ARG;
    bc_engine* e0 = engines[0];
    if (n_engines > GROUP_MAX) return fail(e0, BC_ERR_ARG, "bc_sweep_group: at most 8 handles per group");
    if (n_sweeps < 0 || measure_every < 0) return fail(e0, BC_ERR_ARG, "bc_sweep_group: negative count");
    for (int g = 0; g < n_engines; ++g) {
        bc_engine* e = engines[g];
        if (!e) return fail(e0, BC_ERR_ARG, "bc_sweep_group: null handle");
        for (int h = 0; h < g; ++h)
            if (engines[h] == e) return fail(e0, BC_ERR_ARG, "bc_sweep_group: a handle appears twice");
        if (e->device != e0->device || e->boundary != e0->boundary)
            return fail(e0, BC_ERR_ARG, "bc_sweep_group: the handles must share the device and the boundary condition");
        if (e->ghost) return fail(e0, BC_ERR_UNSUPPORTED, "bc_sweep_group: slab handles cannot join a group");
        if (e->L % 32 || !group_best_rows(e, 2 << 20))
            return fail(e0, BC_ERR_UNSUPPORTED, "bc_sweep_group: a member needs L % 32 == 0 and L >= 128 (use bc_sweep for it)");
        for (int r = 0; r < e->n_rep; ++r)
            if (!e->params_set[r]) return fail(e0, BC_ERR_STATE, "bc_sweep_group: bc_set_params not called for every replica");
        if (measure_every > 0 && e->t != e0->t)
            return fail(e0, BC_ERR_STATE, "bc_sweep_group: measuring needs equal sweep counters on all handles");
    }
    BC_CUDA(e0, cudaSetDevice(e0->device));

    // rows per thread: the longest strips (<= 32 rows, or member 0's "rows_per_thread") that still give the
    // whole group enough threads to fill the chip
    GroupPlan gp;
    gp.n = n_engines;
    gp.open = e0->boundary == BC_OPEN;
    int target = 2;
    if (e0->opt_rows > 0) {
        target = e0->opt_rows;
    } else {
        const long long want = 148LL * 768;
        for (int cand = 32; cand >= 2; cand >>= 1) {
            long long threads = 0;
            for (int g = 0; g < n_engines; ++g) {
                const int R = group_best_rows(engines[g], cand);  // > 0: 2 rows are valid (checked above) and cand is 2^k
                threads += (long long)engines[g]->n_rep * (engines[g]->Ly / R) * (engines[g]->L / 32);
            }
            target = cand;
            if (threads >= want) break;
        }
    }
    unsigned long long ctas = 0;
    long long sites_long = 0, sites_all = 0;  // the pair table pays for strips of >= pair_min_rows rows
    bool have_tab2 = true;
    for (int g = 0; g < n_engines; ++g) {
        const bc_engine* e = engines[g];
        LaunchPlan& p = gp.plan[g];
        p.fast = true; p.small = false;
        p.cpr = e->L / 32;
        p.rows = group_best_rows(e, target);
        if (!p.rows) p.rows = group_best_rows(e, 32);  // a "rows_per_thread" that fits no strip of this member
        p.strips = e->Ly / p.rows;
        p.grid = (unsigned)(((long long)e->n_rep * p.strips * p.cpr) / FAST_THREADS);
        ctas += p.grid;
        if (ctas > 0x7FFFFFFFull) return fail(e0, BC_ERR_UNSUPPORTED, "bc_sweep_group: too many CTAs for one launch");
        gp.cta_end[g] = (unsigned)ctas;
        if (!e->d_tab2) have_tab2 = false;
        sites_all += (long long)e->n_rep * e->Ly * e->L;
        if (p.rows >= e0->opt_pair_min_rows) sites_long += (long long)e->n_rep * e->Ly * e->L;
    }
    // one instantiation per launch: pair-table kernels when the members that profit from them hold most of the sites
    // (a member with shorter strips only stages the 11.4 KB table more often; the results do not depend on it)
    gp.pair = have_tab2 && e0->opt_pair_min_rows > 0 && 2 * sites_long >= sites_all;

    if (gp.pair) {  // 6 CTAs x 30 KB per SM need the large shared-memory carve-out
        static bool carveout_set = false;
        if (!carveout_set) {
            for (int m = 0; m < 2; ++m)
                for (int o = 0; o < 2; ++o) {
                    cudaFuncSetAttribute(pick_group<0>(m, o, true), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                    cudaFuncSetAttribute(pick_group<1>(m, o, true), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                }
            carveout_set = true;
        }
    }

    // join: every member's earlier work (uploads, parameter tables, its obs memset) precedes the group's launches
    for (int g = 0; g < n_engines; ++g) {
        bc_engine* e = engines[g];
        int rc = prepare_obs(e, n_sweeps, measure_every);
        if (rc != BC_OK) { if (e != e0) e0->err = e->err; return rc; }
        if (!e->ev_group) BC_CUDA(e0, cudaEventCreateWithFlags(&e->ev_group, cudaEventDisableTiming));
        if (g > 0) {
            BC_CUDA(e0, cudaEventRecord(e->ev_group, e->stream));
            BC_CUDA(e0, cudaStreamWaitEvent(e0->stream, e->ev_group, 0));
        }
    }
    const int64_t n_meas = (int64_t)e0->pending_sweeps.size();

    BC_CUDA(e0, cudaEventRecord(e0->ev0, e0->stream));
    int64_t done = 0;  // sweeps issued so far (offset from every member's counter)
    const int block = e0->opt_graph_block;
    for (int64_t slot = 0; slot <= n_meas; ++slot) {
        const int64_t until = slot < n_meas ? e0->pending_sweeps[(size_t)slot] - e0->t : n_sweeps;
        GroupPlan pp;
        if (until > done && persist_plan(engines, n_engines, until - done, &pp)) {
            int rc = run_persist(engines, pp, done, until - done);
            if (rc != BC_OK) return rc;
            done = until;
        }
        if (e0->opt_use_graph && block > 0 && until - done >= block) {
            int rc = ensure_group_graph(engines, gp, block);
            if (rc != BC_OK) return rc;
            clock_set_kernel<<<1, 1, 0, e0->stream>>>(e0->d_clock, e0->t + done);
            ++e0->launches;
            for (; done + block <= until; done += block) {
                BC_CUDA(e0, cudaGraphLaunch(e0->group_graph, e0->stream));
                e0->launches += 2 * block + 1;
            }
        }
        for (; done < until; ++done) {
            launch_group_half(engines, gp, 0, false, done, 0, nullptr);
            launch_group_half(engines, gp, 1, false, done, 0, nullptr);
        }
        if (slot < n_meas) {
            launch_group_half(engines, gp, 0, true, done, slot, nullptr);
            launch_group_half(engines, gp, 1, true, done, slot, nullptr);
            ++done;
        }
    }
    BC_CUDA(e0, cudaEventRecord(e0->ev1, e0->stream));
    e0->have_timing = true;
    BC_CUDA(e0, cudaGetLastError());
    // fork: whatever a member does next on its own stream comes after the group's sweeps
    BC_CUDA(e0, cudaEventRecord(e0->ev_group, e0->stream));
    for (int g = 0; g < n_engines; ++g) {
        bc_engine* e = engines[g];
        if (g > 0) BC_CUDA(e0, cudaStreamWaitEvent(e->stream, e0->ev_group, 0));
        e->t += n_sweeps;
        e->pending_meas = (int64_t)e->pending_sweeps.size();
    }
    return 0;
}

extern "C" int64_t bc_read_obs(bc_engine* e, bc_obs* out, int64_t max_records) {
    if (!e || !out) return BC_ERR_ARG;
    const int64_t n = e->pending_meas * e->n_rep;
    if (max_records < n) return fail(e, BC_ERR_ARG, "bc_read_obs: buffer too small");
    BC_CUDA(e, cudaSetDevice(e->device));
    int rc = fetch_pending(e, out);
    if (rc != BC_OK) return rc;
    e->pending_meas = 0;
    return n;
}

extern "C" int bc_measure(bc_engine* e, bc_obs* out) {
    if (!e || !out) return BC_ERR_ARG;
    BC_CUDA(e, cudaSetDevice(e->device));
    unsigned long long* d_raw = nullptr;
    const size_t bytes = (size_t)e->n_rep * OBS_RAW * sizeof(unsigned long long);
    BC_CUDA(e, cudaMalloc(&d_raw, bytes));
    cudaError_t st = cudaMemsetAsync(d_raw, 0, bytes, e->stream);
    const long long total = (long long)e->n_rep * e->Ly * e->L;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (st == cudaSuccess && halo_wait(e, e->stream) != BC_OK) st = cudaErrorUnknown;
    if (st == cudaSuccess) {
        measure_kernel<<<blocks, 256, 0, e->stream>>>(e->d_c[0], e->d_c[1], geom_of(e), e->n_rep, d_raw);
        ++e->launches;
        st = cudaGetLastError();
    }
    if (st == cudaSuccess && e->comm) {
        ncclResult_t ns = g_nccl.AllReduce(d_raw, d_raw, (size_t)e->n_rep * OBS_RAW, ncclUint64, ncclSum, e->comm, e->stream);
        if (ns != ncclSuccess) { cudaFree(d_raw); return fail(e, BC_ERR_NCCL, std::string("bc_measure: ") + g_nccl.GetErrorString(ns)); }
    }
    std::vector<unsigned long long> raw((size_t)e->n_rep * OBS_RAW);
    if (st == cudaSuccess) st = cudaMemcpyAsync(raw.data(), d_raw, bytes, cudaMemcpyDeviceToHost, e->stream);
    if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
    cudaFree(d_raw);
    if (st != cudaSuccess) return fail(e, BC_ERR_CUDA, std::string("bc_measure: ") + cudaGetErrorString(st));
    for (int r = 0; r < e->n_rep; ++r) raw_to_obs(e, &raw[(size_t)r * OBS_RAW], e->t - 1, &out[r]);
    return BC_OK;
}

// ---------------------------------------------------------------------------------------------
// cluster move: Swendsen-Wang on the +-1 spins, all replicas
// ---------------------------------------------------------------------------------------------
// shared by the cluster move and the cluster-size measurement: stage the lattices row-major, draw the bonds
// (threshold thr on Philox stream 3 at counter `cstep`), label connected components (root = smallest site)
// grid of the per-site cluster kernels: x = 256-site blocks of one replica, y = replicas (stride loop beyond 65535)
static dim3 site_grid(const bc_engine* e) {
    const long long N = (long long)e->L * e->L;
    return dim3((unsigned)((N + 255) / 256), (unsigned)std::min(e->n_rep, 65535));
}

static int label_clusters(bc_engine* e, uint32_t thr, int64_t cstep, bool always = false) {
    const long long N = (long long)e->L * e->L, total = N * e->n_rep;
    if (N > 0x7FFFFFFFLL) return fail(e, BC_ERR_UNSUPPORTED, "cluster labelling: site indices are 32-bit (L <= 46340)");
    int rc = ensure_stage(e, (size_t)total);
    if (rc != BC_OK) return rc;
    if (e->cluster_cap < (size_t)total) {
        BC_CUDA(e, cudaStreamSynchronize(e->stream));
        cudaFree(e->d_parent); cudaFree(e->d_bonds); cudaFree(e->d_csize);
        e->d_parent = nullptr; e->d_bonds = nullptr; e->d_csize = nullptr; e->cluster_cap = 0;
        BC_CUDA(e, cudaMalloc(&e->d_parent, (size_t)total * sizeof(int32_t)));
        BC_CUDA(e, cudaMalloc(&e->d_bonds, (size_t)total));
        e->cluster_cap = (size_t)total;
    }
    const uint32_t k0 = (uint32_t)e->seed, k1 = (uint32_t)(e->seed >> 32);
    const uint32_t c_lo = (uint32_t)((uint64_t)cstep & 0xFFFFFFFFu), c_hi = (uint32_t)((uint64_t)cstep >> 32);
    rc = stage_from_device(e, 0, e->n_rep);
    if (rc != BC_OK) return rc;
    const int tiles_x = (e->L + SW_TX - 1) / SW_TX, tiles_y = (e->L + SW_TY - 1) / SW_TY;
    const long long ctas = (long long)tiles_x * tiles_y * e->n_rep;
    if (ctas > 0x7FFFFFFFLL) return fail(e, BC_ERR_UNSUPPORTED, "cluster labelling: too many tiles for one launch");
    sw_local_kernel<<<(unsigned)ctas, SW_THREADS, 0, e->stream>>>(e->d_stage, e->d_bonds, e->d_parent, e->L, tiles_x, tiles_y,
                                                                  e->boundary == BC_OPEN, always ? 1 : 0, thr, k0, k1, c_lo, c_hi);
    sw_border_kernel<<<(unsigned)ctas, SW_TX + SW_TY, 0, e->stream>>>(e->d_bonds, e->d_parent, e->L, tiles_x, tiles_y);
    e->launches += 2;
    BC_CUDA(e, cudaGetLastError());
    return BC_OK;
}

static int bond_threshold_of(bc_engine* e, const char* who, uint32_t* thr) {
    for (int r = 0; r < e->n_rep; ++r) {
        if (!e->params_set[r]) return fail(e, BC_ERR_STATE, std::string(who) + ": bc_set_params not called for every replica");
        if (e->h_TJ[(size_t)r * 2] != e->h_TJ[0] || e->h_TJ[(size_t)r * 2 + 1] != e->h_TJ[1])
            return fail(e, BC_ERR_UNSUPPORTED, std::string(who) + ": all replicas of the handle must share (T, J)");
    }
    // bond probability p = 1 - exp(-2 J / T) (main.cpp:115) as a 32-bit threshold
    const double p = 1 - std::exp(-2 * (1 / e->h_TJ[0]) * e->h_TJ[1]);
    *thr = 0;
    if (p > 0) { double q = std::floor(p * 4294967296.0 + 0.5); *thr = q > 4294967295.0 ? 4294967295u : (uint32_t)q; }
    return BC_OK;
}

extern "C" int bc_cluster_update(bc_engine* e, int64_t n_updates) {
    if (!e || n_updates < 0) return BC_ERR_ARG;
    if (e->ghost) return fail(e, BC_ERR_UNSUPPORTED, "bc_cluster_update: not available on a slab handle");
    uint32_t thr = 0;
    int rc = bond_threshold_of(e, "bc_cluster_update", &thr);
    if (rc != BC_OK) return rc;
    BC_CUDA(e, cudaSetDevice(e->device));
    const long long N = (long long)e->L * e->L;
    const uint32_t k0 = (uint32_t)e->seed, k1 = (uint32_t)(e->seed >> 32);
    for (int64_t u = 0; u < n_updates; ++u) {
        rc = label_clusters(e, thr, e->cluster_step);
        if (rc != BC_OK) return rc;
        const uint32_t c_lo = (uint32_t)((uint64_t)e->cluster_step & 0xFFFFFFFFu), c_hi = (uint32_t)((uint64_t)e->cluster_step >> 32);
        // the bond bytes have served their purpose (border merge): the array now takes one flip bit per cluster root
        const dim3 rgrid((unsigned)((N + 8LL * SW_ROOT_SPAN - 1) / (8LL * SW_ROOT_SPAN)), (unsigned)std::min(e->n_rep, 65535));
        sw_rootbit_kernel<<<rgrid, 256, 0, e->stream>>>(e->d_stage, e->d_parent, e->d_bonds, e->L, e->n_rep, k0, k1, c_lo, c_hi);
        dim3 fgrid = site_grid(e);
        if ((N & 3) == 0) fgrid.x = (unsigned)((N / 4 + 255) / 256);  // four sites per thread
        sw_flip_kernel<<<fgrid, 256, 0, e->stream>>>(e->d_stage, e->d_parent, e->d_bonds, e->L, e->n_rep);
        e->launches += 2;
        BC_CUDA(e, cudaGetLastError());
        rc = stage_to_device(e, 0, e->n_rep);
        if (rc != BC_OK) return rc;
        ++e->cluster_step;
    }
    return BC_OK;
}

// Cluster-size moments of the current configurations (consumer gamma_nu.cpp:35-60):
// kind 0 = geometric "spin" clusters (every bond between equal non-zero neighbours, what export_clusters
// writes to spin/ with bond probability 1, main.cpp:369-370), kind 1 = FK clusters (bond probability
// 1 - exp(-2J/T), main.cpp:371-372; random bonds from the measurement counter `mstep`).
extern "C" int bc_cluster_moments(bc_engine* e, int kind, int64_t mstep, bc_cluster_stats* out) {
    if (!e || !out || (kind != 0 && kind != 1) || mstep < 0) return BC_ERR_ARG;
    if (e->ghost) return fail(e, BC_ERR_UNSUPPORTED, "bc_cluster_moments: not available on a slab handle");
    uint32_t thr = 0xFFFFFFFFu;
    bool always = true;
    if (kind == 1) {
        int rc0 = bond_threshold_of(e, "bc_cluster_moments", &thr);
        if (rc0 != BC_OK) return rc0;
        always = false;
    }
    BC_CUDA(e, cudaSetDevice(e->device));
    // measurement bonds use counters above 2^62 so that they never collide with the cluster move's stream
    int rc = label_clusters(e, thr, ((int64_t)1 << 62) + mstep, always);
    if (rc != BC_OK) return rc;
    const long long N = (long long)e->L * e->L, total = N * e->n_rep;
    if (!e->d_csize) BC_CUDA(e, cudaMalloc(&e->d_csize, e->cluster_cap * sizeof(int32_t)));
    if (!e->d_cmom) BC_CUDA(e, cudaMalloc(&e->d_cmom, (size_t)e->n_rep * 3 * sizeof(unsigned long long)));
    int32_t* d_size = e->d_csize;
    unsigned long long* d_out = e->d_cmom;
    cudaError_t st = cudaMemsetAsync(d_size, 0, (size_t)total * sizeof(int32_t), e->stream);
    if (st == cudaSuccess) st = cudaMemsetAsync(d_out, 0, (size_t)e->n_rep * 3 * sizeof(unsigned long long), e->stream);
    std::vector<unsigned long long> h((size_t)e->n_rep * 3);
    if (st == cudaSuccess) {
        dim3 cgrid = site_grid(e);
        if ((N & 3) == 0) cgrid.x = (unsigned)((N / 4 + 255) / 256);  // four sites per thread
        sw_compress_kernel<<<cgrid, 256, 0, e->stream>>>(e->d_parent, e->L, e->n_rep);
        sw_count_kernel<<<site_grid(e), 256, 0, e->stream>>>(e->d_stage, e->d_parent, d_size, e->L, e->n_rep);
        sw_moments_kernel<<<(unsigned)((total + SW_MOM_PER_CTA - 1) / SW_MOM_PER_CTA), 256, 0, e->stream>>>(d_size, e->L, e->n_rep, d_out);
        e->launches += 3;
        st = cudaGetLastError();
    }
    if (st == cudaSuccess) st = cudaMemcpyAsync(h.data(), d_out, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream);
    if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
    if (st != cudaSuccess) return fail(e, BC_ERR_CUDA, std::string("bc_cluster_moments: ") + cudaGetErrorString(st));
    for (int r = 0; r < e->n_rep; ++r) {
        out[r].sum_size_sq = (int64_t)h[(size_t)r * 3]; out[r].max_size = (int64_t)h[(size_t)r * 3 + 1];
        out[r].n_clusters = (int64_t)h[(size_t)r * 3 + 2];
    }
    return BC_OK;
}

extern "C" int bc_get_cluster_counter(const bc_engine* e, int64_t* c) {
    if (!e || !c) return BC_ERR_ARG;
    *c = e->cluster_step;
    return BC_OK;
}
extern "C" int bc_set_cluster_counter(bc_engine* e, int64_t c) {
    if (!e || c < 0) return BC_ERR_ARG;
    e->cluster_step = c;
    return BC_OK;
}

extern "C" int bc_sync(bc_engine* e) {
    if (!e) return BC_ERR_ARG;
    BC_CUDA(e, cudaSetDevice(e->device));
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    if (e->d_wave) {
        int flags[2] = {0, 0};
        BC_CUDA(e, cudaMemcpy(flags, e->d_wave, sizeof(flags), cudaMemcpyDeviceToHost));
        if (flags[1]) return fail(e, BC_ERR_STATE, "bc_sync: the wavefront kernel timed out waiting for a dependency");
    }
    if (e->d_bar) {
        unsigned bar[3] = {0, 0, 0};
        BC_CUDA(e, cudaMemcpy(bar, e->d_bar, sizeof(bar), cudaMemcpyDeviceToHost));
        if (bar[2]) return fail(e, BC_ERR_STATE, "bc_sync: the persistent kernel timed out at a grid barrier");
    }
    if (e->d_halo_err) {
        int err = 0;
        BC_CUDA(e, cudaMemcpy(&err, e->d_halo_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (err) return fail(e, BC_ERR_STATE, "bc_sync: a slab neighbour did not deliver its halo rows in time");
    }
    return BC_OK;
}

// ---- P2P halo: export / connect CUDA IPC handles -----------------------------------------------------
struct IpcBlob {
    cudaIpcMemHandle_t c[2];
    cudaIpcMemHandle_t flags;
};
static_assert(sizeof(IpcBlob) == 192, "three 64-byte CUDA IPC handles");

extern "C" int bc_slab_ipc_export(bc_engine* e, void* blob192) {
    if (!e || !blob192) return BC_ERR_ARG;
    if (!e->ghost) return fail(e, BC_ERR_STATE, "bc_slab_ipc_export: not a slab handle");
    BC_CUDA(e, cudaSetDevice(e->device));
    IpcBlob b;
    BC_CUDA(e, cudaIpcGetMemHandle(&b.c[0], e->d_c[0]));
    BC_CUDA(e, cudaIpcGetMemHandle(&b.c[1], e->d_c[1]));
    BC_CUDA(e, cudaIpcGetMemHandle(&b.flags, e->d_flags));
    std::memcpy(blob192, &b, sizeof(b));
    return BC_OK;
}

extern "C" int bc_slab_ipc_connect(bc_engine* e, const void* prev_blob192, const void* next_blob192) {
    if (!e) return BC_ERR_ARG;
    if (!e->ghost || e->nranks < 2) return fail(e, BC_ERR_STATE, "bc_slab_ipc_connect: needs a slab handle with n_ranks >= 2");
    BC_CUDA(e, cudaSetDevice(e->device));
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    const bool same = prev_blob192 && next_blob192 && std::memcmp(prev_blob192, next_blob192, sizeof(IpcBlob)) == 0;
    auto open3 = [&](const void* blob, uint8_t* (&c)[2], long long*& flags, int slot) -> int {
        IpcBlob b;
        std::memcpy(&b, blob, sizeof(b));
        void* p[3] = {nullptr, nullptr, nullptr};
        BC_CUDA(e, cudaIpcOpenMemHandle(&p[0], b.c[0], cudaIpcMemLazyEnablePeerAccess));
        BC_CUDA(e, cudaIpcOpenMemHandle(&p[1], b.c[1], cudaIpcMemLazyEnablePeerAccess));
        BC_CUDA(e, cudaIpcOpenMemHandle(&p[2], b.flags, cudaIpcMemLazyEnablePeerAccess));
        c[0] = (uint8_t*)p[0]; c[1] = (uint8_t*)p[1]; flags = (long long*)p[2];
        for (int i = 0; i < 3; ++i) e->ipc_opened[slot * 3 + i] = p[i];
        return BC_OK;
    };
    if (prev_blob192) { int rc = open3(prev_blob192, e->peer_prev_c, e->peer_prev_flags, 0); if (rc != BC_OK) return rc; }
    if (next_blob192) {
        if (same) { e->peer_next_c[0] = e->peer_prev_c[0]; e->peer_next_c[1] = e->peer_prev_c[1]; e->peer_next_flags = e->peer_prev_flags; }
        else { int rc = open3(next_blob192, e->peer_next_c, e->peer_next_flags, 1); if (rc != BC_OK) return rc; }
    }
    e->p2p = true;
    return BC_OK;
}

// ---------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------
extern "C" int64_t bc_launch_count(const bc_engine* e) { return e ? e->launches : (int64_t)BC_ERR_ARG; }

extern "C" int bc_last_sweep_ms(bc_engine* e, float* ms) {
    if (!e || !ms) return BC_ERR_ARG;
    if (!e->have_timing) return fail(e, BC_ERR_STATE, "bc_last_sweep_ms: no bc_sweep yet");
    BC_CUDA(e, cudaSetDevice(e->device));
    BC_CUDA(e, cudaEventSynchronize(e->ev1));
    BC_CUDA(e, cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return BC_OK;
}

extern "C" int bc_timer_start(bc_engine* e) {
    if (!e) return BC_ERR_ARG;
    BC_CUDA(e, cudaSetDevice(e->device));
    BC_CUDA(e, cudaEventRecord(e->tm0, e->stream));
    return BC_OK;
}

extern "C" int bc_timer_stop(bc_engine* e, float* ms) {
    if (!e || !ms) return BC_ERR_ARG;
    BC_CUDA(e, cudaSetDevice(e->device));
    BC_CUDA(e, cudaEventRecord(e->tm1, e->stream));
    BC_CUDA(e, cudaEventSynchronize(e->tm1));
    BC_CUDA(e, cudaEventElapsedTime(ms, e->tm0, e->tm1));
    return BC_OK;
}

extern "C" int bc_set_option(bc_engine* e, const char* name, int64_t value) {
    if (!e || !name) return BC_ERR_ARG;
    const std::string n(name);
    if (n == "rows_per_thread") { e->opt_rows = (int)value; return BC_OK; }
    if (n == "swp_min_rows") { e->opt_swp_min_rows = (int)value; if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; } return BC_OK; }
    if (n == "pair_min_rows") { e->opt_pair_min_rows = (int)value; if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; } return BC_OK; }
    if (n == "wave") { e->opt_wave = (int)value; return BC_OK; }
    if (n == "pdl") { e->opt_pdl = value != 0; if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; } return BC_OK; }
    if (n == "fused_halo") { e->opt_fused_halo = value != 0; return BC_OK; }
    if (n == "overlap") { e->opt_overlap = value != 0; return BC_OK; }
    if (n == "resident") { e->opt_resident = (int)value; return BC_OK; }
    if (n == "persist") { e->opt_persist = (int)value; return BC_OK; }
    if (n == "use_graph") { e->opt_use_graph = value != 0; return BC_OK; }
    if (n == "graph_block") { if (value < 1 || value > 1024) return fail(e, BC_ERR_ARG, "graph_block out of range"); e->opt_graph_block = (int)value; return BC_OK; }
    if (n == "force_generic") { e->opt_force_generic = value != 0; return BC_OK; }
    return fail(e, BC_ERR_ARG, "bc_set_option: unknown option " + n);
}

extern "C" int bc_tie_count(bc_engine* e, int64_t* ties) {
    if (!e || !ties) return BC_ERR_ARG;
    BC_CUDA(e, cudaSetDevice(e->device));
    unsigned long long v = 0;
    BC_CUDA(e, cudaMemcpyAsync(&v, e->d_ties, sizeof(v), cudaMemcpyDeviceToHost, e->stream));
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    *ties = (int64_t)v;
    return BC_OK;
}

extern "C" int bc_describe_plan(bc_engine* e, int* fast, int* rows_per_thread, int* grid, int* block) {
    if (!e) return BC_ERR_ARG;
    const LaunchPlan p = make_plan(e);
    if (resident_plan(e).ok) {  // 2 = shared-memory-resident kernel
        const ResidentPlan r = resident_plan(e);
        if (fast) *fast = 2;
        if (rows_per_thread) *rows_per_thread = 0;
        if (grid) *grid = (int)r.grid;
        if (block) *block = r.T * r.reps_per_cta;
        return BC_OK;
    }
    if (fast) *fast = p.fast ? (p.pair ? 2 + p.pair : 1) : 0;  // 3 = pair table, 4 = + pipelined draws
    if (rows_per_thread) *rows_per_thread = p.rows;
    if (grid) *grid = (int)p.grid;
    if (block) *block = p.fast ? FAST_THREADS : 256;
    return BC_OK;
}
