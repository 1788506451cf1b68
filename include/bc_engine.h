/*
 * bc_engine.h -- C ABI of the B200 Blume-Capel checkerboard Metropolis engine
 *                (libbc_engine.so, built from blume_capel_b200/csrc/).
 *
 * The reference (j2apps/blume_capel) has no plugin / FFI layer: its hot path is the
 * in-process call   refill_random(); step(lattice);   made from main()'s two driver loops
 * (/root/reference/main.cpp:342-348 burn-in, main.cpp:361-368 collection) with the
 * parameters passed as globals (B, D, J at main.cpp:24) and the lattice owned by main()
 * (main.cpp:335).  This header is the boundary a maintainer binds instead of those calls;
 * each entry point names the reference code it replaces.  INTEGRATION.md shows the patch.
 *
 * Conventions
 *   - every function returns 0 on success, a negative bc_status on failure; the text of
 *     the last failure on a handle is bc_last_error(handle) (bc_create failures:
 *     bc_last_error(NULL)).  No C++ exception crosses this boundary.
 *   - the caller owns every host buffer; the engine owns all device memory and one stream.
 *   - a handle is bound to ONE GPU and is not thread-safe; distinct handles may be driven
 *     from distinct host threads / processes (one process per GPU in multi-GPU runs).
 *   - bc_sweep is asynchronous on the engine's stream unless it has to return observables;
 *     bc_sync / bc_download / bc_read_obs wait for it.
 *   - there is NO CPU fallback: without a CUDA device bc_create fails with BC_ERR_CUDA.
 *   - spins on the host side are int8 in {-1,0,+1}, row-major idx = x + y*L, exactly the
 *     reference layout (main.cpp:98-101, int instead of int8).
 */
#ifndef BC_ENGINE_H
#define BC_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BC_ABI_VERSION 2

typedef struct bc_engine bc_engine; /* opaque */

enum bc_status {
    BC_OK = 0,
    BC_ERR_ARG = -1,     /* bad argument (odd L on a torus, replica out of range, ...) */
    BC_ERR_CUDA = -2,    /* CUDA runtime error (text in bc_last_error) */
    BC_ERR_NCCL = -3,    /* NCCL error */
    BC_ERR_STATE = -4,   /* call not valid in this state (parameters not set, ...) */
    BC_ERR_UNSUPPORTED = -5
};

enum bc_boundary { BC_PERIODIC = 0, BC_OPEN = 1 };
enum bc_storage { BC_STORAGE_INT8 = 0 };

/* Exact integer observables of one replica after one measured sweep.
 * E = -J*B + D*Q (H of main.cpp:105-106), m = M/N, q = Q/N. */
typedef struct {
    int64_t sweep; /* global index of the sweep just completed */
    int64_t M;     /* sum_i s_i                                  */
    int64_t Q;     /* sum_i s_i^2                                */
    int64_t B;     /* sum over bonds s_i s_j, each bond (+x,+y of every site) once */
} bc_obs;

/* ---- lifetime ---------------------------------------------------------------------- */

/* Replaces: `int lattice[N]` + globals set up in main() (main.cpp:303-335).
 * L: linear size (run time, not the -DL_MACRO compile-time constant of main.cpp:14-16,39);
 * periodic needs even L.  n_replicas independent lattices of the same L are updated per
 * launch (replica r uses Philox key (seed_lo, seed_hi ^ r)). */
int bc_create(bc_engine** out, int L, int boundary, int n_replicas, int storage, int device, uint64_t seed);
void bc_destroy(bc_engine* e);
const char* bc_last_error(const bc_engine* e);
int bc_abi_version(void);

/* ---- slab decomposition (lattices too large for one GPU; one process per GPU) ---------------------
 * New relative to the reference, whose only multi-process mode is independent Slurm array tasks
 * (job_scripts/runner.sh:19).  The L x L lattice is split into n_ranks row slabs of L/n_ranks rows;
 * after every half-sweep each rank sends the freshly updated colour of its first and last owned row to
 * its ring neighbours over NCCL (ncclSend/ncclRecv on the engine stream).  Results are bit-identical
 * to a single-GPU run because the RNG counter is a function of global site coordinates.
 * bc_slab_unique_id: rank 0 fills 128 bytes (an ncclUniqueId) that the caller distributes to all ranks
 * (bench.py broadcasts it through its process group).  n_ranks == 1 needs no id and exchanges with itself.
 * On a slab handle bc_upload/bc_download move the LOCAL rows (bc_local_rows), observables are global. */
int bc_slab_unique_id(void* id128);
int bc_create_slab(bc_engine** out, int L, int boundary, int n_replicas, int storage, int device, uint64_t seed,
                   int n_ranks, int rank, const void* nccl_id128);
int bc_local_rows(const bc_engine* e, int* rows, int* first_global_row);
/* Halo exchange over peer memory instead of NCCL send/recv: every rank exports 192 bytes (CUDA IPC handles of
 * its two colour arrays and its flag words), the caller hands each rank the blobs of its ring neighbours
 * (NULL at an open edge), and from then on a half-sweep stores its boundary rows straight into the
 * neighbours' ghost rows over NVLink and publishes an epoch flag; the consumer waits on the GPU.  One
 * process per GPU, never two ranks on one GPU.  Collective: all ranks connect at the same point. */
int bc_slab_ipc_export(bc_engine* e, void* blob192);
int bc_slab_ipc_connect(bc_engine* e, const void* prev_blob192, const void* next_blob192);
/* Cluster move on slabs (n_ranks >= 2, at most 8): every rank exports 128 bytes (CUDA IPC handles of its piece of the
 * global union-find over cluster roots and of its tiles' perimeter roots), the caller all-gathers them and hands every
 * rank the n_ranks blobs in rank order.  Collective; afterwards bc_cluster_update works on the decomposition. */
int bc_slab_cluster_export(bc_engine* e, void* blob128);
int bc_slab_cluster_connect(bc_engine* e, const void* blobs);

/* ---- parameters -------------------------------------------------------------------- */

/* Replaces: B = 1/atof(argv[5]); D = atof(argv[6]); J = atof(argv[7]) (main.cpp:321-323).
 * Rebuilds the 54-entry integer acceptance table for `replica` (-1 = all replicas). */
int bc_set_params(bc_engine* e, int replica, double T, double D, double J);

/* The table the kernels use: thr31[54], index 18*(s+1) + 9*flip + (nb+4); accept iff
 * u31 < thr31.  Lets a test compare it with the literal expression of main.cpp:105-109. */
int bc_get_table(const bc_engine* e, int replica, uint32_t thr31[54]);

/* ---- lattice state ----------------------------------------------------------------- */

/* Replaces generate_lattice() (main.cpp:166-171): iid uniform {-1,0,1}, from the Philox
 * init stream (reproducible, unlike the reference's random_device seeding).  replica -1 = all. */
int bc_init_random(bc_engine* e, int replica);

/* Replaces get_lattice_from_burn()'s result (main.cpp:281-299) / export_clusters()' input
 * (main.cpp:232): whole-lattice transfer in the reference layout.  spins: L*L int8. */
int bc_upload(bc_engine* e, int replica, const int8_t* spins);
int bc_download(bc_engine* e, int replica, int8_t* spins);
/* all replicas at once, spins: n_replicas * L*L int8, replica-major */
int bc_upload_all(bc_engine* e, const int8_t* spins);
int bc_download_all(bc_engine* e, int8_t* spins);
/* Asynchronous variants (enqueue on the engine stream and return): `spins` should be pinned host memory
 * and must stay valid -- and, for the upload, unmodified -- until the next bc_sync on this handle.  With
 * two handles a caller overlaps the transfers of one batch with the sweeps of another. */
int bc_upload_all_async(bc_engine* e, const int8_t* spins);
int bc_download_all_async(bc_engine* e, int8_t* spins);

/* 2-bit wire format: whole-lattice transfers are PCIe-bound at one byte per site, so the same transfers are also
 * offered packed, four sites per byte: row-major like the reference's lattice, byte b of a row holds the codes
 * spin + 1 of sites x = 4b .. 4b+3 in bits 0-1, 2-3, 4-5, 6-7, rows padded to whole bytes.  The conversion to and from
 * the engine's layout runs on the device.  replica -1 = all replicas (bc_packed_size(e, n_replicas) bytes); async != 0:
 * enqueue and return like the *_async calls above (pinned memory, valid until bc_sync).  bc_pack_spins /
 * bc_unpack_spins convert between int8 spins and the wire format on the host (rows of L sites). */
int64_t bc_packed_size(const bc_engine* e, int n_lattices);
int bc_upload_packed(bc_engine* e, int replica, const uint8_t* packed, int async);
int bc_download_packed(bc_engine* e, int replica, uint8_t* packed, int async);
int bc_pack_spins(const int8_t* spins, int L, int64_t rows, uint8_t* packed);
int bc_unpack_spins(const uint8_t* packed, int L, int64_t rows, int8_t* spins);

/* Sweep counter = the full RNG state besides the seed (Philox is counter based), so
 * (lattice, seed, counter) is an exact checkpoint. */
int bc_get_sweep_counter(const bc_engine* e, int64_t* t);
int bc_set_sweep_counter(bc_engine* e, int64_t t);

/* ---- the hot path ------------------------------------------------------------------ */

/* Replaces the loop body `refill_random(); step(lattice);` (main.cpp:343-347, 363-367) for
 * the Metropolis part of step() (main.cpp:159-161): n_sweeps checkerboard sweeps (one sweep
 * = N attempts = colour 0 then colour 1) of every replica.  If measure_every > 0, after every
 * sweep t with (t+1) % measure_every == 0 the observables of each replica are accumulated in
 * the same pass; they are written to `out` (host, caller-owned, n_meas*n_replicas records,
 * measurement-major then replica) if out != NULL, and the call then waits for completion.
 * Returns the number of records written (>= 0) or a negative status. */
int64_t bc_sweep(bc_engine* e, int64_t n_sweeps, int64_t measure_every, bc_obs* out);

/* The same sweeps for SEVERAL handles of one device in ONE kernel launch per half-sweep: a finite-size-scaling
 * ladder (BASELINE config 3: L = 256 ... 2048 at the same (T, D) points) or the open-boundary size series (config 5)
 * as the reference runs them, one process per size (job_scripts/runner.sh:14-19), but without a launch -- and a
 * launch tail -- per size: the CTAs of all members form one grid.  Members keep their own size, replicas, tables,
 * seed and sweep counter, and the results are bit-identical to bc_sweep on each handle.  Requirements: 1..8 distinct
 * handles, same device and boundary, no slab handles, every L a multiple of 32 and >= 128 (smaller lattices belong
 * to bc_sweep's shared-memory-resident kernel); with measure_every > 0 all sweep counters must be equal.  The
 * launches go to member 0's stream (its options "rows_per_thread", "pair_min_rows", "use_graph", "graph_block",
 * "pdl" apply; bc_launch_count / bc_last_sweep_ms / bc_last_error report on member 0); the other members' streams are
 * ordered before and after, so any later call on any member sees the swept lattices.  Observables: bc_read_obs on
 * each member.  Returns 0 or a negative status. */
int64_t bc_sweep_group(bc_engine* const* engines, int n_engines, int64_t n_sweeps, int64_t measure_every);

/* Cluster move (the reference interleaves one Wolff cluster per 3L Metropolis attempts,
 * main.cpp:113-154,158-163): n_updates Swendsen-Wang updates of every replica -- bonds between equal non-zero
 * neighbours with probability 1 - exp(-2J/T) (main.cpp:115; zeros never join, main.cpp:123-130,144), every
 * cluster flipped with probability 1/2.  Same stationary distribution as the reference's Wolff move, all
 * clusters at once instead of one.  Every replica uses its own bond probability (a handle may hold a (T, D) grid).
 * On a slab of a multi-rank decomposition the move is collective (every rank calls it) and needs
 * bc_slab_cluster_connect; clusters that cross slab borders are merged through peer memory, the result is identical
 * to the single-GPU move.  Has its own RNG counter. */
int bc_cluster_update(bc_engine* e, int64_t n_updates);
int bc_get_cluster_counter(const bc_engine* e, int64_t* c);
int bc_set_cluster_counter(bc_engine* e, int64_t c);

/* Cluster-size moments of the current configurations (what gamma_nu.cpp:35-60 derives from
 * the line lengths of a cluster file): kind 0 = geometric "spin" clusters (the content of spin/ files, bond
 * probability 1, main.cpp:369-370), kind 1 = FK clusters (bond probability 1 - exp(-2J/T), main.cpp:371-372,
 * bonds drawn from Philox stream 3 at the measurement counter mstep).  out: n_replicas records. */
typedef struct {
    int64_t sum_size_sq; /* sum over clusters of size^2 */
    int64_t max_size;    /* size of the largest cluster */
    int64_t n_clusters;
} bc_cluster_stats;
int bc_cluster_moments(bc_engine* e, int kind, int64_t mstep, bc_cluster_stats* out);

/* Gap-size histogram of the current configurations on the GPU: what
 * corner_contribution/gap_size_statistics.cpp:58-81,100-144 derives from the cluster file of one sample, quirks
 * included (the walk starts at the ROW of a cluster's first site and never records that site, :104-106), for the
 * spin (kind 0) or FK (kind 1, bonds at measurement counter mstep) clusters of every replica.  ADDS one sample per
 * replica to hist[replica][bins]: variant 0 = the reference's (gaps folded to min(gap, L - gap), wrap-around gap,
 * L/2 bins, bin g-1 counts gaps of size g); variant 1 = open-boundary variant, new relative to the reference
 * (unfolded gaps, no wrap-around gap, L-1 bins).  Input of gap_size_analysis.py:5-11. */
int bc_gap_histogram(bc_engine* e, int kind, int64_t mstep, int variant, int64_t* hist);

/* Cluster lists of one replica, formed on the GPU, for the dump writer: replaces form_clusters + the per-cluster
 * sort of export_clusters (main.cpp:179-230, 236-239).  Clusters come in the order of their smallest site and
 * their members ascend, exactly the order of the reference's files.  sites: members of all clusters back to back
 * (capacity L*L); starts[c]: offset of cluster c in sites (capacity L*L + 1; starts[n_clusters] = n_sites);
 * signs[c] = +1 / -1 (capacity L*L).  kind / mstep as in bc_cluster_moments (spin/ files: kind 0; fk/ files: kind 1
 * with mstep = the sample index).  The labelling and the sort are cached per (lattice state, kind, mstep). */
int bc_cluster_lists(bc_engine* e, int replica, int kind, int64_t mstep, uint32_t* sites, uint32_t* starts,
                     int8_t* signs, int64_t* n_sites, int64_t* n_clusters);

/* Fetch the records of the last bc_sweep that was called with out == NULL (waits for it).
 * Returns the number of records written or a negative status. */
int64_t bc_read_obs(bc_engine* e, bc_obs* out, int64_t max_records);

/* Observables of the current configuration (separate reduction kernel, no update). */
int bc_measure(bc_engine* e, bc_obs* out /* n_replicas records */);

int bc_sync(bc_engine* e);

/* ---- introspection for benchmarks / tests ------------------------------------------ */

/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
int64_t bc_launch_count(const bc_engine* e);
/* Device time in ms of the last bc_sweep call, CUDA events on the engine's stream. */
int bc_last_sweep_ms(bc_engine* e, float* ms);
/* CUDA-event stopwatch on the engine's stream (bench.py times K steps with it):
 * bc_timer_start records an event, bc_timer_stop records another, waits, returns the ms between. */
int bc_timer_start(bc_engine* e);
int bc_timer_stop(bc_engine* e, float* ms);
/* Tuning / test knobs (name, value); unknown names give BC_ERR_ARG.  Results never depend on them:
 *   "rows_per_thread" n   rows each thread of the streaming kernel marches over (0 = auto)
 *   "pair_min_rows" n     streaming kernel: use the 54 x 54 pair table (one threshold lookup per two sites)
 *                         when a thread marches over >= n rows (default 8; 0 = never)
 *   "resident" -1|0|1     shared-memory-resident kernel for L <= 448, any L (auto: L <= 128 / off / force)
 *   "use_graph" 0|1, "graph_block" n   CUDA-graph replay of blocks of n unmeasured sweeps
 *   "pdl" 0|1             programmatic dependent launch between half-sweeps
 *   "wave" 0|1            experimental single-launch wavefront kernel (off by default)
 *   "persist" 0|1         experimental persistent kernel: a run of unmeasured sweeps in one cooperative launch with
 *                         grid barriers between half-sweeps (off by default: slower than graph replay)
 *   "overlap" 0|1         slab handles: boundary strips + halo exchange concurrent with the interior
 *   "fused_halo" 0|1      peer-memory slabs: halo stores issued by the boundary-strip kernel itself
 *   "force_generic" 0|1   one-thread-per-site kernel for every size (second implementation, for tests) */
int bc_set_option(bc_engine* e, const char* name, int64_t value);
/* The launch geometry bc_sweep would use now (any pointer may be NULL).  *fast: 0 = one thread per site,
 * 1 = streaming kernel, 2 = shared-memory-resident kernel, 3 = streaming kernel with the pair table. */
int bc_describe_plan(bc_engine* e, int* fast, int* rows_per_thread, int* grid, int* block);
/* Ties (coarse 15-bit draw == coarse threshold) resolved by the fine stream so far. */
int bc_tie_count(bc_engine* e, int64_t* ties);

#ifdef __cplusplus
}
#endif
#endif /* BC_ENGINE_H */
