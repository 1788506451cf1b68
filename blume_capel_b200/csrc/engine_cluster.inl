// engine_cluster.inl -- cluster move, cluster labels and cluster-size moments, bc_sync.
// Part of bc_engine.cu (included there once; not a stand-alone translation unit).

// ---------------------------------------------------------------------------------------------
// cluster move: Swendsen-Wang on the +-1 spins, all replicas, straight on the compact colour arrays
// ---------------------------------------------------------------------------------------------
static int cluster_alloc(bc_engine* e) {
    if (e->cl_par) return BC_OK;
    const size_t tiles_x = (size_t)(e->L + CL_T - 1) / CL_T, tiles_y = (size_t)(e->Ly + CL_T - 1) / CL_T;
    const size_t n_tiles = tiles_x * tiles_y * (size_t)e->n_rep;
    if (n_tiles > 0x7FFFFFFFull) return fail(e, BC_ERR_UNSUPPORTED, "cluster labelling: too many tiles for one launch");
    const size_t own_sites = (size_t)e->Ly * (size_t)e->L;
    BC_CUDA(e, cudaSetDevice(e->device));
    BC_CUDA(e, cudaMalloc(&e->cl_par, n_tiles * CL_SITES * sizeof(uint16_t)));
    BC_CUDA(e, cudaMalloc(&e->cl_hb, n_tiles * CL_T * sizeof(uint64_t)));
    BC_CUDA(e, cudaMalloc(&e->cl_pm, n_tiles * 2 * CL_T * sizeof(uint64_t)));
    BC_CUDA(e, cudaMalloc(&e->cl_vbb, n_tiles * sizeof(uint64_t)));
    BC_CUDA(e, cudaMalloc(&e->cl_proot, n_tiles * 256 * sizeof(uint16_t)));
    BC_CUDA(e, cudaMalloc(&e->cl_gparent, (size_t)e->n_rep * own_sites * sizeof(uint32_t)));
    return BC_OK;
}

static int cluster_args(bc_engine* e, const char* who, bool random_bonds, int64_t cstep, ClArgs* A) {
    const bool multi = e->ghost && e->nranks > 1;
    if (multi && !e->cl_connected)
        return fail(e, BC_ERR_STATE, std::string(who) + ": a slab of a multi-rank decomposition needs bc_slab_cluster_connect first");
    for (int r = 0; r < e->n_rep; ++r)
        if (random_bonds && !e->params_set[r]) return fail(e, BC_ERR_STATE, std::string(who) + ": bc_set_params not called for every replica");
    BC_CUDA(e, cudaSetDevice(e->device));
    int rc = cluster_alloc(e);
    if (rc != BC_OK) return rc;
    memset(A, 0, sizeof(*A));
    A->g.L = e->L; A->g.Wc = e->Wc; A->g.rows_arr = e->rows_arr; A->g.row_first = e->row_first;
    A->g.open = e->boundary == BC_OPEN;
    A->g.Ly = e->Ly; A->g.y_glob0 = e->y_glob0; A->g.multi = multi ? 1 : 0;
    A->g.tiles_x = (e->L + CL_T - 1) / CL_T; A->g.tiles_y = (e->Ly + CL_T - 1) / CL_T;
    A->c0 = e->d_c[0]; A->c1 = e->d_c[1];
    A->out.par = e->cl_par; A->out.hb = e->cl_hb; A->out.pm = e->cl_pm; A->out.vbb = e->cl_vbb; A->out.proot = e->cl_proot; A->out.gparent = e->cl_gparent;
    A->rng.thr = random_bonds ? e->d_bond_thr : nullptr;  // per replica: p = 1 - exp(-2 J / T) (main.cpp:115)
    A->rng.k0 = (uint32_t)e->seed; A->rng.k1 = (uint32_t)(e->seed >> 32);
    A->rng.c_lo = (uint32_t)((uint64_t)cstep & 0xFFFFFFFFu); A->rng.c_hi = (uint32_t)((uint64_t)cstep >> 32);
    A->n_tiles = (uint32_t)((size_t)A->g.tiles_x * A->g.tiles_y * e->n_rep);
    A->rep_stride = (size_t)e->Ly * (size_t)e->L;
    if (multi) {
        for (int k = 0; k < e->nranks; ++k) A->gbase[k] = k == e->rank ? e->cl_gparent : e->cl_peer_gparent[k];
        A->sites_per_rank = (uint32_t)((size_t)e->Ly * (size_t)e->L);
        A->proot_below = e->cl_peer_proot_below;
        A->y_glob0_below = ((e->rank + 1) % e->nranks) * e->Ly;
    } else {
        A->gbase[0] = e->cl_gparent;
        A->sites_per_rank = 0;
    }
    return BC_OK;
}

// Stream-ordered barrier over the ranks of a slab decomposition: the cluster move's phases read and write other
// ranks' memory (the union-find pieces, the perimeter roots of the rank below), so a phase may only start when every
// rank has finished the previous one.  One int all-reduced on the engine stream; no host synchronisation.
static int cluster_barrier(bc_engine* e) {
    if (!(e->ghost && e->nranks > 1)) return BC_OK;
    if (!e->d_cl_barrier) {
        BC_CUDA(e, cudaMalloc(&e->d_cl_barrier, sizeof(int)));
        BC_CUDA(e, cudaMemsetAsync(e->d_cl_barrier, 0, sizeof(int), e->stream));
    }
    BC_NCCL(e, g_nccl.AllReduce(e->d_cl_barrier, e->d_cl_barrier, 1, ncclInt32, ncclSum, e->comm, e->stream));
    return BC_OK;
}

// Peer mapping for the cluster move on slabs: every rank exports 128 bytes (CUDA IPC handles of its piece of the global
// union-find and of its perimeter roots); the caller all-gathers the blobs and hands every rank all of them.
struct ClIpcBlob {
    cudaIpcMemHandle_t gparent, proot;
};
static_assert(sizeof(ClIpcBlob) == 128, "two 64-byte CUDA IPC handles");

extern "C" int bc_slab_cluster_export(bc_engine* e, void* blob128) {
    if (!e || !blob128) return BC_ERR_ARG;
    if (!e->ghost || e->nranks < 2) return fail(e, BC_ERR_STATE, "bc_slab_cluster_export: needs a slab handle with n_ranks >= 2");
    if (e->nranks > CL_MAX_RANKS) return fail(e, BC_ERR_UNSUPPORTED, "bc_slab_cluster_export: at most 8 ranks");
    int rc = cluster_alloc(e);
    if (rc != BC_OK) return rc;
    ClIpcBlob b;
    BC_CUDA(e, cudaIpcGetMemHandle(&b.gparent, e->cl_gparent));
    BC_CUDA(e, cudaIpcGetMemHandle(&b.proot, e->cl_proot));
    std::memcpy(blob128, &b, sizeof(b));
    return BC_OK;
}

extern "C" int bc_slab_cluster_connect(bc_engine* e, const void* blobs /* n_ranks x 128 bytes, in rank order */) {
    if (!e || !blobs) return BC_ERR_ARG;
    if (!e->ghost || e->nranks < 2) return fail(e, BC_ERR_STATE, "bc_slab_cluster_connect: needs a slab handle with n_ranks >= 2");
    if (e->nranks > CL_MAX_RANKS) return fail(e, BC_ERR_UNSUPPORTED, "bc_slab_cluster_connect: at most 8 ranks");
    if (e->cl_connected) return BC_OK;
    BC_CUDA(e, cudaSetDevice(e->device));
    int rc = cluster_alloc(e);
    if (rc != BC_OK) return rc;
    const int below = (e->rank + 1) % e->nranks;
    for (int k = 0; k < e->nranks; ++k) {
        if (k == e->rank) continue;
        ClIpcBlob b;
        std::memcpy(&b, static_cast<const char*>(blobs) + (size_t)k * sizeof(ClIpcBlob), sizeof(b));
        void* p = nullptr;
        BC_CUDA(e, cudaIpcOpenMemHandle(&p, b.gparent, cudaIpcMemLazyEnablePeerAccess));
        e->cl_peer_gparent[k] = static_cast<uint32_t*>(p);
        e->cl_ipc_opened[k] = p;
        if (k == below) {
            void* q = nullptr;
            BC_CUDA(e, cudaIpcOpenMemHandle(&q, b.proot, cudaIpcMemLazyEnablePeerAccess));
            e->cl_peer_proot_below = static_cast<const uint16_t*>(q);
            e->cl_ipc_opened[8] = q;
        }
    }
    e->cl_connected = true;
    return BC_OK;
}

// labels every tile and merges the clusters across tile borders (two launches; between them, and after them, a barrier
// over the ranks of a slab decomposition: the merge reads the rank below's perimeter roots and links into other ranks'
// pieces of the union-find, and what follows walks them)
static int label_clusters(bc_engine* e, const ClArgs& A) {
    cl_label_kernel<<<A.n_tiles, CL_LANES, 0, e->stream>>>(A);
    int rc = cluster_barrier(e);
    if (rc != BC_OK) return rc;
    cl_border_kernel<<<A.n_tiles, 128, 0, e->stream>>>(A);
    e->launches += 2;
    BC_CUDA(e, cudaGetLastError());
    return cluster_barrier(e);
}

extern "C" int bc_cluster_update(bc_engine* e, int64_t n_updates) {
    if (!e || n_updates < 0) return BC_ERR_ARG;
    for (int64_t u = 0; u < n_updates; ++u) {
        ClArgs A;
        int rc = cluster_args(e, "bc_cluster_update", true, e->cluster_step, &A);
        if (rc == BC_OK) rc = label_clusters(e, A);
        if (rc != BC_OK) return rc;
        cl_flip_kernel<<<A.n_tiles, CL_LANES, 0, e->stream>>>(A);
        ++e->launches;
        BC_CUDA(e, cudaGetLastError());
        if (e->ghost) { rc = exchange_both(e); if (rc != BC_OK) return rc; }  // refresh the ghost rows
        rc = cluster_barrier(e);  // the next update re-initialises union-find entries other ranks may still be walking
        if (rc != BC_OK) return rc;
        ++e->cluster_step;
        ++e->state_version;
    }
    return BC_OK;
}

// Cluster labels of the current configurations into e->cl_labels ([replica][L*L], smallest site of the cluster,
// 0xFFFFFFFF for a zero spin): kind 0 = geometric "spin" clusters (every bond between equal non-zero neighbours, what
// export_clusters writes to spin/ with bond probability 1, main.cpp:369-370), kind 1 = FK clusters (bond probability
// 1 - exp(-2J/T), main.cpp:371-372; random bonds from the measurement counter 2^62 + mstep, which never collides
// with the cluster move's counters).
static int cluster_labels(bc_engine* e, const char* who, int kind, int64_t mstep) {
    const size_t N = (size_t)e->L * (size_t)e->L;
    if (e->ghost && e->nranks > 1)
        return fail(e, BC_ERR_UNSUPPORTED, std::string(who) + ": cluster measurements are not available on a slab of a multi-rank decomposition");
    if (N > 0xFFFFFFFFull) return fail(e, BC_ERR_UNSUPPORTED, std::string(who) + ": site labels are 32-bit (L <= 32768)");
    ClArgs A;
    int rc = cluster_args(e, who, kind == 1, ((int64_t)1 << 62) + mstep, &A);
    if (rc != BC_OK) return rc;
    if (!e->cl_labels) BC_CUDA(e, cudaMalloc(&e->cl_labels, (size_t)e->n_rep * N * sizeof(uint32_t)));
    A.labels = e->cl_labels;
    rc = label_clusters(e, A);
    if (rc != BC_OK) return rc;
    cl_labels_kernel<<<A.n_tiles, CL_LANES, 0, e->stream>>>(A);
    ++e->launches;
    BC_CUDA(e, cudaGetLastError());
    return BC_OK;
}

// Cluster-size moments of the current configurations (consumer gamma_nu.cpp:35-60)
extern "C" int bc_cluster_moments(bc_engine* e, int kind, int64_t mstep, bc_cluster_stats* out) {
    if (!e || !out || (kind != 0 && kind != 1) || mstep < 0) return BC_ERR_ARG;
    int rc = cluster_labels(e, "bc_cluster_moments", kind, mstep);
    if (rc != BC_OK) return rc;
    const long long N = (long long)e->L * e->L, total = N * e->n_rep;
    if (!e->d_csize) BC_CUDA(e, cudaMalloc(&e->d_csize, (size_t)total * sizeof(uint32_t)));
    if (!e->d_cmom) BC_CUDA(e, cudaMalloc(&e->d_cmom, (size_t)e->n_rep * 3 * sizeof(unsigned long long)));
    BC_CUDA(e, cudaMemsetAsync(e->d_csize, 0, (size_t)total * sizeof(uint32_t), e->stream));
    BC_CUDA(e, cudaMemsetAsync(e->d_cmom, 0, (size_t)e->n_rep * 3 * sizeof(unsigned long long), e->stream));
    const dim3 cgrid((unsigned)((N + 255) / 256), (unsigned)std::min(e->n_rep, 65535));
    cl_count_kernel<<<cgrid, 256, 0, e->stream>>>(e->cl_labels, e->d_csize, (uint32_t)N, e->n_rep);
    cl_moments_kernel<<<(unsigned)((total + CL_MOM_PER_CTA - 1) / CL_MOM_PER_CTA), 256, 0, e->stream>>>(e->d_csize, N, e->n_rep, e->d_cmom);
    e->launches += 2;
    BC_CUDA(e, cudaGetLastError());
    std::vector<unsigned long long> h((size_t)e->n_rep * 3);
    BC_CUDA(e, cudaMemcpyAsync(h.data(), e->d_cmom, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    for (int r = 0; r < e->n_rep; ++r) {
        out[r].sum_size_sq = (int64_t)h[(size_t)r * 3]; out[r].max_size = (int64_t)h[(size_t)r * 3 + 1];
        out[r].n_clusters = (int64_t)h[(size_t)r * 3 + 2];
    }
    return BC_OK;
}

// ---------------------------------------------------------------------------------------------
// cluster member lists: the occupied sites of every replica sorted by (cluster, site)
// ---------------------------------------------------------------------------------------------
// Labels -> stable LSD radix sort by label (bc_cluster_sort.cuh).  Leaves e->cl_list (pairs (label, site), [replica][N])
// and e->cl_count ([replica] occupied sites).  Cached per (lattice state, kind, mstep): bc_gap_histogram and
// bc_cluster_lists of the same sample share one labelling and one sort.
static int cluster_sorted_lists(bc_engine* e, const char* who, int kind, int64_t mstep) {
    if (e->cl_list && e->list_version == e->state_version && e->list_kind == kind && e->list_mstep == mstep) return BC_OK;
    int rc = cluster_labels(e, who, kind, mstep);
    if (rc != BC_OK) return rc;
    const size_t N = (size_t)e->L * (size_t)e->L;
    const uint32_t n_chunks = (uint32_t)((N + SORT_CHUNK - 1) / SORT_CHUNK);
    if (!e->cl_sort[0]) {
        BC_CUDA(e, cudaMalloc(&e->cl_sort[0], (size_t)e->n_rep * N * sizeof(uint2)));
        BC_CUDA(e, cudaMalloc(&e->cl_sort[1], (size_t)e->n_rep * N * sizeof(uint2)));
        BC_CUDA(e, cudaMalloc(&e->cl_sort_hist, (size_t)e->n_rep * 256 * n_chunks * sizeof(uint32_t)));
        BC_CUDA(e, cudaMalloc(&e->cl_count, (size_t)e->n_rep * sizeof(uint32_t)));
    }
    int bits = 1;
    while (bits < 32 && (1ull << bits) < N) ++bits;  // labels are site indices < N
    const int passes = (bits + 7) / 8;
    const dim3 grid((n_chunks + SORT_WARPS - 1) / SORT_WARPS, (unsigned)e->n_rep);
    for (int p = 0; p < passes; ++p) {
        SortArgs A;
        memset(&A, 0, sizeof(A));
        A.labels = p == 0 ? e->cl_labels : nullptr;
        A.in = p == 0 ? nullptr : e->cl_sort[(p - 1) & 1];
        A.out = e->cl_sort[p & 1];
        A.hist = e->cl_sort_hist; A.count_in = e->cl_count;
        A.n = (uint32_t)N; A.cap = (uint32_t)N; A.n_chunks = n_chunks; A.shift = 8 * p;
        sort_hist_kernel<<<grid, 32 * SORT_WARPS, 0, e->stream>>>(A);
        scan_kernel<<<(unsigned)e->n_rep, 1024, 0, e->stream>>>(e->cl_sort_hist, 256u * n_chunks, (size_t)256 * n_chunks, e->cl_count);
        sort_scatter_kernel<<<grid, 32 * SORT_WARPS, 0, e->stream>>>(A);
        e->launches += 3;
    }
    BC_CUDA(e, cudaGetLastError());
    e->cl_list = e->cl_sort[(passes - 1) & 1];
    e->list_version = e->state_version; e->list_kind = kind; e->list_mstep = mstep;
    return BC_OK;
}

// Gap-size histogram of the current configurations (gap_size_statistics.cpp:58-81,100-144 on the cluster
// file of every sample): ADDS one sample per replica to hist[replica][bins] (variant 0: L/2 bins; 1: L-1 bins).
extern "C" int bc_gap_histogram(bc_engine* e, int kind, int64_t mstep, int variant, int64_t* hist) {
    if (!e || !hist || (kind != 0 && kind != 1) || (variant != 0 && variant != 1) || mstep < 0) return BC_ERR_ARG;
    int rc = cluster_sorted_lists(e, "bc_gap_histogram", kind, mstep);
    if (rc != BC_OK) return rc;
    const size_t bins = variant == 0 ? (size_t)e->L / 2 : (size_t)e->L - 1, total = bins * (size_t)e->n_rep;
    if (total == 0) return BC_OK;
    if (e->gap_hist_cap < total) {
        BC_CUDA(e, cudaStreamSynchronize(e->stream));
        cudaFree(e->d_gap_hist); e->d_gap_hist = nullptr; e->gap_hist_cap = 0;
        BC_CUDA(e, cudaMalloc(&e->d_gap_hist, total * sizeof(unsigned long long)));
        e->gap_hist_cap = total;
    }
    BC_CUDA(e, cudaMemsetAsync(e->d_gap_hist, 0, total * sizeof(unsigned long long), e->stream));
    const size_t N = (size_t)e->L * (size_t)e->L;
    const dim3 grid((unsigned)std::min<size_t>((N + 255) / 256, 148 * 8), (unsigned)e->n_rep);
    gap_hist_kernel<<<grid, 256, 0, e->stream>>>(e->cl_list, e->cl_count, (uint32_t)N, e->L, variant, e->d_gap_hist);
    ++e->launches;
    BC_CUDA(e, cudaGetLastError());
    std::vector<unsigned long long> h(total);
    BC_CUDA(e, cudaMemcpyAsync(h.data(), e->d_gap_hist, total * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < total; ++i) hist[i] += (int64_t)h[i];
    return BC_OK;
}

// Cluster lists of one replica for the dump writer (export_clusters, main.cpp:232-257): clusters in the order of
// their smallest site, members ascending.  sites: the members of all clusters back to back (capacity L*L);
// starts[c]: where cluster c begins in sites (capacity L*L + 1, starts[n_clusters] = n_sites); signs[c] = +-1.
extern "C" int bc_cluster_lists(bc_engine* e, int replica, int kind, int64_t mstep, uint32_t* sites, uint32_t* starts,
                                int8_t* signs, int64_t* n_sites, int64_t* n_clusters) {
    if (!e || !sites || !starts || !signs || !n_sites || !n_clusters || (kind != 0 && kind != 1) || mstep < 0) return BC_ERR_ARG;
    if (replica < 0 || replica >= e->n_rep) return fail(e, BC_ERR_ARG, "bc_cluster_lists: replica out of range");
    const bool cached = e->cl_list && e->list_version == e->state_version && e->list_kind == kind && e->list_mstep == mstep &&
                        e->heads_version == e->state_version && e->heads_kind == kind && e->heads_mstep == mstep;
    const size_t N = (size_t)e->L * (size_t)e->L;
    if (!cached) {
        int rc = cluster_sorted_lists(e, "bc_cluster_lists", kind, mstep);
        if (rc != BC_OK) return rc;
        if (!e->cl_sites) {
            BC_CUDA(e, cudaMalloc(&e->cl_sites, (size_t)e->n_rep * N * sizeof(uint32_t)));
            BC_CUDA(e, cudaMalloc(&e->cl_starts, (size_t)e->n_rep * N * sizeof(uint32_t)));
            BC_CUDA(e, cudaMalloc(&e->cl_ordinal, (size_t)e->n_rep * N * sizeof(uint32_t)));
            BC_CUDA(e, cudaMalloc(&e->cl_signs, (size_t)e->n_rep * N));
            BC_CUDA(e, cudaMalloc(&e->cl_nclusters, (size_t)e->n_rep * sizeof(uint32_t)));
        }
        const dim3 grid((unsigned)std::min<size_t>((N + 255) / 256, 148 * 8), (unsigned)e->n_rep);
        // the ordinals of elements beyond a replica's count are never read: zero them so that the scan totals the heads
        BC_CUDA(e, cudaMemsetAsync(e->cl_ordinal, 0, (size_t)e->n_rep * N * sizeof(uint32_t), e->stream));
        cluster_heads_kernel<<<grid, 256, 0, e->stream>>>(e->cl_list, e->cl_count, (uint32_t)N, e->cl_ordinal);
        scan_kernel<<<(unsigned)e->n_rep, 1024, 0, e->stream>>>(e->cl_ordinal, (uint32_t)N, N, e->cl_nclusters);
        cluster_lists_kernel<<<grid, 256, 0, e->stream>>>(e->cl_list, e->cl_count, (uint32_t)N, e->cl_ordinal, e->d_c[0], e->d_c[1],
                                                         e->L, e->Wc, e->rows_arr, e->row_first, e->cl_sites, e->cl_starts, e->cl_signs);
        e->launches += 3;
        BC_CUDA(e, cudaGetLastError());
        e->heads_version = e->state_version; e->heads_kind = kind; e->heads_mstep = mstep;
    }
    uint32_t cnt = 0, ncl = 0;
    BC_CUDA(e, cudaMemcpyAsync(&cnt, e->cl_count + replica, sizeof(cnt), cudaMemcpyDeviceToHost, e->stream));
    BC_CUDA(e, cudaMemcpyAsync(&ncl, e->cl_nclusters + replica, sizeof(ncl), cudaMemcpyDeviceToHost, e->stream));
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    if (cnt) BC_CUDA(e, cudaMemcpyAsync(sites, e->cl_sites + (size_t)replica * N, (size_t)cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    if (ncl) {
        BC_CUDA(e, cudaMemcpyAsync(starts, e->cl_starts + (size_t)replica * N, (size_t)ncl * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
        BC_CUDA(e, cudaMemcpyAsync(signs, e->cl_signs + (size_t)replica * N, (size_t)ncl, cudaMemcpyDeviceToHost, e->stream));
    }
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    starts[ncl] = cnt;
    *n_sites = cnt; *n_clusters = ncl;
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
    { int rcf = check_device_flags(e, "bc_sync"); if (rcf != BC_OK) return rcf; }
    return BC_OK;
}
