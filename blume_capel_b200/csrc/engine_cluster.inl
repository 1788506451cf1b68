// engine_cluster.inl -- cluster move and cluster-size moments, bc_sync.
// Part of bc_engine.cu (included there once; not a stand-alone translation unit).

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
    { int rcf = check_device_flags(e, "bc_sync"); if (rcf != BC_OK) return rcf; }
    return BC_OK;
}
