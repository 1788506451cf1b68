// engine_state.inl -- lattice state: initialisation, uploads / downloads in the reference layout, sweep counter.
// Part of bc_engine.cu (included there once; not a stand-alone translation unit).

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
    ++e->state_version;
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
    ++e->state_version;
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
    if (wait) {
        BC_CUDA(e, cudaStreamSynchronize(e->stream));
        return check_device_flags(e, "bc_download");
    }
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

// ---- 2-bit wire format (bc_util_kernels.cuh): four sites per byte between host and device ---------------------
static size_t packed_bytes(const bc_engine* e, int n) { return (size_t)n * e->Ly * (size_t)((e->L + 3) / 4); }

static int packed_transfer(bc_engine* e, int r0, int n, uint8_t* host, bool upload, bool wait) {
    BC_CUDA(e, cudaSetDevice(e->device));
    const size_t bytes = packed_bytes(e, n);
    int rc = ensure_stage(e, bytes);
    if (rc != BC_OK) return rc;
    uint8_t* stage = reinterpret_cast<uint8_t*>(e->d_stage);
    const bool vec = e->L % 64 == 0 && e->Wc == e->L / 2 && n <= 65535;
    const dim3 vgrid((unsigned)(((long long)e->Ly * (e->L / 64) + 255) / 256), (unsigned)n);
    if (upload) {
        ++e->state_version;
        BC_CUDA(e, cudaMemcpyAsync(stage, host, bytes, cudaMemcpyHostToDevice, e->stream));
        if (vec) unpack2_kernel<<<vgrid, 256, 0, e->stream>>>(stage, e->d_c[0], e->d_c[1], geom_of(e), r0);
        else unpack2_any_kernel<<<(unsigned)(((long long)n * e->Ly * e->L + 255) / 256), 256, 0, e->stream>>>(stage, e->d_c[0], e->d_c[1], geom_of(e), r0, n);
        ++e->launches;
        BC_CUDA(e, cudaGetLastError());
        rc = exchange_both(e);
        if (rc != BC_OK) return rc;
    } else {
        if (vec) pack2_kernel<<<vgrid, 256, 0, e->stream>>>(stage, e->d_c[0], e->d_c[1], geom_of(e), r0);
        else pack2_any_kernel<<<(unsigned)((bytes + 255) / 256), 256, 0, e->stream>>>(stage, e->d_c[0], e->d_c[1], geom_of(e), r0, n);
        ++e->launches;
        BC_CUDA(e, cudaGetLastError());
        BC_CUDA(e, cudaMemcpyAsync(host, stage, bytes, cudaMemcpyDeviceToHost, e->stream));
    }
    if (wait) {
        BC_CUDA(e, cudaStreamSynchronize(e->stream));
        return check_device_flags(e, upload ? "bc_upload_packed" : "bc_download_packed");
    }
    return BC_OK;
}

extern "C" int64_t bc_packed_size(const bc_engine* e, int n_lattices) {
    if (!e || n_lattices < 0) return BC_ERR_ARG;
    return (int64_t)packed_bytes(e, n_lattices);
}
extern "C" int bc_upload_packed(bc_engine* e, int replica, const uint8_t* packed, int async) {
    if (!e || !packed) return BC_ERR_ARG;
    if (replica < -1 || replica >= e->n_rep) return fail(e, BC_ERR_ARG, "bc_upload_packed: replica out of range");
    return packed_transfer(e, replica < 0 ? 0 : replica, replica < 0 ? e->n_rep : 1, const_cast<uint8_t*>(packed), true, !async);
}
extern "C" int bc_download_packed(bc_engine* e, int replica, uint8_t* packed, int async) {
    if (!e || !packed) return BC_ERR_ARG;
    if (replica < -1 || replica >= e->n_rep) return fail(e, BC_ERR_ARG, "bc_download_packed: replica out of range");
    return packed_transfer(e, replica < 0 ? 0 : replica, replica < 0 ? e->n_rep : 1, packed, false, !async);
}
// host-side conversion between the reference layout (int8 spins) and the wire format, row by row
extern "C" int bc_pack_spins(const int8_t* spins, int L, int64_t rows, uint8_t* packed) {
    if (!spins || !packed || L < 1 || rows < 0) return BC_ERR_ARG;
    const int row_bytes = (L + 3) / 4;
    for (int64_t y = 0; y < rows; ++y) {
        const int8_t* s = spins + y * L;
        uint8_t* p = packed + y * row_bytes;
        for (int b = 0; b < row_bytes; ++b) {
            unsigned v = 0;
            for (int i = 0; i < 4; ++i) v |= (unsigned)((4 * b + i < L ? s[4 * b + i] : 0) + 1) << (2 * i);
            p[b] = (uint8_t)v;
        }
    }
    return BC_OK;
}
extern "C" int bc_unpack_spins(const uint8_t* packed, int L, int64_t rows, int8_t* spins) {
    if (!spins || !packed || L < 1 || rows < 0) return BC_ERR_ARG;
    const int row_bytes = (L + 3) / 4;
    for (int64_t y = 0; y < rows; ++y)
        for (int x = 0; x < L; ++x) spins[y * L + x] = (int8_t)((int)((packed[y * row_bytes + (x >> 2)] >> (2 * (x & 3))) & 3u) - 1);
    return BC_OK;
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
