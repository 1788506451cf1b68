// engine_halo.inl -- slab halo exchange over NCCL or peer memory (between half-sweeps of a slab handle).
// Part of bc_engine.cu (included there once; not a stand-alone translation unit).

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
        halo_signal_kernel<<<1, 1, 0, stream>>>(e->d_halo_clock, prev >= 0 ? e->peer_prev_flags : nullptr,
                                                next >= 0 ? e->peer_next_flags : nullptr);
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
    if (!e->p2p) return BC_OK;
    const bool periodic = e->boundary == BC_PERIODIC;
    const int need_top = (e->rank > 0 || periodic) ? 1 : 0, need_bottom = (e->rank + 1 < e->nranks || periodic) ? 1 : 0;
    halo_wait_kernel<<<1, 1, 0, stream>>>(e->d_flags, e->d_halo_clock, need_top, need_bottom,
                                          (long long)e->opt_halo_timeout_ms * 1000000LL, e->d_halo_err);
    ++e->launches;
    BC_CUDA(e, cudaGetLastError());
    return BC_OK;
}

// Sticky device-side failure flags (a slab neighbour that did not deliver its halo rows in time; the experimental
// kernels' bounded waits): read after the stream has been synchronised, by every synchronising entry point.
static int check_device_flags(bc_engine* e, const char* who) {
    if (e->d_halo_err) {
        int err = 0;
        BC_CUDA(e, cudaMemcpy(&err, e->d_halo_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (err) return fail(e, BC_ERR_STATE, std::string(who) + ": a slab neighbour did not deliver its halo rows in time "
                                              "(results since then are invalid; option \"halo_timeout_ms\")");
    }
#ifdef BC_EXPERIMENTAL
    if (e->d_wave) {
        int flags[2] = {0, 0};
        BC_CUDA(e, cudaMemcpy(flags, e->d_wave, sizeof(flags), cudaMemcpyDeviceToHost));
        if (flags[1]) return fail(e, BC_ERR_STATE, std::string(who) + ": the wavefront kernel timed out waiting for a dependency");
    }
    if (e->d_bar) {
        unsigned bar[3] = {0, 0, 0};
        BC_CUDA(e, cudaMemcpy(bar, e->d_bar, sizeof(bar), cudaMemcpyDeviceToHost));
        if (bar[2]) return fail(e, BC_ERR_STATE, std::string(who) + ": the persistent kernel timed out at a grid barrier");
    }
#endif
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
