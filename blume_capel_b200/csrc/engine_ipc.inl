// engine_ipc.inl -- peer-memory slab halos: export / connect CUDA IPC handles.
// Part of bc_engine.cu (included there once; not a stand-alone translation unit).

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
    if (!e->ghost) return fail(e, BC_ERR_STATE, "bc_slab_ipc_connect: needs a slab handle");
    BC_CUDA(e, cudaSetDevice(e->device));
    BC_CUDA(e, cudaStreamSynchronize(e->stream));
    if (e->nranks == 1) {
        // A single slab is its own ring neighbour: with two NULL blobs the peer pointers are the handle's own arrays
        // and flags, and the HALO kernels (boundary CTAs first, in-kernel wait / halo stores / epoch signal) run on
        // one GPU exactly as they do across GPUs.  This is how the single-GPU tests cover them.
        if (prev_blob192 || next_blob192) return fail(e, BC_ERR_ARG, "bc_slab_ipc_connect: a one-rank slab connects to itself (pass NULL, NULL)");
        for (int c = 0; c < 2; ++c) { e->peer_prev_c[c] = e->d_c[c]; e->peer_next_c[c] = e->d_c[c]; }
        e->peer_prev_flags = e->d_flags; e->peer_next_flags = e->d_flags;
        e->p2p = true;
        return BC_OK;
    }
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
