// engine_introspection.inl -- introspection, options, plan description.
// Part of bc_engine.cu (included there once; not a stand-alone translation unit).

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
    if (n == "pair_min_rows") { e->opt_pair_min_rows = (int)value; if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; } return BC_OK; }
#ifdef BC_EXPERIMENTAL
    if (n == "wave") { e->opt_wave = (int)value; return BC_OK; }
    if (n == "persist") { e->opt_persist = (int)value; return BC_OK; }
#else
    if (n == "wave" || n == "persist") {
        if (value == 0) return BC_OK;
        return fail(e, BC_ERR_UNSUPPORTED, "bc_set_option: " + n + " is an experimental kernel (a measured negative result); build with -DBC_EXPERIMENTAL");
    }
#endif
    if (n == "halo_timeout_ms") { if (value < 1 || value > 3600000) return fail(e, BC_ERR_ARG, "halo_timeout_ms out of range"); e->opt_halo_timeout_ms = (int)value; return BC_OK; }
    if (n == "pdl") { e->opt_pdl = value != 0; if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; } return BC_OK; }
    if (n == "fused_halo") { e->opt_fused_halo = value != 0; if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; } return BC_OK; }
    if (n == "overlap") { e->opt_overlap = value != 0; return BC_OK; }
    if (n == "resident") { e->opt_resident = (int)value; return BC_OK; }
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
    if (fast) *fast = p.fast ? (p.pair ? 3 : 1) : 0;  // 3 = streaming kernel with the pair table
    if (rows_per_thread) *rows_per_thread = p.rows;
    if (grid) *grid = (int)p.grid;
    if (block) *block = p.fast ? FAST_THREADS : 256;
    return BC_OK;
}
