// engine_group.inl -- bc_sweep_group (several lattice sizes in one launch per half-sweep), bc_read_obs, bc_measure.
// Part of bc_engine.cu (included there once; not a stand-alone translation unit).

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
        sig.push_back((uint64_t)es[g]->L); sig.push_back((uint64_t)es[g]->boundary); sig.push_back((uint64_t)gp.cta_end[g]);
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
        // open lattices: 8-row strips at most (config 5, L = 128..1024 x 64 replicas: 1431 updates/ns against 1369 with
        // the 16-row strips the thread count alone would allow -- the open-boundary kernels stage one row ahead only
        // and carry the edge handling at the register limit; tools/tune_group.py)
        for (int cand = gp.open ? 8 : 32; cand >= 2; cand >>= 1) {
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

    if (gp.pair) {  // 7 CTAs x 31 KB per SM need the large shared-memory carve-out (a per-device function attribute)
        bool& carveout_set = e0->group_carveout_set;
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
#ifdef BC_EXPERIMENTAL
        GroupPlan pp;
        if (until > done && persist_plan(engines, n_engines, until - done, &pp)) {
            int rc = run_persist(engines, pp, done, until - done);
            if (rc != BC_OK) return rc;
            done = until;
        }
#endif
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
        if (n_sweeps > 0) ++e->state_version;
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
    { int rcf = check_device_flags(e, "bc_measure"); if (rcf != BC_OK) return rcf; }
    for (int r = 0; r < e->n_rep; ++r) raw_to_obs(e, &raw[(size_t)r * OBS_RAW], e->t - 1, &out[r]);
    return BC_OK;
}
