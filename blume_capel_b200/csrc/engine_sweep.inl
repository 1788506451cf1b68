// engine_sweep.inl -- the hot path: launch plans, half-sweep launches, CUDA-graph replay, wavefront / persistent / resident paths, bc_sweep.
// Part of bc_engine.cu (included there once; not a stand-alone translation unit).

// ---------------------------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------------------------

template <int COL, bool SMALL, bool MEAS>
static fast_fn fast_open(bool open) {
    return open ? (fast_fn)sweep_fast_kernel<COL, SMALL, MEAS, true> : (fast_fn)sweep_fast_kernel<COL, SMALL, MEAS, false>;
}

template <int COL, bool MEAS>
static fast_fn fast_halo(bool open, int pair) {
    if (pair) return open ? (fast_fn)sweep_fast_kernel<COL, false, MEAS, true, true, 1> : (fast_fn)sweep_fast_kernel<COL, false, MEAS, false, true, 1>;
    return open ? (fast_fn)sweep_fast_kernel<COL, false, MEAS, true, true> : (fast_fn)sweep_fast_kernel<COL, false, MEAS, false, true>;
}

template <int COL, bool MEAS>
static fast_fn fast_pair(bool open) {
    return open ? (fast_fn)sweep_fast_kernel<COL, false, MEAS, true, false, 1>
                : (fast_fn)sweep_fast_kernel<COL, false, MEAS, false, false, 1>;
}

static void pair_kernels(fast_fn (&out)[16]) {
    int n = 0;
    for (int open = 0; open < 2; ++open) {
        out[n++] = fast_pair<0, false>(open); out[n++] = fast_pair<0, true>(open);
        out[n++] = fast_pair<1, false>(open); out[n++] = fast_pair<1, true>(open);
        out[n++] = fast_halo<0, false>(open, 1); out[n++] = fast_halo<0, true>(open, 1);
        out[n++] = fast_halo<1, false>(open, 1); out[n++] = fast_halo<1, true>(open, 1);
    }
}

static fast_fn pick_fast(int col, bool small, bool meas, bool open, bool halo = false, int pair = 0) {
    if (halo && !small) {
        if (col == 0) return meas ? fast_halo<0, true>(open, pair) : fast_halo<0, false>(open, pair);
        return meas ? fast_halo<1, true>(open, pair) : fast_halo<1, false>(open, pair);
    }
    if (pair && !small) {
        if (col == 0) return meas ? fast_pair<0, true>(open) : fast_pair<0, false>(open);
        return meas ? fast_pair<1, true>(open) : fast_pair<1, false>(open);
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
    int pair = 0;  // 1: pair-table kernels (one replica per CTA, enough rows per thread to amortise the 11.4 KB table)
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
            // the tallest strip that still leaves enough threads to fill the chip (7 CTAs of 128 per SM); with the
            // branch-free row loop 32 rows beat 64 on the bench batch (1938 against 1919 updates/ns: twice the CTAs,
            // a fuller last wave) and 16 (1904: the per-strip prologue -- pair table, first rows -- weighs more)
            const long long want = 148LL * 768;
            const int cand[5] = {32, 16, 8, 4, 2};
            for (int i = 0; i < 5 && !R; ++i)
                if (rows_valid(Ly, p.cpr, cand[i]) && (long long)e->n_rep * (Ly / cand[i]) * p.cpr >= want) R = cand[i];
            if (!R) R = rows_valid(Ly, p.cpr, 2) ? 2 : 1;
        }
        p.rows = R;
        p.strips = Ly / R;
        p.small = (R & 1) != 0 || (((long long)p.strips * p.cpr) % FAST_THREADS) != 0;
        if (!p.small && e->d_tab2 && e->opt_pair_min_rows > 0 && R >= e->opt_pair_min_rows)
            p.pair = 1;
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
    P.halo_flags = nullptr; P.halo_clock = nullptr; P.halo_done = nullptr; P.halo_err = nullptr;
    P.peer_prev_flags = nullptr; P.peer_next_flags = nullptr;
    P.halo_timeout_ns = 0; P.halo_ctas_per_rep = 0; P.halo_need = 0;
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

// Peer-memory slab whose whole half-sweep is ONE launch of the HALO instantiation (boundary strips first, waits,
// halo stores and epoch signal inside the kernel): graph replay and programmatic dependent launch apply to it.
static bool slab_fused(const bc_engine* e, const LaunchPlan& plan) {
    return e->ghost && e->p2p && e->opt_fused_halo && plan.fast && !plan.small;
}

// part: 0 = all strips, 1 = the two boundary strips of the slab, 2 = the interior strips
static int launch_half(bc_engine* e, const LaunchPlan& plan, int col, bool measure, int64_t t, int64_t slot,
                       const int64_t* t_ptr = nullptr, int part = 0, cudaStream_t stream = nullptr) {
    if (!stream) stream = e->stream;
    SweepParams P = sweep_params(e, plan, col, t, slot, t_ptr);
    const bool fused_halo = part == 0 && slab_fused(e, plan);
    if (fused_halo) {
        const bool periodic = e->boundary == BC_PERIODIC;
        const bool has_prev = e->rank > 0 || periodic, has_next = e->rank + 1 < e->nranks || periodic;
        if (has_prev) { P.peer_prev = e->peer_prev_c[col]; P.peer_prev_flags = e->peer_prev_flags; }
        if (has_next) { P.peer_next = e->peer_next_c[col]; P.peer_next_flags = e->peer_next_flags; }
        P.halo_flags = e->d_flags; P.halo_clock = e->d_halo_clock; P.halo_done = e->d_halo_done; P.halo_err = e->d_halo_err;
        P.halo_timeout_ns = (long long)e->opt_halo_timeout_ms * 1000000LL;
        // CTAs per replica that hold (part of) the first or the last strip: see halo_cta_order
        const long long C = (long long)plan.strips * plan.cpr / FAST_THREADS;
        const long long nb = plan.cpr >= FAST_THREADS ? plan.cpr / FAST_THREADS : 1;
        P.halo_ctas_per_rep = (int)std::min<long long>(2 * nb, C);
        P.halo_need = (has_prev ? 1 : 0) | (has_next ? 2 : 0);
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
    { int rcf = check_device_flags(e, "observables"); if (rcf != BC_OK) return rcf; }
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
    if (slab_fused(e, plan)) return launch_half(e, plan, col, measure, t, slot);
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
    rc = exchange_halo(e, col, e->halo_stream);
    if (rc != BC_OK) return rc;
    BC_CUDA(e, cudaEventRecord(e->ev_halo, e->halo_stream));
    launch_half(e, plan, col, measure, t, slot, nullptr, 2, e->stream);
    BC_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_halo, 0));
    return BC_OK;
}

// (Re)capture the graph of `block` unmeasured sweeps for the current plan: 2*block kernels reading the
// device clock plus one node that advances it.
static int ensure_graph(bc_engine* e, const LaunchPlan& plan, int block) {
    const int halo_sig = slab_fused(e, plan) ? 1 : 0;
    if (e->graph_exec && e->graph_rows == plan.rows && e->graph_grid == plan.grid && e->graph_block == block &&
        e->graph_halo == halo_sig) return BC_OK;
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
    e->graph_rows = plan.rows; e->graph_grid = plan.grid; e->graph_block = block; e->graph_halo = halo_sig;
    return BC_OK;
}

#ifdef BC_EXPERIMENTAL
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

#endif  // BC_EXPERIMENTAL

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

#ifdef BC_EXPERIMENTAL
// ---- persistent kernel: a whole run of unmeasured sweeps in one cooperative launch -------------------------
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

#endif  // BC_EXPERIMENTAL

// n unmeasured sweeps starting at sweep index t0
static int run_plain_sweeps(bc_engine* e, const LaunchPlan& plan, int64_t t0, int64_t n) {
#ifdef BC_EXPERIMENTAL
    if (wave_eligible(e, plan, n)) return run_wave(e, plan, t0, n);
    if (plan.fast) {
        GroupPlan gp;
        bc_engine* one[1] = {e};
        if (persist_plan(one, 1, n, &gp)) return run_persist(one, gp, t0 - e->t, n);
    }
#endif
    int64_t done = 0;
    const int block = e->opt_graph_block;
    if (e->opt_use_graph && (!e->ghost || slab_fused(e, plan)) && block > 0 && n >= block) {
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
    bool pair = false;  // pair table staged behind the lattice (one replica per CTA; launch_resident: calls of >= 32 sweeps)
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
    r.pair = e->d_tab2 != nullptr && r.reps_per_cta == 1 && e->opt_pair_min_rows > 0 && r.smem + TAB2_WORDS * 4 <= 227 * 1024;
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
    const bool meas = measure_every > 0;
    const bool pair = rp.pair && n >= 32;  // staging 11.4 KB pays for itself after a few dozen sweeps
    typedef void (*res_fn)(const ResidentParams);
    // [pair][meas][any][open]
    static const res_fn fns[2][2][2][2] = {
        {{{sweep_resident_kernel<false, false, false, false>, sweep_resident_kernel<true, false, false, false>},
          {sweep_resident_kernel<false, true, false, false>, sweep_resident_kernel<true, true, false, false>}},
         {{sweep_resident_kernel<false, false, true, false>, sweep_resident_kernel<true, false, true, false>},
          {sweep_resident_kernel<false, true, true, false>, sweep_resident_kernel<true, true, true, false>}}},
        {{{sweep_resident_kernel<false, false, false, true>, sweep_resident_kernel<true, false, false, true>},
          {sweep_resident_kernel<false, true, false, true>, sweep_resident_kernel<true, true, false, true>}},
         {{sweep_resident_kernel<false, false, true, true>, sweep_resident_kernel<true, false, true, true>},
          {sweep_resident_kernel<false, true, true, true>, sweep_resident_kernel<true, true, true, true>}}}};
    const res_fn fn = fns[pair][meas][any][open];
    const int fi = 8 * pair + 4 * meas + 2 * any + open;
    if (!e->resident_attr_set[fi]) {
        BC_CUDA(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        e->resident_attr_set[fi] = true;
    }
    ResidentParams P;
    P.c0 = e->d_c[0]; P.c1 = e->d_c[1];
    P.tab16 = e->d_tab16; P.tab31 = e->d_tab31; P.tab2 = e->d_tab2;
    P.obs = e->d_obs; P.ties = e->d_ties;
    P.t0 = t0; P.n_sweeps = n; P.measure_every = measure_every;
    P.L = e->L; P.Wc = e->Wc; P.cpr = e->Wc / 16; P.n_replicas = e->n_rep;
    P.threads_per_rep = rp.T; P.reps_per_cta = rp.reps_per_cta;
    P.key0 = (uint32_t)e->seed; P.key1 = (uint32_t)(e->seed >> 32);
    fn<<<rp.grid, rp.T * rp.reps_per_cta, rp.smem + (pair ? TAB2_WORDS * 4 : 0), e->stream>>>(P);
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
    if (n_sweeps > 0) ++e->state_version;
    e->pending_meas = n_meas;
    if (out && n_meas > 0) {
        int rc = fetch_pending(e, out);
        if (rc != BC_OK) return rc;
        e->pending_meas = 0;  // fetched (and, for slabs, already all-reduced in place)
        return n_meas * e->n_rep;
    }
    return 0;
}
