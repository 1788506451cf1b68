"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Sizes are chosen so the oracle finishes in seconds; full-size (BASELINE.json) checks use
size-independent properties in test_gpu_fullsize.py.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TC = (0.608, 1.966, 1.0)  # tricritical point of main.cpp:34-36


def set_opt(eng, name, value):
    """set_option; the experimental kernels ("wave", "persist": measured negative results) are only compiled with
    -DBC_EXPERIMENTAL, their tests skip on a default build."""
    try:
        eng.set_option(name, value)
    except Exception as ex:
        if getattr(ex, "code", 0) == -5:
            eng.close()
            pytest.skip(f"{name}: experimental kernel not compiled in (build with -DBC_EXPERIMENTAL)")
        raise


def run_pair(bc, oracle, L, boundary, n_rep, n_sweeps, seed, params=TC, opts=(), measure_every=1, start="random"):
    eng = bc.Engine(L, boundary, n_rep, 0, seed)
    for k, v in opts:
        set_opt(eng, k, v)
    eng.set_params(*params)
    tab = oracle.build_table(*params)
    ref = np.zeros((n_rep, L, L), np.int8)
    if start == "random":
        eng.init_random()
        for r in range(n_rep):
            ref[r] = oracle.init_random(L, seed, r)
    else:
        rng = np.random.default_rng(seed)
        ref[:] = rng.integers(-1, 2, size=ref.shape, dtype=np.int8)
        eng.upload(ref)
    got0 = eng.download()
    assert np.array_equal(got0, ref), "initial lattice differs"
    obs = eng.sweep(n_sweeps, measure_every)
    got = eng.download()
    ref_obs = []
    for r in range(n_rep):
        ref_obs.append(oracle.sweep(ref[r], tab, seed, r, 0, n_sweeps, measure_every, boundary))
    plan = eng.plan()
    plan["ties"] = eng.tie_count
    eng.close()
    return got, ref, obs, ref_obs, plan


@pytest.mark.parametrize("L", [2, 4, 6, 8, 12, 16, 24, 48])
def test_generic_kernel_periodic(bc, oracle, L):
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, L, bc.PERIODIC, 3, 20, 0xBC00 + L, opts=(("resident", 0),))
    assert not plan["fast"]
    assert np.array_equal(got, ref)
    for r in range(3):
        for f in ("sweep", "M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ref_obs[r][f]), (r, f)


@pytest.mark.parametrize("L", [3, 5, 8, 17, 33])
def test_generic_kernel_open(bc, oracle, L):
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, L, bc.OPEN, 2, 15, 0x0BE0 + L, opts=(("resident", 0),))
    assert not plan["fast"]
    assert np.array_equal(got, ref)
    for r in range(2):
        for f in ("M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ref_obs[r][f]), (r, f)


@pytest.mark.parametrize("L,rows", [(32, 1), (32, 0), (64, 0), (64, 2), (64, 4), (96, 0), (96, 3), (128, 0),
                                    (128, 8), (128, 16), (256, 16)])
@pytest.mark.parametrize("boundary", [0, 1])
def test_fast_kernel(bc, oracle, L, rows, boundary):
    n_rep = 3 if L <= 128 else 1
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, L, boundary, n_rep, 12, 0xFA57 + L + rows,
                                            opts=(("rows_per_thread", rows), ("resident", 0)))
    assert plan["fast"] and not plan["resident"]
    if rows:
        assert plan["rows_per_thread"] == rows
    assert np.array_equal(got, ref), f"lattice differs, plan={plan}"
    for r in range(n_rep):
        for f in ("sweep", "M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ref_obs[r][f]), (r, f, plan)


@pytest.mark.parametrize("L,n_rep,rows", [(256, 2, 8), (256, 1, 16), (512, 2, 16), (1024, 1, 32)])
@pytest.mark.parametrize("boundary", [0, 1])
@pytest.mark.parametrize("params", [TC, (0.3, 1.0, 1.0), (5.0, -2.0, -1.0)])
def test_pair_table_kernel(bc, oracle, L, n_rep, rows, boundary, params):
    """Streaming kernel with the 54 x 54 pair table (one threshold lookup per two sites) against the oracle,
    with and without fused observables, at parameter points whose tables hold the extreme entries
    (always accept: E15 = 32767; never accept: E15 = 0)."""
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, L, boundary, n_rep, 6, 0x9A1 + L + rows, params=params,
                                            opts=(("rows_per_thread", rows), ("resident", 0), ("pair_min_rows", 8)),
                                            measure_every=2)
    assert plan["pair"] and plan["rows_per_thread"] == rows
    assert np.array_equal(got, ref), f"lattice differs, plan={plan}"
    for r in range(n_rep):
        for f in ("sweep", "M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ref_obs[r][f]), (r, f, plan)


@pytest.mark.parametrize("params", [(0.608, 1.966, 1.0), (2.2691853, -1000.0, 1.0), (1.0, 0.5, 0.0), (0.3, 1.0, 1.0),
                                    (5.0, -2.0, -1.0)])
def test_parameter_points(bc, oracle, params):
    """Ising limit D=-1000 (main.cpp:27-29), J=0, low T, antiferromagnetic J."""
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, 64, bc.PERIODIC, 2, 10, 77, params=params, start="upload")
    assert np.array_equal(got, ref)
    assert np.array_equal(obs[:, 1]["B"], ref_obs[1]["B"])


def test_measure_every_and_counter(bc, oracle):
    got, ref, obs, ref_obs, _ = run_pair(bc, oracle, 64, 0, 2, 23, 5, measure_every=5)
    assert obs.shape == (4, 2)
    assert list(obs[:, 0]["sweep"]) == [4, 9, 14, 19]
    assert np.array_equal(obs[:, 0]["M"], ref_obs[0]["M"])
    assert np.array_equal(got, ref)


def test_table_equals_oracle(bc, oracle):
    eng = bc.Engine(32, 0, 2, 0, 1)
    for p in [(0.608, 1.966, 1.0), (2.2691853, -1000.0, 1.0), (0.1, 3.0, 1.0), (10.0, 0.0, 0.0)]:
        eng.set_params(*p)
        assert np.array_equal(eng.table(1), oracle.build_table(*p))
    eng.set_params(1.0, 1.0, 1.0, replica=1)
    assert np.array_equal(eng.table(1), oracle.build_table(1.0, 1.0, 1.0))
    assert np.array_equal(eng.table(0), oracle.build_table(10.0, 0.0, 0.0))
    eng.close()


def test_per_replica_parameters(bc, oracle):
    """Replicas of one launch may sit at different (T, D) (finite-size-scaling grids)."""
    L, n_rep, seed = 64, 4, 99
    eng = bc.Engine(L, 0, n_rep, 0, seed)
    pts = [(0.58 + 0.02 * r, 1.94 + 0.01 * r, 1.0) for r in range(n_rep)]
    for r, p in enumerate(pts):
        eng.set_params(*p, replica=r)
    eng.init_random()
    obs = eng.sweep(8, 1)
    got = eng.download()
    for r, p in enumerate(pts):
        ref = oracle.init_random(L, seed, r)
        ro = oracle.sweep(ref, oracle.build_table(*p), seed, r, 0, 8, 1)
        assert np.array_equal(got[r], ref), r
        assert np.array_equal(obs[:, r]["B"], ro["B"])
    eng.close()


def test_standalone_measure(bc, oracle):
    rng = np.random.default_rng(3)
    for L, b in [(10, 0), (64, 0), (7, 1), (64, 1)]:
        s = rng.integers(-1, 2, size=(2, L, L), dtype=np.int8)
        eng = bc.Engine(L, b, 2, 0, 0)
        eng.upload(s)
        m = eng.measure()
        for r in range(2):
            assert (m[r]["M"], m[r]["Q"], m[r]["B"]) == oracle.observables(s[r], b)
        eng.close()


def test_checkpoint_resume(bc):
    """(lattice, seed, sweep counter) is a complete checkpoint."""
    a = bc.Engine(128, 0, 1, 0, 4242)
    a.set_params(*TC)
    a.init_random()
    a.sweep(10)
    full = a.download(0)
    b = bc.Engine(128, 0, 1, 0, 4242)
    b.set_params(*TC)
    b.init_random()
    b.sweep(4)
    half = b.download(0)
    t = b.sweep_counter
    b.close()
    c = bc.Engine(128, 0, 1, 0, 4242)
    c.set_params(*TC)
    c.upload(half, 0)
    c.sweep_counter = t
    c.sweep(6)
    assert np.array_equal(c.download(0), full)
    a.close(); c.close()


def test_ties_are_resolved_like_the_oracle(bc, oracle):
    """Enough attempts that coarse ties (2^-15 per uphill attempt) occur; lattice must still match."""
    L, n = 128, 40
    oracle.tie_count(reset=True)
    eng = bc.Engine(L, 0, 1, 0, 31337)
    eng.set_params(*TC)
    eng.init_random()
    eng.sweep(n, 0)
    got, gpu_ties = eng.download(0), eng.tie_count
    eng.close()
    ref = oracle.init_random(L, 31337, 0)
    oracle.sweep(ref, oracle.build_table(*TC), 31337, 0, 0, n, 0, 0)
    assert np.array_equal(got, ref)
    assert oracle.tie_count() > 0  # the case was exercised
    assert gpu_ties == oracle.tie_count()  # the kernels take the fine-stream path exactly when the oracle does


def test_errors(bc):
    with pytest.raises(bc.BCError):
        bc.Engine(33, bc.PERIODIC)  # odd torus
    with pytest.raises(bc.BCError):
        bc.Engine(32, 7)
    e = bc.Engine(32)
    with pytest.raises(bc.BCError):
        e.sweep(1)  # parameters not set
    with pytest.raises(bc.BCError):
        e.set_params(-1.0, 0.0)
    with pytest.raises(bc.BCError):
        e.set_option("nope", 1)
    e.close()


def test_committed_golden_trajectories(bc):
    """GPU against tests/golden/trajectories.npz (written by the oracle, make_golden.py)."""
    import json, os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "trajectories.npz"))
    meta = json.loads(str(d["meta"]))
    for i, m in enumerate(meta):
        # replica r of a batch == single engine with seed ^ (r << 32)
        e = bc.Engine(m["L"], m["boundary"], m["replica"] + 1, 0, m["seed"])
        e.set_params(m["T"], m["D"], m["J"])
        e.init_random()
        assert np.array_equal(e.download(m["replica"]), d[f"init_{i}"]), i
        obs = e.sweep(m["n_sweeps"], 1)[:, m["replica"]]
        assert np.array_equal(e.download(m["replica"]), d[f"final_{i}"]), i
        assert np.array_equal(np.stack([obs["M"], obs["Q"], obs["B"]]), d[f"obs_{i}"]), i
        e.close()


@pytest.mark.parametrize("boundary", [0, 1])
def test_single_rank_slab_equals_plain_engine(bc, oracle, boundary):
    """The slab code path (ghost rows + halo exchange, here with itself) is bit-identical to the
    wrap-around path and to the oracle; exercises bc_create_slab with n_ranks = 1 on one GPU."""
    L, n, seed = 128, 10, 0x51AB
    e = bc.Engine(L, boundary, 2, 0, seed, slab=(1, 0, None))
    assert (e.rows, e.first_row) == (L, 0)
    e.set_params(*TC)
    e.init_random()
    obs = e.sweep(n, 1)
    got = e.download()
    m = e.measure()
    e.close()
    tab = oracle.build_table(*TC)
    for r in range(2):
        ref = oracle.init_random(L, seed, r)
        ro = oracle.sweep(ref, tab, seed, r, 0, n, 1, boundary)
        assert np.array_equal(got[r], ref)
        for f in ("M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ro[f]), f
        assert (m[r]["M"], m[r]["Q"], m[r]["B"]) == oracle.observables(ref, boundary)


@pytest.mark.parametrize("L,n_rep,rows", [(1024, 2, 32), (2048, 2, 32), (1024, 2, 64), (2048, 1, 64)])
def test_headline_plan_matches_oracle(bc, oracle, L, n_rep, rows):
    """The exact plan the bench line runs (bench.py: 64 x L=4096 -> 32-row strips, pair-table kernel with the
    branch-free row loop and the warp-cooperative tie repair, bc_describe_plan == 3) against the oracle at a size the
    oracle can afford, measured and unmeasured sweeps mixed, graph replay included (36 > graph_block); 64-row
    strips (the round-1 plan) as well.  Thousands of ties per run: the tie counts must agree too."""
    n, every = 36, 9
    oracle.tie_count(reset=True)
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, L, bc.PERIODIC, n_rep, n, 0xBE7C4 + L, measure_every=every,
                                            opts=(("rows_per_thread", rows), ("resident", 0)))
    assert plan["pair"] and plan["rows_per_thread"] == rows, plan
    assert np.array_equal(got, ref), f"lattice differs, plan={plan}"
    for r in range(n_rep):
        for f in ("sweep", "M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ref_obs[r][f]), (r, f, plan)
    assert plan["ties"] == oracle.tie_count() > 1000, (plan, oracle.tie_count())


@pytest.mark.parametrize("boundary,rows,slab", [(0, 8, False), (1, 8, False), (0, 16, False), (1, 16, False), (0, 8, True), (1, 16, True)])
def test_tie_repair_at_lattice_edges(bc, oracle, boundary, rows, slab):
    """The out-of-line tie repair of the pair-table kernels where its addressing is special: a small lattice for many
    sweeps, so that ties (about 1400 per run) fall on the first and last rows (periodic wrap / open edge rows), on the
    first and last chunk of a row (neighbour word from the other end of the row / the open edge) and, on a one-rank
    slab, on the rows that travel to the neighbour's ghost rows after the repair.  Lattices, fused observables and tie
    counts against the oracle."""
    L, n_rep, n, every, seed = 256, 2, 400, 50, 0x71E + rows + boundary
    oracle.tie_count(reset=True)
    e = bc.Engine(L, boundary, n_rep, 0, seed, slab=(1, 0, None)) if slab else bc.Engine(L, boundary, n_rep, 0, seed)
    if slab:
        e.ipc_connect(None, None)
    e.set_option("rows_per_thread", rows)
    e.set_option("resident", 0)
    e.set_params(*TC)
    e.init_random()
    plan = e.plan()
    assert plan["pair"] and plan["rows_per_thread"] == rows, plan
    obs = e.sweep(n, every)
    got, ties = e.download(), e.tie_count
    e.close()
    tab = oracle.build_table(*TC)
    for r in range(n_rep):
        ref = oracle.init_random(L, seed, r)
        ro = oracle.sweep(ref, tab, seed, r, 0, n, every, boundary)
        assert np.array_equal(got[r], ref), (boundary, rows, slab, r)
        for f in ("sweep", "M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ro[f]), f
    assert ties == oracle.tie_count() > 500, (ties, oracle.tie_count())


def test_headline_plan_is_what_bench_runs(bc):
    """bench.py's workload (64 replicas of L = 4096) resolves by itself to 32-row strips with the pair table
    (8192 CTAs, 7 per SM: 7.9 waves)."""
    e = bc.Engine(4096, bc.PERIODIC, 64, 0, 1)
    e.set_params(*TC)
    plan = e.plan()
    e.close()
    assert plan["pair"] and plan["rows_per_thread"] == 32 and plan["grid"] == 8192, plan


@pytest.mark.parametrize("L,n_rep,rows,boundary", [(128, 2, 0, 0), (256, 1, 8, 0), (256, 2, 16, 1), (512, 2, 32, 0),
                                                   (1024, 1, 64, 0), (1024, 1, 32, 1), (4096, 1, 64, 0)])
def test_fused_halo_slab_kernels_match_oracle(bc, oracle, L, n_rep, rows, boundary):
    """The HALO instantiations of the streaming kernel (a slab's half-sweep as ONE launch: boundary CTAs first,
    in-kernel wait on the neighbours' epoch flags, halo rows stored into the neighbours' ghost rows, epoch signal
    from the last boundary CTA; plain / pair-table variants, graph replay) on ONE GPU: a one-rank slab
    connected to itself is its own ring neighbour.  Lattices and fused observables against the oracle."""
    n, every, seed = (20, 5, 0x4A10 + L + rows) if L <= 1024 else (4, 2, 0x4A10)
    e = bc.Engine(L, boundary, n_rep, 0, seed, slab=(1, 0, None))
    e.ipc_connect(None, None)
    e.set_option("rows_per_thread", rows)
    e.set_option("graph_block", 4)
    e.set_params(*TC)
    e.init_random()
    l0 = e.launch_count
    obs = e.sweep(n, every)
    # one kernel per half-sweep (+ the graph's clock kernels): no separate wait / push / signal launches
    assert e.launch_count - l0 <= 3 * n, (e.launch_count - l0, n)  # the unfused path issues 8 and more per sweep
    got = e.download()
    m = e.measure()
    e.sync()
    e.close()
    tab = oracle.build_table(*TC)
    for r in range(n_rep):
        ref = oracle.init_random(L, seed, r)
        ro = oracle.sweep(ref, tab, seed, r, 0, n, every, boundary)
        assert np.array_equal(got[r], ref), (L, rows, boundary)
        for f in ("sweep", "M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ro[f]), f
        assert (m[r]["M"], m[r]["Q"], m[r]["B"]) == oracle.observables(ref, boundary)


@pytest.mark.parametrize("L,n_rep", [(64, 3), (128, 1), (24, 2)])
def test_graph_replay_equals_direct_launches(bc, L, n_rep):
    """Blocks of unmeasured sweeps are replayed from a CUDA graph with a device-resident sweep
    counter; results must not depend on it (nor on how a run is split into bc_sweep calls)."""
    outs = []
    for use_graph, block, chunks in ((0, 16, (50,)), (1, 16, (50,)), (1, 4, (7, 20, 23)), (1, 16, (33, 17))):
        e = bc.Engine(L, 0, n_rep, 0, 2718)
        e.set_option("use_graph", use_graph)
        e.set_option("graph_block", block)
        e.set_params(*TC)
        e.init_random()
        obs = [e.sweep(c, 10) for c in chunks]
        outs.append((e.download(), np.concatenate([o["B"] for o in obs])))
        assert e.sweep_counter == 50
        e.close()
    for lat, b in outs[1:]:
        assert np.array_equal(lat, outs[0][0])
        assert np.array_equal(b, outs[0][1])


@pytest.mark.parametrize("L,n_rep,force", [(32, 1, -1), (32, 7, -1), (64, 1, -1), (64, 3, -1), (96, 2, -1), (128, 2, -1),
                                           (256, 2, 1), (448, 1, 1)])
@pytest.mark.parametrize("boundary", [0, 1])
def test_resident_kernel(bc, oracle, L, n_rep, force, boundary):
    """Small lattices stay in shared memory for the whole call (one launch for all sweeps)."""
    n = 9 if L <= 128 else 4
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, L, boundary, n_rep, n, 0x2E51 + L, measure_every=2,
                                            opts=(("resident", force),))
    assert plan["resident"], plan
    assert np.array_equal(got, ref)
    for r in range(n_rep):
        for f in ("sweep", "M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ref_obs[r][f]), (r, f)


@pytest.mark.parametrize("L,boundary,n_rep,force", [(2, 0, 5, -1), (4, 0, 3, -1), (6, 0, 9, -1), (8, 0, 4, -1), (12, 0, 5, -1),
                                                    (16, 0, 1, -1), (24, 0, 3, -1), (48, 0, 2, -1), (100, 0, 2, -1),
                                                    (200, 0, 2, 1), (34, 0, 3, -1), (62, 0, 2, -1),
                                                    (3, 1, 5, -1), (5, 1, 2, -1), (8, 1, 3, -1), (17, 1, 4, -1), (33, 1, 2, -1),
                                                    (31, 1, 2, -1), (48, 1, 2, -1), (65, 1, 2, -1), (127, 1, 1, -1)])
def test_resident_kernel_any_size(bc, oracle, L, boundary, n_rep, force):
    """Sizes that are not a multiple of 32 (the reference's 8, 12, 16, 24, 48, 96 of gen_files.py:11; odd L with open
    edges) also stay in shared memory: padded rows, wrap-around patches, masked stores and observables."""
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, L, boundary, n_rep, 11, 0xA11 + L, measure_every=1,
                                            opts=(("resident", force),), start="random" if L % 2 == 0 else "upload")
    assert plan["resident"], plan
    assert np.array_equal(got, ref)
    for r in range(n_rep):
        for f in ("sweep", "M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ref_obs[r][f]), (r, f)


@pytest.mark.parametrize("L,n_rep,boundary,rows", [(128, 1, 0, 0), (128, 5, 1, 2), (256, 2, 0, 0), (256, 1, 1, 4), (512, 1, 0, 0),
                                                   (512, 3, 1, 8), (1024, 1, 0, 0), (1024, 2, 1, 2)])
def test_persistent_kernel_matches_oracle(bc, oracle, L, n_rep, boundary, rows):
    """Runs of unmeasured sweeps in one cooperative launch with grid barriers between half-sweeps."""
    e_opts = (("resident", 0), ("persist", 1), ("rows_per_thread", rows))
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, L, boundary, n_rep, 14, 0x9E25 + L, measure_every=5, opts=e_opts)
    assert np.array_equal(got, ref)
    for r in range(n_rep):
        for f in ("sweep", "M", "Q", "B"):
            assert np.array_equal(obs[:, r][f], ref_obs[r][f]), (r, f)


def test_persistent_kernel_is_one_launch_and_equals_graph_replay(bc):
    res = []
    for persist in (1, 0):
        e = bc.Engine(2048, 0, 1, 0, 31)
        set_opt(e, "persist", persist)
        e.set_params(*TC)
        e.init_random()
        e.sync()
        l0 = e.launch_count
        e.sweep(50)
        n_launch = e.launch_count - l0
        e.sweep(7, 7)
        res.append((e.download(), n_launch, e.tie_count, e.measure()))
        e.sync()
        e.close()
    assert np.array_equal(res[0][0], res[1][0]) and res[0][2] == res[1][2]
    assert np.array_equal(res[0][3], res[1][3])
    assert res[0][1] == 1 and res[1][1] >= 100


def test_resident_is_one_launch_and_equals_streaming(bc):
    res = []
    for resident in (1, 0):
        e = bc.Engine(64, 0, 5, 0, 99)
        e.set_option("resident", resident)
        e.set_params(*TC)
        e.init_random()
        e.sync()
        l0 = e.launch_count
        a = e.sweep(40, 4)
        b = e.sweep(17, 4)
        n_launch = e.launch_count - l0
        res.append((e.download(), np.concatenate([a["B"], b["B"]]), n_launch, e.tie_count))
        e.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert res[0][2] == 2 and res[1][2] > 100
    assert res[0][3] == res[1][3]


def test_async_transfers_and_two_handles(bc, oracle):
    """bc_upload_all_async / bc_download_all_async on two handles (two streams) used ping-pong, as
    bench.py's end-to-end loop does; every batch must come back exactly as the oracle computes it."""
    L, n_rep, n = 64, 3, 6
    tab = oracle.build_table(*TC)
    engs = [bc.Engine(L, 0, n_rep, 0, 1000 + k) for k in range(2)]
    for e in engs:
        e.set_params(*TC)
    rng = np.random.default_rng(8)
    bufs = [np.ascontiguousarray(rng.integers(-1, 2, size=(n_rep, L, L), dtype=np.int8)) for _ in range(2)]
    want = [b.copy() for b in bufs]
    counters = [0, 0]
    for step in range(4):
        k = step & 1
        e = engs[k]
        e.sync()
        e.upload_async(bufs[k])
        e.sweep(n, 0, fetch=False)
        e.download_async(bufs[k])
        for r in range(n_rep):
            oracle.sweep(want[k][r], tab, 1000 + k, r, counters[k], n, 0)
        counters[k] += n
    for k in range(2):
        engs[k].sync()
        assert np.array_equal(bufs[k], want[k]), k
        engs[k].close()


def test_resident_per_replica_parameters_and_counter_offset(bc, oracle):
    L, n_rep, seed = 64, 4, 31
    e = bc.Engine(L, 1, n_rep, 0, seed)
    pts = [(0.5 + 0.1 * r, 1.9 + 0.02 * r, 1.0) for r in range(n_rep)]
    for r, p in enumerate(pts):
        e.set_params(*p, replica=r)
    e.init_random()
    e.sweep_counter = (1 << 33) + 5  # exercises the high word of the Philox counter
    obs = e.sweep(7, 3)
    assert e.plan()["resident"]
    got = e.download()
    for r, p in enumerate(pts):
        ref = oracle.init_random(L, seed, r)
        ro = oracle.sweep(ref, oracle.build_table(*p), seed, r, (1 << 33) + 5, 7, 3, 1)
        assert np.array_equal(got[r], ref), r
        assert np.array_equal(obs[:, r]["B"], ro["B"]) and np.array_equal(obs[:, r]["sweep"], ro["sweep"])
    e.close()


def test_randomised_configurations(bc, oracle):
    """40 seeded random configurations across every dispatch path (generic / streaming / resident / slab
    handle), both boundaries, ragged sizes, random parameters, replica counts, call splits and counter
    offsets: lattice and observables must equal the oracle bit for bit."""
    rng = np.random.default_rng(20261020)
    for case in range(40):
        boundary = int(rng.integers(0, 2))
        kind = rng.choice(["generic", "stream", "resident", "slab", "resident_any"])
        if kind in ("generic", "resident_any"):
            L = int(rng.choice([2, 4, 6, 10, 14, 18, 22, 30, 34, 50]) if boundary == 0 else rng.integers(2, 60))
        else:
            L = int(rng.choice([32, 64, 96, 128, 160, 256]))
        n_rep = int(rng.integers(1, 5))
        T = float(rng.uniform(0.3, 3.0)); D = float(rng.uniform(-2.0, 2.5)); J = float(rng.choice([1.0, 1.0, 0.5, -1.0]))
        seed = int(rng.integers(0, 2**63))
        t0 = int(rng.choice([0, 7, (1 << 32) - 3, (1 << 40) + 1]))
        chunks = [int(c) for c in rng.integers(1, 6, size=int(rng.integers(1, 4)))]
        me = int(rng.choice([0, 1, 2, 3]))
        slab = (1, 0, None) if kind == "slab" else None
        e = bc.Engine(L, boundary, n_rep, 0, seed, slab=slab)
        if kind == "stream":
            e.set_option("resident", 0)
            e.set_option("rows_per_thread", int(rng.choice([0, 1, 2, 4, 8])))
            e.set_option("pair_min_rows", int(rng.choice([0, 2, 8])))
        if kind == "resident":
            e.set_option("resident", 1)
        if kind == "generic":
            e.set_option("resident", 0)  # one thread per site instead of the padded shared-memory kernel
        e.set_params(T, D, J)
        e.init_random()
        e.sweep_counter = t0
        tab = oracle.build_table(T, D, J)
        ref = np.stack([oracle.init_random(L, seed, r) for r in range(n_rep)])
        t = t0
        for c in chunks:
            obs = e.sweep(c, me)
            for r in range(n_rep):
                ro = oracle.sweep(ref[r], tab, seed, r, t, c, me, boundary)
                if me:
                    for f in ("sweep", "M", "Q", "B"):
                        assert np.array_equal(obs[:, r][f], ro[f]), (case, kind, L, boundary, f)
            t += c
        assert np.array_equal(e.download(), ref), (case, kind, L, boundary, n_rep, T, D, J, t0, chunks)
        e.close()


@pytest.mark.parametrize("L,n_rep,rows", [(128, 1, 4), (128, 3, 8), (256, 2, 4), (512, 1, 4), (1024, 1, 8)])
@pytest.mark.parametrize("boundary", [0, 1])
def test_wavefront_kernel_matches_oracle(bc, oracle, L, n_rep, rows, boundary):
    """All half-sweeps of a call in one launch, ordered by per-block progress counters instead of kernel
    boundaries; must equal the oracle (and hence the launch-per-half-sweep path) bit for bit."""
    n = 7
    got, ref, obs, ref_obs, plan = run_pair(bc, oracle, L, boundary, n_rep, n, 0x3A7E + L, measure_every=0,
                                            opts=(("resident", 0), ("wave", 1), ("rows_per_thread", rows)))
    assert np.array_equal(got, ref)


def test_wavefront_kernel_large_and_launch_count(bc):
    """L = 4096 and 8192 (one and two 128-chunk blocks per strip): wavefront == streaming, one launch per call."""
    for L, n in ((4096, 6), (8192, 4)):
        res = []
        for wave in (1, 0):
            e = bc.Engine(L, 0, 1, 0, 0xABC)
            set_opt(e, "wave", wave)
            e.set_params(*TC)
            e.init_random()
            e.sync()
            l0 = e.launch_count
            e.sweep(n)
            launches = e.launch_count - l0
            e.sync()
            res.append((e.download(0), launches))
            e.close()
        assert np.array_equal(res[0][0], res[1][0]), L
        assert res[0][1] == 1 and res[1][1] >= 2 * n


@pytest.mark.parametrize("L,boundary,n_rep,slab", [(64, 0, 3, False), (4096, 0, 1, False), (100, 0, 2, False), (33, 1, 2, False),
                                                 (256, 0, 2, True)])
def test_packed_wire_format_roundtrip(bc, L, boundary, n_rep, slab):
    """bc_upload_packed / bc_download_packed (2 bits per site on the wire, converted on the device) move the same
    lattices as the int8 calls: vector kernels for L % 64 == 0, per-byte kernels otherwise, slab geometry included."""
    rng = np.random.default_rng(L)
    s = rng.integers(-1, 2, size=(n_rep, L, L), dtype=np.int8)
    e = bc.Engine(L, boundary, n_rep, 0, 5, slab=(1, 0, None) if slab else None)
    e.set_params(*TC)
    e.upload_packed(bc.pack_spins(s))
    assert np.array_equal(e.download(), s)
    e.sweep(3)
    want = e.download()
    got = bc.unpack_spins(e.download_packed(), L).reshape(n_rep, L, L)
    assert np.array_equal(got, want)
    one = bc.unpack_spins(e.download_packed(replica=n_rep - 1), L)
    assert np.array_equal(one, want[n_rep - 1])
    e.upload_packed(bc.pack_spins(s[0]), replica=n_rep - 1)
    assert np.array_equal(e.download(n_rep - 1), s[0])
    e.close()
