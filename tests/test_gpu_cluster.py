"""f-3: Swendsen-Wang cluster move on the GPU against the CPU oracle (bit-exact) and exact enumeration."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TC = (0.608, 1.966, 1.0)


@pytest.mark.parametrize("L,boundary,T,D", [(8, 0, 0.608, 1.966), (24, 0, 2.0, 0.0), (64, 0, 0.608, 1.966), (64, 1, 1.2, 0.5),
                                            (33, 1, 0.608, 1.966), (256, 0, 2.2691853, -1000.0), (512, 0, 0.5, 1.0)])
def test_cluster_update_matches_oracle(bc, oracle, L, boundary, T, D):
    n_rep, seed = 2, 0xC1 + L
    e = bc.Engine(L, boundary, n_rep, 0, seed)
    e.set_params(T, D, 1.0)
    e.init_random()
    ref = np.stack([oracle.init_random(L, seed, r) for r in range(n_rep)])
    tab = oracle.build_table(T, D, 1.0)
    for step in range(3):  # sweeps and cluster moves interleaved, like step() of main.cpp:158-163
        e.sweep(2)
        e.cluster_update(1)
        for r in range(n_rep):
            oracle.sweep(ref[r], tab, seed, r, 2 * step, 2, 0, boundary)
            oracle.cluster_update(ref[r], T, 1.0, seed, r, step, boundary)
        assert np.array_equal(e.download(), ref), (step,)
    assert e.cluster_counter == 3
    e.close()


def test_cluster_update_low_temperature_giant_cluster(bc, oracle):
    """All-plus start at low T: one cluster spans the lattice (long union-find chains)."""
    L, seed = 128, 9
    s = np.ones((1, L, L), np.int8)
    s[0, 5, 7] = 0
    s[0, 40:60, 10:30] = -1
    e = bc.Engine(L, 0, 1, 0, seed)
    e.set_params(0.3, 1.0, 1.0)
    e.upload(s)
    e.cluster_update(2)
    ref = s[0].copy()
    for c in range(2):
        oracle.cluster_update(ref, 0.3, 1.0, seed, 0, c)
    assert np.array_equal(e.download(0), ref)
    e.close()


@pytest.mark.parametrize("L,boundary", [(320, 0), (200, 1), (70, 0)])
def test_clusters_spanning_many_tiles(bc, oracle, L, boundary):
    """The labelling works tile by tile (64 x 64 in shared memory) and merges across tile borders afterwards:
    a spiral-like ordered configuration whose clusters cross every border, sizes that clip the last tiles."""
    seed, T = 4242, 0.9
    s = np.ones((2, L, L), np.int8)
    s[0, ::7, : L - 3] = -1          # long horizontal stripes, connected to nothing else
    s[0, 3::7, 5:] = 0
    s[1, :, ::5] = -1                # vertical stripes
    s[1, L // 2, :] = 1              # one row joins all the +1 columns
    e = bc.Engine(L, boundary, 2, 0, seed)
    e.set_params(T, 0.0, 1.0)
    e.upload(s)
    for kind in (0, 1):
        got = e.cluster_moments(kind, mstep=3)
        for r in range(2):
            want = oracle.cluster_moments(s[r], kind, T, 1.0, seed, r, 3, boundary)
            assert (got[r]["sum_size_sq"], got[r]["max_size"], got[r]["n_clusters"]) == want, (kind, r)
    ref = s.copy()
    for c in range(3):
        e.cluster_update(1)
        for r in range(2):
            oracle.cluster_update(ref[r], T, 1.0, seed, r, c, boundary)
        assert np.array_equal(e.download(), ref), c
    e.close()


def test_metropolis_plus_cluster_statistics_L4(bc):
    """Sweeps interleaved with cluster moves still sample the Boltzmann distribution (exact enumeration)."""
    exact = np.array([0.13644431, 0.65236652, 0.66068962])
    L, n_rep = 4, 4096
    e = bc.Engine(L, 0, n_rep, 0, 77)
    e.set_params(*TC)
    e.init_random()
    acc = []
    for it in range(400):
        e.sweep(5)
        e.cluster_update(1)
        if it >= 100:
            m = e.measure()
            N = L * L
            acc.append([((-TC[2] * m["B"] + TC[1] * m["Q"]) / N).mean(), (np.abs(m["M"]) / N).mean(), (m["Q"] / N).mean()])
    e.close()
    acc = np.array(acc)
    mean = acc.mean(0)
    err = acc.std(0, ddof=1) / np.sqrt(len(acc) / 4)  # crude allowance for autocorrelation
    assert np.all(np.abs(mean - exact) < 5 * err + 5e-4), (mean, err)


def test_cluster_update_per_replica_temperatures(bc, oracle):
    """BASELINE config 3 is an 8 x 8 (T, D) grid: every replica of a handle has its own bond probability
    1 - exp(-2J/T) (main.cpp:115), so the cluster move works on a ladder handle with 8 distinct temperatures."""
    L, n_rep, seed = 128, 8, 0x7E3
    Ts = np.linspace(0.58, 0.64, n_rep)
    e = bc.Engine(L, 0, n_rep, 0, seed)
    for r, T in enumerate(Ts):
        e.set_params(float(T), 1.94 + 0.005 * r, 1.0, replica=r)
    e.init_random()
    ref = np.stack([oracle.init_random(L, seed, r) for r in range(n_rep)])
    for step in range(2):
        e.sweep(3)
        e.cluster_update(1)
        for r, T in enumerate(Ts):
            oracle.sweep(ref[r], oracle.build_table(float(T), 1.94 + 0.005 * r, 1.0), seed, r, 3 * step, 3, 0, 0)
            oracle.cluster_update(ref[r], float(T), 1.0, seed, r, step, 0)
    assert np.array_equal(e.download(), ref)
    e.close()


@pytest.mark.parametrize("boundary", [0, 1])
def test_cluster_update_on_a_one_rank_slab(bc, oracle, boundary):
    """Slab handles (ghost rows in the colour arrays) take the cluster move when they hold the whole lattice."""
    L, seed, T = 192, 0x51AB, 0.608
    e = bc.Engine(L, boundary, 1, 0, seed, slab=(1, 0, None))
    e.set_params(T, 1.966, 1.0)
    e.init_random()
    ref = oracle.init_random(L, seed, 0)
    tab = oracle.build_table(T, 1.966, 1.0)
    for step in range(2):
        e.cluster_update(1)
        e.sweep(2)  # reads the ghost rows the cluster move has to refresh
        oracle.cluster_update(ref, T, 1.0, seed, 0, step, boundary)
        oracle.sweep(ref, tab, seed, 0, 2 * step, 2, 0, boundary)
    assert np.array_equal(e.download(0), ref)
    e.close()


def test_cluster_update_full_size(bc, oracle):
    """L = 4096 (4096 tiles, clusters across thousands of tile borders) against the oracle."""
    L, seed, T = 4096, 0xF00D, 0.608
    e = bc.Engine(L, 0, 1, 0, seed)
    e.set_params(T, 1.966, 1.0)
    e.init_random()
    e.sweep(20)
    s = e.download(0)
    e.cluster_update(1)
    ref = s.copy()
    oracle.cluster_update(ref, T, 1.0, seed, 0, 0, 0)
    assert np.array_equal(e.download(0), ref)
    want = oracle.cluster_moments(ref, 1, T, 1.0, seed, 0, 2, 0)
    got = e.cluster_moments(1, mstep=2)[0]
    assert (got["sum_size_sq"], got["max_size"], got["n_clusters"]) == want
    e.close()


def test_cluster_update_errors(bc):
    e = bc.Engine(32, 0, 2, 0, 1)
    e.set_params(1.0, 0.0, 1.0, replica=0)
    with pytest.raises(bc.BCError):
        e.cluster_update(1)  # parameters of replica 1 not set
    e.close()


def test_cluster_moves_match_the_committed_vectors(bc):
    """tests/golden/cluster_moves.npz (written by make_cluster_golden.py from the oracle): the GPU reproduces the
    lattices after interleaved sweeps and cluster moves and the cluster-size moments, without the oracle in the loop."""
    import json
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "cluster_moves.npz"))
    for i, m in enumerate(json.loads(str(z["meta"]))):
        n_rep = m["replica"] + 1  # the fixture's replica index selects the Philox key
        e = bc.Engine(m["L"], m["boundary"], n_rep, 0, m["seed"])
        e.set_params(m["T"], m["D"], m["J"])
        e.init_random()
        for c in range(m["rounds"]):
            e.sweep(m["sweeps"])
            e.cluster_update(1)
        assert np.array_equal(e.download(m["replica"]), z[f"final_{i}"]), i
        for kind in (0, 1):
            got = e.cluster_moments(kind, mstep=5)[m["replica"]]
            assert (got["sum_size_sq"], got["max_size"], got["n_clusters"]) == tuple(z[f"moments_{i}"][kind]), (i, kind)
        e.close()
