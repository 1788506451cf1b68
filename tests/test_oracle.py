"""CPU tests that pin the oracle (no GPU).

The reference has no tests or golden vectors for the Metropolis path, so
the oracle is pinned by: Random123 Philox4x32-10 known answers, the literal reference
expression for the 54 thresholds (main.cpp:94,105-106,109), exact enumeration of small tori
(BASELINE.md section 3), and statistics of the unmodified reference (tests/golden).
"""
import json
import math
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TC = (0.608, 1.966, 1.0)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kats:
        assert list(oracle.philox(ctr, key)) == want


@pytest.mark.parametrize("T,D,J", [TC, (2.2691853, -1000.0, 1.0), (1.0, 0.5, 0.0), (0.3, 1.0, 1.0), (5.0, -2.0, -1.0)])
def test_table_against_literal_reference_expression(oracle, T, D, J):
    tab = oracle.build_table(T, D, J)
    B = 1 / T
    for cur in (-1, 0, 1):
        for flip in (0, 1):
            for nb in range(-4, 5):
                # literal main.cpp:94, :105-106, :109
                proposal = (cur + 2 + flip) % 3 - 1
                de = -J * (proposal - cur) * nb + D * (proposal * proposal - cur * cur)
                want = 2**31 if de <= 0 else min(2**31, max(0, int(math.floor(math.exp(-B * de) * 2**31 + 0.5))))
                got = int(tab[18 * (cur + 1) + 9 * flip + (nb + 4)])
                assert got == want, (cur, flip, nb)
                de_o, p_o = oracle.reference_de(cur, flip, nb, D, J)
                assert p_o == proposal and de_o == de
                assert proposal != cur and proposal in (-1, 0, 1)
    # spin-flip symmetry of the Hamiltonian: entry i mirrors entry 53 - i
    assert np.array_equal(tab, tab[::-1])


def test_table_golden_fixture(oracle):
    with open(os.path.join(GOLD, "tables.json")) as f:
        gold = json.load(f)
    for item in gold:
        assert list(map(int, oracle.build_table(item["T"], item["D"], item["J"]))) == item["thr31"]


def test_impossible_moves_have_zero_threshold(oracle):
    tab = oracle.build_table(2.2691853, -1000.0, 1.0)  # Ising limit of main.cpp:27-29
    # +-1 -> 0 costs ~1000 J: never accepted; 0 -> +-1 always accepted
    for S in range(9):
        assert tab[18 * 0 + 9 * 0 + S] == 0      # -1 -> 0
        assert tab[18 * 2 + 9 * 1 + S] == 0      # +1 -> 0
        assert tab[18 * 1 + 9 * 0 + S] == 2**31  # 0 -> +1
        assert tab[18 * 1 + 9 * 1 + S] == 2**31  # 0 -> -1


def test_exact_enumeration_matches_baseline_table(oracle):
    want = {2: (0.09138349, 0.66747018, 0.66832228, 0.49549719, 0.70152109),
            3: (0.12142487, 0.66077427, 0.66441978, 0.49535728, 1.41112734)}
    for L, w in want.items():
        got = oracle.enumerate_exact(L, *TC)
        assert np.allclose([got[k] for k in ("E", "absm", "q", "U4", "chi")], w, atol=2e-8)


def test_golden_trajectories(oracle):
    """Committed fixtures (tests/golden/make_golden.py): lattices + observables after n sweeps."""
    d = np.load(os.path.join(GOLD, "trajectories.npz"))
    meta = json.loads(str(d["meta"]))
    for i, m in enumerate(meta):
        tab = oracle.build_table(m["T"], m["D"], m["J"])
        s = oracle.init_random(m["L"], m["seed"], m["replica"])
        assert np.array_equal(s, d[f"init_{i}"])
        obs = oracle.sweep(s, tab, m["seed"], m["replica"], 0, m["n_sweeps"], 1, m["boundary"])
        assert np.array_equal(s, d[f"final_{i}"])
        assert np.array_equal(np.stack([obs["M"], obs["Q"], obs["B"]]), d[f"obs_{i}"])


def test_oracle_statistics_vs_exact_enumeration(oracle):
    """Checkerboard + Philox + integer thresholds sample the Boltzmann distribution of the
    reference Hamiltonian: L=4 torus, 3e5 sweeps, against brute-force enumeration."""
    L, n = 4, 300_000
    exact = {"E": 0.13644431, "absm": 0.65236652, "q": 0.66068962, "U4": 0.49561525}
    s = oracle.init_random(L, 2025, 0)
    obs = oracle.sweep(s, oracle.build_table(*TC), 2025, 0, 0, n, 1)[2000:]
    N = L * L
    m = obs["M"] / N
    E = (-TC[2] * obs["B"] + TC[1] * obs["Q"]) / N
    q = obs["Q"] / N

    def err(x, nb=50):  # binning error
        b = x[: len(x) // nb * nb].reshape(nb, -1).mean(1)
        return b.std(ddof=1) / math.sqrt(nb)

    assert abs(E.mean() - exact["E"]) < 5 * err(E)
    assert abs(np.abs(m).mean() - exact["absm"]) < 5 * err(np.abs(m))
    assert abs(q.mean() - exact["q"]) < 5 * err(q)
    # U4 is a ratio of moments: jackknife over 50 bins for its error bar
    nb = 50
    mb = m[: len(m) // nb * nb].reshape(nb, -1)

    def u4_of(x):
        return 1 - (x**4).mean() / (3 * (x**2).mean() ** 2)

    jk = np.array([u4_of(np.delete(mb, i, axis=0)) for i in range(nb)])
    u4_err = math.sqrt((nb - 1) / nb * ((jk - jk.mean()) ** 2).sum())
    assert abs(u4_of(mb) - exact["U4"]) < 5 * u4_err and u4_err < 0.01, (u4_of(mb), exact["U4"], u4_err)


def test_J0_limit(oracle):
    T, D = 1.0, 0.7
    L = 32
    s = oracle.init_random(L, 1, 0)
    obs = oracle.sweep(s, oracle.build_table(T, D, 0.0), 1, 0, 0, 600, 1)[100:]
    q = obs["Q"].mean() / (L * L)
    q_exact = 2 * math.exp(-D / T) / (1 + 2 * math.exp(-D / T))
    assert abs(q - q_exact) < 4e-3


def test_observables_and_open_boundary(oracle):
    rng = np.random.default_rng(1)
    for L in (3, 4, 7, 16):
        s = rng.integers(-1, 2, size=(L, L), dtype=np.int8)
        s64 = s.astype(np.int64)
        M, Q, B = oracle.observables(s, oracle.PERIODIC)
        assert (M, Q) == (s64.sum(), (s64 * s64).sum())
        assert B == (s64 * np.roll(s64, -1, 1)).sum() + (s64 * np.roll(s64, -1, 0)).sum()
        _, _, Bo = oracle.observables(s, oracle.OPEN)
        assert Bo == (s64[:, :-1] * s64[:, 1:]).sum() + (s64[:-1] * s64[1:]).sum()


def test_periodic_odd_L_rejected(oracle):
    s = np.zeros((5, 5), np.int8)
    with pytest.raises(ValueError):
        oracle.sweep(s, oracle.build_table(*TC), 1, 0, 0, 1, 0, oracle.PERIODIC)
    oracle.sweep(s, oracle.build_table(*TC), 1, 0, 0, 1, 0, oracle.OPEN)


def test_row_range_half_sweeps_compose(oracle):
    """Slab decomposition on the CPU: updating row ranges separately == whole-lattice half-sweep
    (the RNG counter depends only on global coordinates)."""
    L = 16
    tab = oracle.build_table(*TC)
    a = oracle.init_random(L, 9, 0)
    b = a.copy()
    oracle.sweep(a, tab, 9, 0, 0, 3, 0)
    for t in range(3):
        for col in (0, 1):
            snap = b.copy()  # both slabs read the pre-half-sweep state of the other colour
            parts = []
            for (y0, y1) in ((0, 5), (5, 16)):
                w = snap.copy()
                oracle.half_sweep_rows(w, tab, 9, 0, t, col, y0, y1)
                parts.append(w[y0:y1])
            b = np.concatenate(parts, 0)
    assert np.array_equal(a, b)


def test_cluster_move_preserves_the_boltzmann_distribution(oracle):
    """Metropolis sweeps interleaved with Swendsen-Wang moves on the +-1 spins against exact enumeration
    (fast-mixing points; the tricritical point is covered on the GPU with thousands of replicas)."""
    for (T, D) in ((2.0, 0.0), (2.27, -1000.0)):
        ex = oracle.enumerate_exact(4, T, D, 1.0)
        tab = oracle.build_table(T, D, 1.0)
        s = oracle.init_random(4, 5, 0)
        E, M = [], []
        for it in range(30000):
            oracle.sweep(s, tab, 5, 0, it, 1, 0)
            oracle.cluster_update(s, T, 1.0, 5, 0, it)
            m, q, b = oracle.observables(s)
            E.append((-b + D * q) / 16)
            M.append(abs(m) / 16)
        assert abs(np.mean(E[1000:]) - ex["E"]) < 6e-3, (T, D)
        assert abs(np.mean(M[1000:]) - ex["absm"]) < 6e-3, (T, D)


def test_cluster_move_never_touches_zeros_and_only_flips_signs(oracle):
    rng = np.random.default_rng(4)
    s = rng.integers(-1, 2, size=(16, 16), dtype=np.int8)
    before = s.copy()
    oracle.cluster_update(s, 0.608, 1.0, 11, 0, 0)
    assert np.array_equal(s == 0, before == 0)
    assert np.array_equal(np.abs(s), np.abs(before))
    # J = 0: no bonds, every non-zero site is its own cluster and flips with probability 1/2
    t = np.ones((64, 64), np.int8)
    oracle.cluster_update(t, 1.0, 0.0, 11, 0, 0)
    assert 0.4 < (t == -1).mean() < 0.6


def test_cluster_move_golden_fixture(oracle):
    """The cluster-move spec (bond / flip streams, smallest-site roots) is pinned by committed vectors."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_cluster_golden", os.path.join(GOLD, "make_cluster_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    z = np.load(os.path.join(GOLD, "cluster_moves.npz"))
    cases = json.loads(str(z["meta"]))
    assert cases == gen.CASES
    for i, m in enumerate(cases):
        s, mom = gen.run_case(m)
        assert np.array_equal(s, z[f"final_{i}"]), i
        assert np.array_equal(mom, z[f"moments_{i}"]), i
