"""f-4 on the GPU: gap-size histogram, cluster lists and dump text from device-side labels."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "blume_capel_b200", "host")
TC = (0.608, 1.966, 1.0)


def gap_text_from_labels(spins, labels):
    """The cluster file export_clusters writes (main.cpp:236-256) from per-site cluster labels (smallest site)."""
    flat, lab = spins.reshape(-1), labels.reshape(-1).astype(np.int64)
    occ = np.flatnonzero(flat != 0)
    order = occ[np.argsort(lab[occ], kind="stable")]
    keys = lab[order]
    out = []
    bounds = np.flatnonzero(np.diff(keys)) + 1
    for grp in np.split(order, bounds):
        if grp.size == 0:
            continue
        deltas = np.diff(np.concatenate(([0], grp)))
        out.append(("+ " if flat[grp[0]] > 0 else "- ") + "".join(f"{d} " for d in deltas) + "\n")
    return "".join(out) + "\n"


@pytest.mark.parametrize("L,boundary,n_rep", [(8, 0, 3), (48, 0, 4), (64, 0, 3), (64, 1, 2), (100, 0, 2), (129, 1, 2), (256, 0, 1)])
def test_gap_histogram_matches_oracle(bc, oracle, L, boundary, n_rep):
    """bc_gap_histogram (labels, radix sort by cluster, one thread per member) against bc_oracle_gap_histogram, which is
    pinned to the unmodified gap_size_statistics tool (tests/test_oracle.py); spin and FK clusters, the reference's
    variant and the open variant, sizes that clip the last tiles."""
    seed = 0x6A9 + L
    e = bc.Engine(L, boundary, n_rep, 0, seed)
    e.set_params(*TC)
    if L % 2 == 0 or boundary == 1:
        rng = np.random.default_rng(L)
        s = rng.integers(-1, 2, size=(n_rep, L, L), dtype=np.int8)
        e.upload(s)
    e.sweep(30)
    e.cluster_update(1)
    s = e.download()
    for kind in (0, 1):
        for variant in (0, 1):
            got = e.gap_histogram(kind, mstep=11, variant=variant)
            got = e.gap_histogram(kind, mstep=11, variant=variant, out=got)  # accumulates: two identical samples
            for r in range(n_rep):
                want = oracle.gap_histogram(s[r], kind, TC[0], TC[2], seed, r, 11, boundary, variant)
                assert np.array_equal(got[r], 2 * want), (kind, variant, r)
    assert np.array_equal(e.download(), s)  # measuring does not change the lattices
    e.close()


@pytest.mark.parametrize("L,boundary", [(16, 0), (64, 0), (96, 1), (320, 0)])
def test_cluster_lists_give_the_reference_file_bytes(bc, oracle, tmp_path, L, boundary):
    """bc_cluster_lists (clusters formed on the GPU) printed by the host == the file the host union-find writer
    produces for the same lattice (which tests/test_gap_format.py pins byte for byte to dumps of the unmodified
    reference), and == the text built from the oracle's labels for FK clusters."""
    seed = 0x11575 + L
    e = bc.Engine(L, boundary, 2, 0, seed)
    e.set_params(*TC)
    e.init_random()
    e.sweep(40)
    s = e.download()
    gaptool = os.path.join(HOST, "bc_gaptool")
    for r in range(2):
        for kind in (0, 1):
            sites, starts, signs = e.cluster_lists(r, kind, mstep=4)
            text = "".join(("+ " if signs[c] > 0 else "- ") +
                           "".join(f"{d} " for d in np.diff(np.concatenate(([0], sites[starts[c]:starts[c + 1]].astype(np.int64))))) + "\n"
                           for c in range(len(signs))) + "\n"
            lab = oracle.cluster_labels(s[r], kind, TC[0], TC[2], seed, r, 4, boundary)
            assert text == gap_text_from_labels(s[r], lab), (r, kind)
            if kind == 0:
                raw, txt = tmp_path / "s.bin", tmp_path / "s.txt"
                raw.write_bytes(s[r].tobytes())
                subprocess.run([gaptool, "write", str(L), "periodic" if boundary == 0 else "open", str(raw), str(txt)], check=True)
                assert txt.read_text() == text
    e.close()


def test_gap_histogram_full_size(bc, oracle):
    """L = 4096 (the size the text pipeline cannot handle: about 50 MB of cluster text per sample) against the oracle."""
    L, seed = 4096, 0x9A9
    e = bc.Engine(L, 0, 1, 0, seed)
    e.set_params(*TC)
    e.init_random()
    e.sweep(30)
    s = e.download(0)
    got = e.gap_histogram(0, 0, 0)[0]
    want = oracle.gap_histogram(s, 0, TC[0], TC[2], seed, 0, 0, 0, 0)
    assert np.array_equal(got, want)
    sites, starts, signs = e.cluster_lists(0, 0, 0)
    lab = oracle.cluster_labels(s, 0, TC[0], TC[2], seed, 0, 0, 0).reshape(-1)
    occ = np.flatnonzero(s.reshape(-1) != 0)
    order = occ[np.argsort(lab[occ], kind="stable")]
    assert np.array_equal(sites.astype(np.int64), order)
    assert np.array_equal(sites[starts[:-1]].astype(np.int64), np.unique(lab[occ]))
    e.close()


def test_bc_gap_stats_gpu_is_byte_identical_to_the_reference_tool(bc, tmp_path):
    """The unmodified gap_size_statistics binary (oracle/_ref, built from corner_contribution/gap_size_statistics.cpp)
    on cluster files of 100 runs of L = 48 and 64 (its hard-wired sizes and run count, :171-186) against
    bc_gap_histogram on the same configurations: the per-run files must be the same bytes."""
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "gap_size_statistics")
    gaptool = os.path.join(HOST, "bc_gaptool")
    if not os.access(ref_tool, os.X_OK):
        pytest.skip("oracle/_ref/gap_size_statistics not built")
    runs, samples = 100, 2
    inp, out = tmp_path / "in", tmp_path / "out"
    mine = {}
    for L in (48, 64):
        (out / str(L)).mkdir(parents=True)
        e = bc.Engine(L, 0, runs, 0, 77 + L)
        e.set_params(*TC)
        e.init_random()
        hist = np.zeros((runs, L // 2), np.int64)
        for k in range(samples):
            e.sweep(25)
            e.cluster_update(1)
            e.gap_histogram(0, 0, 0, out=hist)
            s = e.download()
            for r in range(runs):
                d = inp / str(L) / str(r)
                d.mkdir(parents=True, exist_ok=True)
                raw = tmp_path / "l.bin"
                raw.write_bytes(s[r].tobytes())
                subprocess.run([gaptool, "write", str(L), "periodic", str(raw), str(d / f"{k}.txt")], check=True)
        e.close()
        mine[L] = hist
    subprocess.run([ref_tool, str(inp), str(out)], check=True)
    for L in (48, 64):
        for r in range(runs):
            want = (out / str(L) / f"{r}.txt").read_text()
            got = f"{samples}\n" + "".join(f"{v} " for v in mine[L][r]) + "\n"
            assert got == want, (L, r)


def test_bc_gap_stats_cli_on_gpu(tmp_path):
    tool = os.path.join(HOST, "bc_gap_stats")
    subprocess.run([tool, "run", str(tmp_path / "gap"), "0.608", "1.966", "1.0", "--L", "48,4096", "--runs", "10", "--samples", "2",
                    "--burn-sweeps", "40", "--sweeps-per-sample", "10", "--kind", "fk", "--seed", "2", "--corner", str(tmp_path / "c.csv")],
                   check=True)
    for L in (48, 4096):
        rows = (tmp_path / "gap" / str(L) / "3.txt").read_text().split("\n")
        assert rows[0] == "2" and len(rows[1].split()) == L // 2 and sum(map(int, rows[1].split())) > 0
    c = (tmp_path / "c.csv").read_text().strip().splitlines()
    assert c[0] == "Batch,L,Corner_Contribution,Error" and len(c) == 3 and float(c[1].split(",")[2]) > 0
    # open-boundary variant: L - 1 bins
    subprocess.run([tool, "run", str(tmp_path / "gapo"), "0.608", "1.966", "1.0", "--L", "64", "--runs", "4", "--samples", "2",
                    "--burn-sweeps", "40", "--boundary", "open", "--seed", "2"], check=True)
    assert len((tmp_path / "gapo" / "64" / "0.txt").read_text().split("\n")[1].split()) == 63


def test_cluster_measurements_on_a_one_rank_slab(bc, oracle):
    """Slab handles that hold the whole lattice (ghost rows in the colour arrays) take the cluster measurements too."""
    L, seed = 128, 0x51AC
    e = bc.Engine(L, 0, 2, 0, seed, slab=(1, 0, None))
    e.set_params(*TC)
    e.init_random()
    e.sweep(20)
    s = e.download()
    got = e.gap_histogram(1, mstep=3, variant=0)
    mom = e.cluster_moments(1, mstep=3)
    for r in range(2):
        assert np.array_equal(got[r], oracle.gap_histogram(s[r], 1, TC[0], TC[2], seed, r, 3, 0, 0))
        assert (mom[r]["sum_size_sq"], mom[r]["max_size"], mom[r]["n_clusters"]) == oracle.cluster_moments(s[r], 1, TC[0], TC[2], seed, r, 3, 0)
    sites, starts, signs = e.cluster_lists(1, 0, 0)
    lab = oracle.cluster_labels(s[1], 0, TC[0], TC[2], seed, 1, 0, 0).reshape(-1)
    occ = np.flatnonzero(s[1].reshape(-1) != 0)
    assert np.array_equal(sites.astype(np.int64), occ[np.argsort(lab[occ], kind="stable")])
    e.close()


def test_cluster_list_cache_follows_the_lattice(bc, oracle):
    """The labelling and the sort are cached per (lattice state, kind, mstep): a sweep, a cluster move or an upload
    in between must invalidate them."""
    L = 64
    e = bc.Engine(L, 0, 1, 0, 3)
    e.set_params(*TC)
    e.init_random()
    a = e.cluster_lists(0, 0, 0)
    b = e.cluster_lists(0, 0, 0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    e.sweep(1)
    c = e.cluster_lists(0, 0, 0)
    s = e.download(0)
    assert len(c[0]) == int((s != 0).sum()) and not (len(a[0]) == len(c[0]) and np.array_equal(a[0], c[0]))
    e.upload(np.ones((1, L, L), np.int8))
    d = e.cluster_lists(0, 0, 0)
    assert len(d[2]) == 1 and len(d[0]) == L * L and d[2][0] == 1  # one cluster: the whole all-plus lattice
    h = e.gap_histogram(0, 0, 0)
    assert np.array_equal(h[0], oracle.gap_histogram(np.ones((L, L), np.int8), 0, TC[0], TC[2], 3, 0, 0, 0, 0)) and h.sum() == L * L - 1
    e.close()
