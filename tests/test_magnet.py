"""f-2: the reference's susceptibility table (corner_contribution/magnet.cpp) from fused magnetisations."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TOOL = os.path.join(ROOT, "blume_capel_b200", "host", "bc_magnet")


def test_table_is_byte_identical_to_the_reference_tool(tmp_path):
    """magnet_expected.csv was written by the unmodified magnet.cpp from cluster files of the same samples
    (tests/golden/make_magnet_golden.py); bc_magnet must print the same bytes from the magnetisations."""
    if not os.access(TOOL, os.X_OK):
        pytest.skip("bc_magnet not built (needs libbc_engine.so)")
    out = tmp_path / "m.csv"
    subprocess.run([TOOL, str(out), "--from-m", os.path.join(GOLD, "magnet_samples.txt")], check=True)
    assert out.read_bytes() == open(os.path.join(GOLD, "magnet_expected.csv"), "rb").read()


@pytest.mark.gpu
def test_bc_magnet_on_gpu_is_byte_identical_to_the_reference_tool(tmp_path):
    """bc_magnet's table from the fused magnetisations of a GPU run == the table the UNMODIFIED magnet binary
    (oracle/_ref/magnet, built from corner_contribution/magnet.cpp) prints from the cluster files of the very same
    samples: the run is repeated through the Python binding with bc_magnet's seeds and sweep schedule
    (bc_magnet.cpp: seed + L, burn, counter aligned to the sample spacing), every sample of every run is written
    as a spin/ cluster file, and the reference tool reads them (its hard-wired 100 runs of L = 48 and 64)."""
    import numpy as np
    import blume_capel_b200 as bc
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "magnet")
    gaptool = os.path.join(ROOT, "blume_capel_b200", "host", "bc_gaptool")
    if not os.access(ref_tool, os.X_OK):
        pytest.skip("oracle/_ref/magnet not built")
    runs, samples, burn, sps, seed = 100, 3, 60, 20, 5
    out = tmp_path / "chi.csv"
    subprocess.run([TOOL, str(out), "0.608", "1.966", "1.0", "--L", "48,64", "--runs", str(runs), "--samples", str(samples),
                    "--burn-sweeps", str(burn), "--sweeps-per-sample", str(sps), "--seed", str(seed)], check=True)
    inp = tmp_path / "in"
    for L in (48, 64):
        e = bc.Engine(L, 0, runs, 0, seed + L)
        e.set_params(0.608, 1.966, 1.0)
        e.init_random()
        e.sweep(burn)
        e.sweep_counter = (burn + sps - 1) // sps * sps
        for k in range(samples):
            e.sweep(sps)
            s = e.download()
            for r in range(runs):
                d = inp / str(L) / str(r)
                d.mkdir(parents=True, exist_ok=True)
                raw = tmp_path / "l.bin"
                raw.write_bytes(s[r].tobytes())
                subprocess.run([gaptool, "write", str(L), "periodic", str(raw), str(d / f"{k}.txt")], check=True)
        e.close()
    want = tmp_path / "ref.csv"
    subprocess.run([ref_tool, str(inp), str(want)], check=True)
    rows = out.read_text().strip().splitlines()
    assert rows[0] == "batch,L,magnetic_susceptibility,standard_error" and len(rows) == 1 + 2 * 10
    assert out.read_bytes() == want.read_bytes()


GAMMA_TOOL = os.path.join(ROOT, "blume_capel_b200", "host", "bc_gamma_nu")


def test_gamma_nu_table_is_byte_identical_to_the_reference_tool(tmp_path):
    """gamma_nu_expected.txt was written by the unmodified gamma_nu.cpp from cluster files; bc_gamma_nu prints the
    same bytes from (sum size^2, max size) per sample -- including the reference's +1 from the blank last line."""
    if not os.access(GAMMA_TOOL, os.X_OK):
        pytest.skip("bc_gamma_nu not built")
    out = tmp_path / "g.txt"
    subprocess.run([GAMMA_TOOL, str(out), "--from-s", os.path.join(GOLD, "gamma_nu_samples.txt")], check=True)
    assert out.read_bytes() == open(os.path.join(GOLD, "gamma_nu_expected.txt"), "rb").read()


@pytest.mark.gpu
def test_cluster_moments_on_gpu_match_oracle(tmp_path):
    import numpy as np
    import blume_capel_b200 as bc
    from oracle import oracle
    rng = np.random.default_rng(6)
    for L, b in ((12, 0), (32, 0), (64, 1), (100, 0)):
        s = rng.integers(-1, 2, size=(3, L, L), dtype=np.int8)
        e = bc.Engine(L, b, 3, 0, 321)
        e.set_params(0.608, 1.966, 1.0)
        e.upload(s)
        for kind in (0, 1):
            got = e.cluster_moments(kind, mstep=7)
            for r in range(3):
                want = oracle.cluster_moments(s[r], kind, 0.608, 1.0, 321, r, 7, b)
                assert (got[r]["sum_size_sq"], got[r]["max_size"], got[r]["n_clusters"]) == want, (L, b, kind, r)
        assert np.array_equal(e.download(), s)  # measuring does not change the lattice
        e.close()


@pytest.mark.gpu
def test_bc_gamma_nu_on_gpu(tmp_path):
    out = tmp_path / "g.txt"
    subprocess.run([GAMMA_TOOL, str(out), "0.608", "1.966", "1.0", "--L", "12,32", "--runs", "8", "--samples", "10",
                    "--burn-sweeps", "300", "--kind", "fk", "--seed", "3"], check=True)
    rows = out.read_text().strip().splitlines()
    assert rows[0] == "L Chi SE" and len(rows) == 3
    for r in rows[1:]:
        L, chi, se = r.split(" ")
        assert int(L) in (12, 32) and float(chi) >= 0 and float(se) >= 0


def test_gap_histogram_is_byte_identical_to_the_reference_tool(tmp_path):
    """gap_stats_expected.json holds per-run outputs of the unmodified gap_size_statistics.cpp; bc_gap_stats
    must reproduce them from cluster files that bc_gaptool writes for the same lattices."""
    import json
    import numpy as np
    tool = os.path.join(ROOT, "blume_capel_b200", "host", "bc_gap_stats")
    gaptool = os.path.join(ROOT, "blume_capel_b200", "host", "bc_gaptool")
    if not (os.access(tool, os.X_OK) and os.access(gaptool, os.X_OK)):
        pytest.skip("host tools not built")
    lat = np.load(os.path.join(GOLD, "gap_stats_lattices.npz"))
    want = json.load(open(os.path.join(GOLD, "gap_stats_expected.json")))
    for key, expected in want.items():
        L = int(key[1:3])
        files = []
        for i in range(3):
            raw = tmp_path / f"{key}_{i}.bin"
            raw.write_bytes(lat[f"{key}_s{i}"].astype(np.int8).tobytes())
            txt = tmp_path / f"{key}_{i}.txt"
            subprocess.run([gaptool, "write", str(L), "periodic", str(raw), str(txt)], check=True)
            files.append(str(txt))
        out = tmp_path / f"{key}.out"
        subprocess.run([tool, "hist", str(out), str(L)] + files, check=True)
        assert out.read_text() == expected, key
