"""The C++ host driver (reference CLI + gap-format files) on a real GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BC_RUN = os.path.join(ROOT, "blume_capel_b200", "host", "bc_run")


def test_bc_run_reference_cli_roundtrip(tmp_path, oracle, bc):
    from oracle import ref_runner
    L, seed = 32, 12345
    args = ["--L", str(L), "--seed", str(seed), "--burn-sweeps", "50", "--sweeps-per-sample", "20", "--mkdir",
            "--obs", str(tmp_path / "obs.csv")]
    # run root burn nsteps T D J  (main.cpp:316-324); burn=2: burn in then collect
    r = subprocess.run([BC_RUN, "3", "data", "2", "4", "0.608", "1.966", "1.0"] + args, cwd=tmp_path, check=True,
                       capture_output=True, text=True)
    assert "burning 32" in r.stdout and "burnt 32" in r.stdout  # main.cpp:338,351
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    assert summary["samples"] == 4
    burn = ref_runner.parse_gap_file(tmp_path / "data" / "burn" / "32_burn.txt", L)
    # the burn file is the engine's lattice after init + 50 sweeps: reproduce through the Python binding
    e = bc.Engine(L, 0, 1, 0, seed)
    e.set_params(0.608, 1.966, 1.0)
    e.init_random()
    e.sweep(50)
    assert np.array_equal(e.download(0), burn)
    e.close()
    # collection restarts from the burn file with the sweep counter continuing at 50
    ref = burn.copy()
    tab = oracle.build_table(0.608, 1.966, 1.0)
    rows = open(tmp_path / "obs.csv").read().strip().splitlines()[1:]
    for i in range(4):
        oracle.sweep(ref, tab, seed, 0, 50 + 20 * i, 20, 0)
        got = ref_runner.parse_gap_file(tmp_path / "data" / "spin" / "32" / "3" / f"{i}.txt", L)
        assert np.array_equal(got, ref), i
        fk = ref_runner.parse_gap_file(tmp_path / "data" / "fk" / "32" / "3" / f"{i}.txt", L)
        assert np.array_equal(fk, ref)
        M, Q, B = oracle.observables(ref)
        f = rows[i].split(",")
        assert (int(f[2]), int(f[3]), int(f[4])) == (M, Q, B)


def test_bc_run_burn_only_and_resume(tmp_path):
    args = ["--L", "64", "--seed", "7", "--burn-sweeps", "10", "--sweeps-per-sample", "5", "--mkdir", "--no-fk"]
    subprocess.run([BC_RUN, "0", "d", "1", "0", "0.608", "1.966", "1.0"] + args, cwd=tmp_path, check=True,
                   capture_output=True)
    assert os.path.exists(tmp_path / "d" / "burn" / "64_burn.txt")
    assert not os.path.exists(tmp_path / "d" / "spin" / "64" / "0" / "0.txt")
    subprocess.run([BC_RUN, "0", "d", "0", "2", "0.608", "1.966", "1.0"] + args, cwd=tmp_path, check=True,
                   capture_output=True)
    assert os.path.exists(tmp_path / "d" / "spin" / "64" / "0" / "1.txt")
    assert not os.path.exists(tmp_path / "d" / "fk" / "64" / "0" / "0.txt")


def test_bc_run_missing_burn_file_is_an_error(tmp_path):
    r = subprocess.run([BC_RUN, "0", "nowhere", "0", "1", "0.608", "1.966", "1.0", "--L", "32"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode != 0 and "burn file" in r.stderr


def test_bc_run_exact_checkpoint(tmp_path):
    """(lattice, seed, sweep counter, cluster counter) restores a run exactly: 6 samples in one go equal
    3 samples + --save-state, then --load-state + the remaining 3 (cluster moves interleaved)."""
    common = ["--L", "64", "--seed", "99", "--burn-sweeps", "30", "--sweeps-per-sample", "12", "--cluster-every", "4",
              "--mkdir", "--no-fk"]
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    subprocess.run([BC_RUN, "0", "d", "2", "6", "0.608", "1.966", "1.0"] + common, cwd=a, check=True, capture_output=True)
    subprocess.run([BC_RUN, "0", "d", "2", "3", "0.608", "1.966", "1.0"] + common + ["--save-state", "state.bin"], cwd=b,
                   check=True, capture_output=True)
    subprocess.run([BC_RUN, "0", "d", "0", "6", "0.608", "1.966", "1.0"] + common + ["--load-state", "state.bin"], cwd=b,
                   check=True, capture_output=True)
    for i in range(6):
        fa = (a / "d" / "spin" / "64" / "0" / f"{i}.txt").read_bytes()
        fb = (b / "d" / "spin" / "64" / "0" / f"{i}.txt").read_bytes()
        assert fa == fb, i
