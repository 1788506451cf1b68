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


BC_RUNNER = os.path.join(ROOT, "blume_capel_b200", "host", "bc_runner")


def _tree(base):
    out = {}
    for d, _, files in os.walk(base):
        for f in files:
            p = os.path.join(d, f)
            out[os.path.relpath(p, base)] = open(p, "rb").read()
    return out


@pytest.mark.parametrize("cluster_every", [0, 3])
def test_bc_runner_equals_one_bc_run_per_size_and_run(tmp_path, cluster_every):
    """The whole campaign of job_scripts/runner.sh in one process: every (size, run) must produce the files
    bc_run r ... --L L --seed (seed ^ (r << 32)) produces."""
    seed = 0x1234ABCD5678
    flags = ["--burn-sweeps", "30", "--sweeps-per-sample", "8", "--mkdir"]
    if cluster_every:
        flags += ["--cluster-every", str(cluster_every)]
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    r = subprocess.run([BC_RUNNER, "data", "0.608", "1.966", "1.0", "--L", "8,12,32", "--runs", "3,2,2", "--nsamples", "2",
                        "--seed", str(seed), "--obs"] + flags, cwd=a, check=True, capture_output=True, text=True)
    lines = [json.loads(x) for x in r.stdout.strip().splitlines()]
    assert [x["L"] for x in lines] == [8, 12, 32] and [x["runs"] for x in lines] == [3, 2, 2]
    for L, runs in ((8, 3), (12, 2), (32, 2)):
        for run in range(runs):
            subprocess.run([BC_RUN, str(run), "data", "2", "2", "0.608", "1.966", "1.0", "--L", str(L), "--seed",
                            str(seed ^ (run << 32))] + flags, cwd=b, check=True, capture_output=True)
    ta, tb = _tree(a / "data" / "spin"), _tree(b / "data" / "spin")
    assert sorted(ta) == sorted(tb) and len(ta) == (3 + 2 + 2) * 2
    assert ta == tb
    assert _tree(a / "data" / "fk") == _tree(b / "data" / "fk")
    # the burn file is run 0's lattice; bc_run wrote it last for its last run, so compare with a fresh run 0
    rows = open(a / "data" / "obs" / "32.csv").read().strip().splitlines()
    assert rows[0] == "run,sample,sweep,M,Q,B" and len(rows) == 1 + 2 * 2


def test_bc_runner_lockstep_and_burn_file_start(tmp_path):
    """--burn 1 writes the burn files, --burn 0 starts every run from them; --lockstep (one launch per half-sweep for
    all sizes) gives the same files as one thread per size."""
    common = ["data", "0.608", "1.966", "1.0", "--L", "128,256", "--runs", "2", "--seed", "77", "--burn-sweeps", "12",
              "--sweeps-per-sample", "20", "--no-fk", "--mkdir"]
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    for d in (a, b):
        subprocess.run([BC_RUNNER] + common + ["--burn", "1", "--nsamples", "0"], cwd=d, check=True, capture_output=True)
        assert os.path.exists(d / "data" / "burn" / "128_burn.txt") and not os.path.exists(d / "data" / "spin" / "128" / "0" / "0.txt")
    subprocess.run([BC_RUNNER] + common + ["--burn", "0", "--nsamples", "3"], cwd=a, check=True, capture_output=True)
    subprocess.run([BC_RUNNER] + common + ["--burn", "0", "--nsamples", "3", "--lockstep"], cwd=b, check=True, capture_output=True)
    ta, tb = _tree(a / "data"), _tree(b / "data")
    assert len(ta) == 2 + 2 * 2 * 3 and ta == tb
    # same start, different keys: the two runs of a size differ after the first sample
    assert ta["spin/128/0/0.txt"] != ta["spin/128/1/0.txt"]


def test_bc_runner_argument_errors(tmp_path):
    r = subprocess.run([BC_RUNNER, "d", "0.6", "1.9", "1.0", "--L", "64,128", "--sweeps-per-sample", "4", "--lockstep"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 2 and "lockstep" in r.stderr
    r = subprocess.run([BC_RUNNER, "d", "0.6", "1.9", "1.0", "--L", "8,12", "--runs", "1,2,3"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 2
    r = subprocess.run([BC_RUNNER, "d", "0.6", "1.9", "1.0", "--L", "8", "--runs", "1", "--nsamples", "1", "--burn", "0"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "burn file" in r.stderr


def test_bc_runner_grid_of_points_and_shards(tmp_path):
    """--T-list / --D-list: every (T, D) point gets its own tree, replica = point * runs + run; with 2 runs per point
    the runs of point k are the runs of a single-point campaign with seed ^ ((2 k) << 32).  --shard k/n splits the
    points over processes (one per GPU)."""
    seed = 99
    base = ["--L", "8,32", "--runs", "2", "--nsamples", "2", "--burn-sweeps", "20", "--sweeps-per-sample", "6", "--mkdir", "--obs"]
    a, b, c = tmp_path / "a", tmp_path / "b", tmp_path / "c"
    for d in (a, b, c):
        d.mkdir()
    grid = ["--T-list", "0.6,0.7", "--D-list", "1.9,1.95"]
    r = subprocess.run([BC_RUNNER, "data", "0", "0", "1.0", "--seed", str(seed)] + grid + base, cwd=a, check=True,
                       capture_output=True, text=True)
    lines = [json.loads(x) for x in r.stdout.strip().splitlines()]
    assert [(x["L"], x["point"]) for x in lines] == [(8, 0), (8, 1), (8, 2), (8, 3), (32, 0), (32, 1), (32, 2), (32, 3)]
    assert (lines[1]["T"], lines[1]["D"]) == (0.6, 1.95) and (lines[2]["T"], lines[2]["D"]) == (0.7, 1.9)
    # point 2 = (0.7, 1.9): the same files as a single-point campaign whose replicas 0, 1 carry the keys of replicas 4, 5
    subprocess.run([BC_RUNNER, "data", "0.7", "1.9", "1.0", "--seed", str(seed ^ (4 << 32))] + base, cwd=b, check=True,
                   capture_output=True)
    assert _tree(a / "data" / "p2") == _tree(b / "data")
    # two shards (one process per GPU) cover the same points and files between them; the Philox keys follow a process's
    # own replica numbering, so a sharded campaign is reproducible for the same --shard, not identical to the unsharded one
    for k in (0, 1):
        r = subprocess.run([BC_RUNNER, "data", "0", "0", "1.0", "--seed", str(seed), "--shard", f"{k}/2"] + grid + base,
                           cwd=c, check=True, capture_output=True, text=True)
        assert sorted({json.loads(x)["point"] for x in r.stdout.strip().splitlines()}) == [k, k + 2]
    assert sorted(_tree(a / "data")) == sorted(_tree(c / "data"))
    # shard 0 starts with point 0 at replicas 0, 1 like the unsharded run: those files are identical
    assert _tree(a / "data" / "p0") == _tree(c / "data" / "p0")
