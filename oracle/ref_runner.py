"""Runs the UNMODIFIED reference simulator built into oracle/_ref (TEST INFRASTRUCTURE ONLY).

The binaries are compiled from /root/reference/main.cpp by oracle/Makefile with the flags of
job_scripts/burner.sh:11; nothing here reads /root/reference at run time.  CLI contract:
``tri-L run root burn nsteps T D J`` (main.cpp:316-324); with burn=0 it loads
./{root}/burn/{L}_burn.txt (main.cpp:358) and writes ./{root}/spin/{L}/{run}/{i}.txt and
./{root}/fk/{L}/{run}/{i}.txt in gap format (main.cpp:232-257, 369-372).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def binary(L: int) -> str:
    return os.path.join(REF_DIR, f"tri-{L}")


def available(L: int) -> bool:
    return os.access(binary(L), os.X_OK)


def write_gap_file(path: str, spins: np.ndarray):
    """One cluster line per sign holding all sites of that sign -- a valid burn file for
    get_lattice_from_burn (main.cpp:281-299), which only needs sign + delta-coded sites."""
    flat = np.asarray(spins).reshape(-1)
    with open(path, "w") as f:
        for sign, ch in ((1, "+"), (-1, "-")):
            idx = np.flatnonzero(flat == sign)
            if idx.size == 0:
                continue
            deltas = np.diff(np.concatenate(([0], idx)))
            f.write(ch + " " + " ".join(map(str, deltas)) + " \n")
        f.write("\n")


def parse_gap_file(path: str, L: int) -> np.ndarray:
    """Gap format (main.cpp:236-256) -> L x L int8 lattice (unlisted sites are 0)."""
    s = np.zeros(L * L, np.int8)
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok or tok[0] not in "+-":
                continue
            sign = 1 if tok[0] == "+" else -1
            pos = np.cumsum(np.array(tok[1:], dtype=np.int64))
            s[pos] = sign
    return s.reshape(L, L)


def run_reference(L: int, T: float, D: float, J: float, nsteps: int, run: int = 0, start: np.ndarray | None = None,
                  keep: bool = False, workdir: str | None = None):
    """Runs `tri-L run root 0 nsteps T D J` from an all-+ (or given) burn file.
    Returns (list of L x L lattices, wall seconds)."""
    if not available(L):
        raise FileNotFoundError(f"{binary(L)} not built (oracle/Makefile needs /root/reference)")
    base = workdir or tempfile.mkdtemp(prefix="bcref_")
    root = "r"
    for sub in (f"{root}/burn", f"{root}/spin/{L}/{run}", f"{root}/fk/{L}/{run}"):
        os.makedirs(os.path.join(base, sub), exist_ok=True)
    if start is None:
        start = np.ones((L, L), np.int8)
    write_gap_file(os.path.join(base, root, "burn", f"{L}_burn.txt"), start)
    t0 = time.perf_counter()
    subprocess.run([binary(L), str(run), root, "0", str(nsteps), repr(T), repr(D), repr(J)], cwd=base, check=True,
                   stdout=subprocess.DEVNULL)
    wall = time.perf_counter() - t0
    out = [parse_gap_file(os.path.join(base, root, "spin", str(L), str(run), f"{i}.txt"), L) for i in range(nsteps)]
    if not keep and workdir is None:
        shutil.rmtree(base, ignore_errors=True)
    return out, wall


def metropolis_attempts(L: int, nsteps: int) -> int:
    """Metropolis attempts the reference performs for `nsteps` samples: 9N step() calls per
    sample (main.cpp:362), each L*3L attempts (main.cpp:158-161)."""
    N = L * L
    return nsteps * 9 * N * 3 * N
