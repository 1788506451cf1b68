"""Golden for the gap-size histogram: runs the UNMODIFIED corner_contribution/gap_size_statistics.cpp
(oracle/_ref/gap_size_statistics) on cluster files of random lattices and keeps, for a few runs, the input
cluster files' lattices and the tool's per-run output.
    python tests/golden/make_gap_stats_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.dirname(os.path.abspath(__file__))
TOOL = os.path.join(ROOT, "blume_capel_b200", "host", "bc_gaptool")
REF = os.path.join(ROOT, "oracle", "_ref", "gap_size_statistics")
KEEP = 4  # runs per size kept as fixtures

rng = np.random.default_rng(20261022)
base = tempfile.mkdtemp(prefix="bcgap_")
lattices = {}
for L in (48, 64):
    os.makedirs(os.path.join(base, "out", str(L)))
    for run in range(100):
        d = os.path.join(base, "in", str(L), str(run))
        os.makedirs(d)
        p_plus = rng.uniform(0.15, 0.6)
        for i in range(3):
            s = rng.choice(np.array([-1, 0, 1], np.int8), size=(L, L), p=[0.7 - p_plus, 0.3, p_plus])
            raw = os.path.join(d, "lat.bin")
            open(raw, "wb").write(s.tobytes())
            subprocess.run([TOOL, "write", str(L), "periodic", raw, os.path.join(d, f"{i}.txt")], check=True)
            os.remove(raw)
            if run < KEEP:
                lattices[f"L{L}_run{run}_s{i}"] = s
subprocess.run([REF, os.path.join(base, "in"), os.path.join(base, "out")], check=True)
out = {}
for L in (48, 64):
    for run in range(KEEP):
        out[f"L{L}_run{run}"] = open(os.path.join(base, "out", str(L), f"{run}.txt")).read()
np.savez_compressed(os.path.join(HERE, "gap_stats_lattices.npz"), **lattices)
import json
json.dump(out, open(os.path.join(HERE, "gap_stats_expected.json"), "w"), indent=0)
print(out["L48_run0"][:200])
