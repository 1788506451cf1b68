"""Golden for the susceptibility table: runs the UNMODIFIED corner_contribution/magnet.cpp (built into
oracle/_ref/magnet) on cluster files written for random lattices, and records its CSV together with the
per-sample magnetisations it derived.  tests/test_magnet.py feeds those magnetisations to bc_magnet.
    python tests/golden/make_magnet_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
TOOL = os.path.join(ROOT, "blume_capel_b200", "host", "bc_gaptool")
MAGNET = os.path.join(ROOT, "oracle", "_ref", "magnet")

rng = np.random.default_rng(20261020)
base = tempfile.mkdtemp(prefix="bcmag_")
lines = []
for L in (48, 64):
    for run in range(100):
        d = os.path.join(base, "in", str(L), str(run))
        os.makedirs(d)
        p_plus = rng.uniform(0.1, 0.6)
        for i in range(3):
            s = rng.choice(np.array([-1, 0, 1], np.int8), size=(L, L), p=[0.7 - p_plus, 0.3, p_plus])
            raw = os.path.join(d, "lat.bin")
            open(raw, "wb").write(s.tobytes())
            subprocess.run([TOOL, "write", str(L), "periodic", raw, os.path.join(d, f"{i}.txt")], check=True)
            os.remove(raw)
            lines.append(f"{L} {run} {int(s.astype(np.int64).sum())}")
out = os.path.join(base, "magnet.txt")
subprocess.run([MAGNET, os.path.join(base, "in"), out], check=True)
open(os.path.join(HERE, "magnet_samples.txt"), "w").write("\n".join(lines) + "\n")
open(os.path.join(HERE, "magnet_expected.csv"), "wb").write(open(out, "rb").read())
print(open(out).read()[:300])
