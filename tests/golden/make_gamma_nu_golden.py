"""Golden for the cluster-size table: runs the UNMODIFIED corner_contribution/gamma_nu.cpp (oracle/_ref/gamma_nu)
on cluster files of random lattices and records its output with the per-sample (sum size^2, max size).
    python tests/golden/make_gamma_nu_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TOOL = os.path.join(ROOT, "blume_capel_b200", "host", "bc_gaptool")
GAMMA = os.path.join(ROOT, "oracle", "_ref", "gamma_nu")

rng = np.random.default_rng(20261021)
base = tempfile.mkdtemp(prefix="bcgam_")
lines = []
for L in (12, 16, 24, 32):
    for run in range(24):
        d = os.path.join(base, "in", str(L), str(run))
        os.makedirs(d)
        p_plus = rng.uniform(0.15, 0.6)
        for i in range(3):
            s = rng.choice(np.array([-1, 0, 1], np.int8), size=(L, L), p=[0.7 - p_plus, 0.3, p_plus])
            raw = os.path.join(d, "lat.bin")
            open(raw, "wb").write(s.tobytes())
            subprocess.run([TOOL, "write", str(L), "periodic", raw, os.path.join(d, f"{i}.txt")], check=True)
            os.remove(raw)
            ssq, mx, _ = oracle.cluster_moments(s, 0, 1.0, 1.0, 0)
            lines.append(f"{L} {run} {ssq} {mx}")
out = os.path.join(base, "gamma.txt")
subprocess.run([GAMMA, os.path.join(base, "in"), out], check=True, stdout=subprocess.DEVNULL)
open(os.path.join(HERE, "gamma_nu_samples.txt"), "w").write("\n".join(lines) + "\n")
open(os.path.join(HERE, "gamma_nu_expected.txt"), "wb").write(open(out, "rb").read())
print(open(out).read())
