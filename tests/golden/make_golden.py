"""Regenerates tests/golden/{tables.json,trajectories.npz} from the CPU oracle.

The reference cannot generate trajectory vectors (non-reproducible mt19937 seeded from
random_device, main.cpp:47-48,307-309), so trajectory fixtures pin the SPEC (DESIGN.md
section 3) as implemented by oracle/bc_oracle.c at the commit that wrote them; the GPU and
any later oracle must reproduce them bit for bit.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

points = [(0.608, 1.966, 1.0), (2.2691853, -1000.0, 1.0), (1.0, 0.5, 0.0), (0.3, 1.0, 1.0), (5.0, -2.0, -1.0)]
tables = [{"T": T, "D": D, "J": J, "thr31": [int(x) for x in oracle.build_table(T, D, J)]} for T, D, J in points]
with open(os.path.join(HERE, "tables.json"), "w") as f:
    json.dump(tables, f, indent=0)

cases = [
    dict(L=4, boundary=0, seed=0x5EED0001, replica=0, n_sweeps=50, T=0.608, D=1.966, J=1.0),
    dict(L=12, boundary=0, seed=0x5EED0001, replica=3, n_sweeps=30, T=0.608, D=1.966, J=1.0),
    dict(L=64, boundary=0, seed=0x5EED0001, replica=0, n_sweeps=25, T=0.608, D=1.966, J=1.0),
    dict(L=64, boundary=1, seed=0x5EED0005, replica=1, n_sweeps=25, T=0.608, D=1.966, J=1.0),
    dict(L=9, boundary=1, seed=7, replica=0, n_sweeps=40, T=1.2, D=0.3, J=1.0),
    dict(L=96, boundary=0, seed=(0xABCD << 32) | 0x1234, replica=2, n_sweeps=12, T=2.2691853, D=-1000.0, J=1.0),
    dict(L=128, boundary=0, seed=0x5EED0002, replica=0, n_sweeps=16, T=0.608, D=1.966, J=1.0),
]
out = {"meta": json.dumps(cases)}
for i, m in enumerate(cases):
    tab = oracle.build_table(m["T"], m["D"], m["J"])
    s = oracle.init_random(m["L"], m["seed"], m["replica"])
    out[f"init_{i}"] = s.copy()
    obs = oracle.sweep(s, tab, m["seed"], m["replica"], 0, m["n_sweeps"], 1, m["boundary"])
    out[f"final_{i}"] = s
    out[f"obs_{i}"] = np.stack([obs["M"], obs["Q"], obs["B"]])
np.savez_compressed(os.path.join(HERE, "trajectories.npz"), **out)
print("wrote", len(tables), "tables and", len(cases), "trajectories")
