"""Regenerates tests/golden/cluster_moves.npz from the CPU oracle: seeded lattices, sweeps interleaved with
Swendsen-Wang cluster moves (bc_oracle_cluster_update) and the cluster-size moments of the result.

Like the trajectory fixtures these pin the SPEC of the cluster move (bond stream 3, flip
stream 4, root = smallest site of a cluster) as implemented at the commit that wrote them: the reference's Wolff move
(main.cpp:113-154) draws from a non-reproducible mt19937 and has no vectors of its own.  Run from the repo root:
    python tests/golden/make_cluster_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    dict(L=8, boundary=0, seed=0xC1A0, replica=0, T=0.608, D=1.966, J=1.0, rounds=4, sweeps=2),
    dict(L=24, boundary=0, seed=0xC1A1, replica=2, T=2.0, D=0.0, J=1.0, rounds=3, sweeps=1),
    dict(L=33, boundary=1, seed=0xC1A2, replica=1, T=0.608, D=1.966, J=1.0, rounds=3, sweeps=2),
    dict(L=64, boundary=0, seed=0xC1A3, replica=0, T=2.2691853, D=-1000.0, J=1.0, rounds=2, sweeps=3),
    dict(L=96, boundary=1, seed=(7 << 40) | 5, replica=3, T=1.2, D=0.5, J=1.0, rounds=2, sweeps=1),
    dict(L=160, boundary=0, seed=0xC1A5, replica=0, T=0.608, D=1.966, J=1.0, rounds=2, sweeps=2),
]


def run_case(m):
    tab = oracle.build_table(m["T"], m["D"], m["J"])
    s = oracle.init_random(m["L"], m["seed"], m["replica"])
    for c in range(m["rounds"]):  # sweeps then one cluster move, like step() of main.cpp:158-163
        oracle.sweep(s, tab, m["seed"], m["replica"], c * m["sweeps"], m["sweeps"], 0, m["boundary"])
        oracle.cluster_update(s, m["T"], m["J"], m["seed"], m["replica"], c, m["boundary"])
    mom = [oracle.cluster_moments(s, kind, m["T"], m["J"], m["seed"], m["replica"], 5, m["boundary"]) for kind in (0, 1)]
    return s, np.array(mom, np.int64)


if __name__ == "__main__":
    out = {"meta": json.dumps(CASES)}
    for i, m in enumerate(CASES):
        s, mom = run_case(m)
        out[f"final_{i}"] = s
        out[f"moments_{i}"] = mom
    np.savez_compressed(os.path.join(HERE, "cluster_moves.npz"), **out)
    print("wrote", len(CASES), "cluster-move cases")
