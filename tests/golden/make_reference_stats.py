"""Generates tests/golden/reference_stats.json by running the UNMODIFIED reference
(/root/reference/main.cpp compiled into oracle/_ref by oracle/Makefile) in this container.

Procedure (statistical tier): all-'+' burn file, `tri-L run root 0 nsamp T D J`
for 8 run indices, first 2 samples of each run dropped, observables recomputed from the spin/
gap-format dumps (listed +-1 sites, zeros elsewhere), jackknife over the 8 runs.
    python tests/golden/make_reference_stats.py [L ...]
"""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle, ref_runner  # noqa: E402

T, D, J = 0.608, 1.966, 1.0
PLAN = {4: 2500, 8: 500, 16: 250, 32: 40, 64: 6}  # samples per run
RUNS = 8


def one_run(args):
    L, nsamp, run = args
    lat, wall = ref_runner.run_reference(L, T, D, J, nsamp, run=run)
    obs = np.array([oracle.observables(s) for s in lat[2:]], dtype=np.float64)
    return obs, wall


def stats(obs_runs, L):
    N = L * L

    def est(runs):
        o = np.concatenate(runs)
        m = o[:, 0] / N
        q = o[:, 1] / N
        E = (-J * o[:, 2] + D * o[:, 1]) / N
        m2, m4 = (m**2).mean(), (m**4).mean()
        return np.array([E.mean(), np.abs(m).mean(), q.mean(), 1 - m4 / (3 * m2 * m2), N * (m2 - np.abs(m).mean() ** 2)])

    full = est(obs_runs)
    jk = np.array([est(obs_runs[:i] + obs_runs[i + 1:]) for i in range(len(obs_runs))])
    n = len(obs_runs)
    err = np.sqrt((n - 1) / n * ((jk - jk.mean(0)) ** 2).sum(0))
    return full, err


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or sorted(PLAN)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_stats.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    with ProcessPoolExecutor(max_workers=min(RUNS, os.cpu_count() or 1)) as ex:
        for L in sizes:
            res = list(ex.map(one_run, [(L, PLAN[L], r) for r in range(RUNS)]))
            full, err = stats([r[0] for r in res], L)
            out[str(L)] = {"T": T, "D": D, "J": J, "runs": RUNS, "samples_per_run": PLAN[L], "dropped": 2,
                           "names": ["E", "absm", "q", "U4", "chi"], "mean": full.tolist(), "err": err.tolist(),
                           "wall_s_per_run": float(np.mean([r[1] for r in res]))}
            print(L, out[str(L)], flush=True)
            json.dump(out, open(path, "w"), indent=1)
