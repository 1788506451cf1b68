"""Measures BASELINE.json configs 1, 3 and 5 on one GPU (configs 2 and 4 are bench.py / tools/bench_slab.py).
Prints one JSON line per config; results are recorded under profiles/."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import blume_capel_b200 as bc

T, D, J = 0.608, 1.966, 1.0
SPLIT = int(os.environ.get("BC_GROUP_SPLIT", "0"))


def config1():
    """L=64 periodic, one (T,D) near the tricritical point, 1e5 sweeps, measured every sweep."""
    L, n = 64, 100_000
    e = bc.Engine(L, bc.PERIODIC, 1, 0, 0x5EED0001)
    e.set_params(T, D, J)
    e.init_random()
    e.sync()
    t0 = time.perf_counter()
    obs = e.sweep(n, 1)
    wall = time.perf_counter() - t0
    ms = e.last_sweep_ms()
    N = L * L
    m = obs[:, 0]["M"][n // 10:] / N
    E = (-J * obs[:, 0]["B"] + D * obs[:, 0]["Q"])[n // 10:] / N
    out = {"config": 1, "workload": "L=64 periodic, (T,D)=(0.608,1.966), 1e5 checkerboard sweeps, observables every sweep, 1 replica",
           "plan": e.plan(), "device_ms": ms, "wall_s": wall, "updates_per_ns": N * n / (ms * 1e6),
           "E_per_site": float(E.mean()), "abs_m": float(np.abs(m).mean()),
           "reference_wall_s_same_attempts": "115 s on 1 core for 1.1e5 sweep-equivalents + Wolff (BASELINE.md section 2)"}
    e.close()
    return out


def config3(sweeps=200, ladder_rows=int(os.environ.get("BC_LADDER_ROWS", "8"))):
    """Finite-size-scaling ladder: L = 32..2048, one GPU's share = 8 (T,D) points x 7 sizes, concurrent engines.
    The engines share the GPU, so each one may use longer strips (fewer, more efficient CTAs) than the plan it
    would pick for itself alone: ladder_rows > 0 sets "rows_per_thread" for the streaming sizes."""
    sizes = [32, 64, 128, 256, 512, 1024, 2048]
    pts = [(0.58 + 0.06 * i / 7, 1.94 + 0.05 * (i % 8) / 7) for i in range(8)]
    engines = []
    for L in sizes:
        e = bc.Engine(L, bc.PERIODIC, len(pts), 0, 0x5EED0003 + L)
        if ladder_rows and L >= 256:
            e.set_option("rows_per_thread", min(ladder_rows, L // 8))
        for r, (t, d) in enumerate(pts):
            e.set_params(t, d, J, replica=r)
        e.init_random()
        e.sweep(4)
        engines.append(e)
    for e in engines:
        e.sync()
    t0 = time.perf_counter()
    for e in engines:  # asynchronous: each engine has its own stream, small sizes overlap with large ones
        e.sweep(sweeps, sweeps, fetch=False)
    for e in engines:
        e.sync()
    wall = time.perf_counter() - t0
    sites = sum(L * L * len(pts) for L in sizes)
    per = {str(e.L): e.last_sweep_ms() for e in engines}
    out = {"config": 3, "workload": "FSS ladder L=32..2048 x 8 (T,D) points per GPU (1/8 of the 64-point grid), per-replica tables",
           "sweeps": sweeps, "ladder_rows": ladder_rows, "sites_per_sweep": sites, "wall_ms": wall * 1e3, "updates_per_ns": sites * sweeps / (wall * 1e9),
           "device_ms_per_size": per, "plans": {str(e.L): e.plan() for e in engines}}
    for e in engines:
        e.close()
    return out


def config5(sweeps=100, n_rep=64):
    """Open-boundary lattices L = 64..1024, 64 replicas per size."""
    res = {}
    tot_sites, tot_ms = 0, 0.0
    for L in (64, 128, 256, 512, 1024):
        e = bc.Engine(L, bc.OPEN, n_rep, 0, 0x5EED0005 + L)
        e.set_params(T, D, J)
        e.init_random()
        e.sweep(4)
        e.sync()
        e.sweep(sweeps, sweeps, fetch=False)
        e.sync()
        ms = e.last_sweep_ms()
        res[str(L)] = {"ms": ms, "updates_per_ns": L * L * n_rep * sweeps / (ms * 1e6), "plan": e.plan()}
        tot_sites += L * L * n_rep
        tot_ms += ms
        e.close()
    return {"config": 5, "workload": f"open boundary, L=64..1024, {n_rep} replicas per size, {sweeps} sweeps, observables at the end",
            "per_size": res, "updates_per_ns_overall": tot_sites * sweeps / (tot_ms * 1e6)}


def config3_group(sweeps=200, rows=int(os.environ.get("BC_GROUP_ROWS", "0"))):
    """The same ladder with the streaming sizes (L >= 128) in ONE launch per half-sweep (bc_sweep_group); the
    shared-memory-resident sizes (L = 32, 64) run on their own streams next to it."""
    sizes = [32, 64, 128, 256, 512, 1024, 2048]
    pts = [(0.58 + 0.06 * i / 7, 1.94 + 0.05 * (i % 8) / 7) for i in range(8)]
    engines = []
    for L in sizes:
        e = bc.Engine(L, bc.PERIODIC, len(pts), 0, 0x5EED0003 + L)
        for r, (t, d) in enumerate(pts):
            e.set_params(t, d, J, replica=r)
        e.init_random()
        e.sweep(4)
        engines.append(e)
    small = [e for e in engines if e.L < 128]
    group = sorted((e for e in engines if e.L >= 128), key=lambda e: -e.L)  # big lattices' CTAs first
    if rows:
        group[0].set_option("rows_per_thread", rows)
    chains = [group[:1], group[1:]] if SPLIT else [group]  # two chains: one fills the other's launch tails
    for c in chains:
        bc.sweep_group(c, 16)  # builds the graph
    for e in engines:
        e.sync()
    t0 = time.perf_counter()
    for e in small:
        e.sweep(sweeps, sweeps, fetch=False)
    for c in chains:
        bc.sweep_group(c, sweeps, sweeps, fetch=False)
    for e in engines:
        e.sync()
    wall = time.perf_counter() - t0
    sites = sum(L * L * len(pts) for L in sizes)
    gsites = sum(e.L * e.L * len(pts) for e in group)
    gms = max(c[0].last_sweep_ms() for c in chains)
    out = {"config": 3, "mode": "group launch (bc_sweep_group) for L >= 128", "sweeps": sweeps, "rows_option": rows, "chains": len(chains),
           "sites_per_sweep": sites, "wall_ms": wall * 1e3, "updates_per_ns": sites * sweeps / (wall * 1e9),
           "group_device_ms": gms, "group_updates_per_ns": gsites * sweeps / (gms * 1e6),
           "small_device_ms": {str(e.L): e.last_sweep_ms() for e in small}}
    for e in engines:
        e.close()
    return out


def config5_group(sweeps=100, n_rep=64, rows=int(os.environ.get("BC_GROUP_ROWS", "0"))):
    """Open-boundary series with L = 128..1024 in one launch per half-sweep; L = 64 on the resident kernel."""
    engines = []
    for L in (64, 128, 256, 512, 1024):
        e = bc.Engine(L, bc.OPEN, n_rep, 0, 0x5EED0005 + L)
        e.set_params(T, D, J)
        e.init_random()
        e.sweep(4)
        engines.append(e)
    small, group = engines[:1], engines[:0:-1]
    if rows:
        group[0].set_option("rows_per_thread", rows)
    chains = [group[:1], group[1:]] if SPLIT else [group]
    for c in chains:
        bc.sweep_group(c, 16)
    for e in engines:
        e.sync()
    t0 = time.perf_counter()
    small[0].sweep(sweeps, sweeps, fetch=False)
    for c in chains:
        bc.sweep_group(c, sweeps, sweeps, fetch=False)
    for e in engines:
        e.sync()
    wall = time.perf_counter() - t0
    sites = sum(e.L * e.L * n_rep for e in engines)
    gsites = sum(e.L * e.L * n_rep for e in group)
    gms = max(c[0].last_sweep_ms() for c in chains)
    out = {"config": 5, "mode": "group launch (bc_sweep_group) for L >= 128", "sweeps": sweeps, "rows_option": rows, "chains": len(chains),
           "wall_ms": wall * 1e3, "updates_per_ns_overall": sites * sweeps / (wall * 1e9),
           "group_device_ms": gms, "group_updates_per_ns": gsites * sweeps / (gms * 1e6), "L64_device_ms": small[0].last_sweep_ms()}
    for e in engines:
        e.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "3", "5"]
    for w in which:
        print(json.dumps({"1": config1, "3": config3, "5": config5, "3g": config3_group, "5g": config5_group}[w]()), flush=True)
