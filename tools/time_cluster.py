"""Timing of bc_cluster_update / bc_cluster_moments (development aid)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc

for L, n_rep, T, D in ((4096, 1, 0.608, 1.966), (1024, 64, 0.608, 1.966), (4096, 4, 2.2691853, -1000.0), (256, 64, 0.608, 1.966)):
    e = bc.Engine(L, bc.PERIODIC, n_rep, 0, 7)
    e.set_params(T, D, 1.0)
    e.init_random()
    e.sweep(50)
    e.cluster_update(1)
    e.sync()
    l0 = e.launch_count
    t0 = time.perf_counter()
    n = 5
    e.cluster_update(n)
    e.sync()
    dt = (time.perf_counter() - t0) / n
    t1 = time.perf_counter()
    e.sweep(10)
    e.sync()
    ds = (time.perf_counter() - t1) / 10
    e.cluster_moments(0)
    t2 = time.perf_counter()
    e.cluster_moments(0)
    dm = time.perf_counter() - t2
    print(json.dumps({"L": L, "n_rep": n_rep, "T": T, "D": D, "ms_per_cluster_update": dt * 1e3, "launches_per_update": (e.launch_count - l0 - 0) / n,
                      "sites_per_ns": L * L * n_rep / (dt * 1e9), "ms_per_sweep": ds * 1e3, "ms_per_cluster_moments": dm * 1e3}), flush=True)
    e.close()
