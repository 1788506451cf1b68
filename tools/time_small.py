"""Sweep time of the reference's own lattice sizes (gen_files.py:11) on the shared-memory-resident kernels vs the
one-thread-per-site kernel (development aid)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc

for L in (8, 12, 16, 24, 32, 48, 64, 96, 128):
    for n_rep in (1, 100):
        row = {"L": L, "n_rep": n_rep}
        for resident in (-1, 0):
            e = bc.Engine(L, bc.PERIODIC, n_rep, 0, 5)
            e.set_option("resident", resident)
            e.set_params(0.608, 1.966, 1.0)
            e.init_random()
            n = 20000 if resident else 2000
            e.sweep(100)
            e.sync()
            t0 = time.perf_counter()
            e.sweep(n, 1, fetch=False)
            e.sync()
            dt = time.perf_counter() - t0
            row["resident" if resident else "per_site"] = {"us_per_sweep": dt / n * 1e6, "updates_per_ns": L * L * n_rep * n / dt * 1e-9,
                                                           "plan": e.plan()["resident"]}
            e.close()
        print(json.dumps(row), flush=True)
