"""Resident (shared-memory) kernel throughput for the reference's sizes with many replicas, and the latency of a single
replica (development aid).  BC_ENGINE_LIB selects the build."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc
lib = os.path.basename(os.environ.get("BC_ENGINE_LIB", "default"))
for L, n_rep, sweeps, me in ((64, 4096, 200, 0), (64, 4096, 200, 1), (32, 8192, 200, 0), (128, 1024, 200, 0), (48, 4096, 200, 0),
                             (16, 16384, 200, 0), (256, 512, 100, 0), (64, 100, 2000, 1), (64, 1, 20000, 1), (128, 100, 1000, 0)):
    e = bc.Engine(L, bc.PERIODIC, n_rep, 0, 5)
    e.set_option("resident", 1)
    e.set_params(0.608, 1.966, 1.0)
    e.init_random(); e.sweep(20); e.sync()
    best = 0
    for _ in range(3):
        e.timer_start(); e.sweep(sweeps, me, fetch=False); ms = e.timer_stop()
        best = max(best, L * L * n_rep * sweeps / (ms * 1e6))
    print(json.dumps({"lib": lib, "L": L, "n_rep": n_rep, "measure_every": me, "updates_per_ns": round(best, 1),
                      "us_per_sweep": round(L * L * n_rep / best / 1e3, 3)}), flush=True)
    e.close()
