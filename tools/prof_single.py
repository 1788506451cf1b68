"""Launch-bound case for the record: one L=4096 lattice (and 2048, 8192), wall time per half-sweep under CUDA-graph
replay + programmatic dependent launch (run without ncu) -- to be set against the kernel's own duration from the ncu
launch list of the same command (gpu__time_duration, serialised, no overlap): the difference is the dependent-launch
floor that temporal blocking would have to remove."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc
for L in (int(a) for a in (sys.argv[1:] or ["4096", "2048", "8192"])):
    e = bc.Engine(L, bc.PERIODIC, 1, 0, 5)
    e.set_params(0.608, 1.966, 1.0)
    e.init_random()
    e.sweep(64)
    e.sync()
    best = 1e30
    for _ in range(3):
        e.timer_start(); e.sweep(400); best = min(best, e.timer_stop())
    plan = e.plan()
    print(json.dumps({"L": L, "us_per_half_sweep": round(best * 1e3 / 800, 3), "updates_per_ns": round(L * L * 400 / (best * 1e6), 1),
                      "rows": plan["rows_per_thread"], "grid": plan["grid"], "pair": plan["pair"]}), flush=True)
    e.close()
