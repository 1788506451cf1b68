"""Quick device-side timing of bc_sweep variants (development aid, not the bench).
usage: python tools/tune.py [quick|full]   (BC_ENGINE_LIB selects an alternative build)"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blume_capel_b200 as bc

TC = (0.608, 1.966, 1.0)

def run(L, n_rep, rows, sweeps, measure_every=0, boundary=0, opts=()):
    e = bc.Engine(L, boundary, n_rep, 0, 1)
    e.set_option("rows_per_thread", rows)
    for k, v in opts:
        e.set_option(k, v)
    e.set_params(*TC)
    e.init_random()
    e.sweep(5, 0)
    e.sync()
    best = 1e30
    for rep in range(3):
        e.sweep(sweeps, measure_every, fetch=False)
        e.sync()
        best = min(best, e.last_sweep_ms())
    plan = e.plan()
    e.close()
    ups = L * L * n_rep * sweeps / (best * 1e6)
    return ups, best, plan

if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    tag = os.environ.get("BC_ENGINE_LIB", "default")
    cases = []
    if mode == "quick":
        for L, nr, sw in ((4096, 1, 400), (2048, 8, 200), (1024, 8, 400), (512, 8, 800), (8192, 1, 100), (2048, 1, 800)):
            cases.append((L, nr, 0, sw, 0, 0, (("wave", 0),)))
            for rows in (4, 8, 16):
                cases.append((L, nr, rows, sw, 0, 0, (("wave", 1),)))
        cases.append((4096, 64, 8, 10, 0, 0))
        cases.append((4096, 64, 32, 10, 0, 0))
        cases.append((4096, 64, 0, 10, 1, 0))
        cases.append((4096, 1, 4, 200, 0, 0))
    else:
        for n_rep, sweeps in ((1, 200), (64, 10)):
            for rows in (2, 4, 8, 16, 32, 64):
                cases.append((4096, n_rep, rows, sweeps, 0, 0))
        cases.append((4096, 64, 0, 10, 1))
        cases.append((4096, 64, 0, 10, 0, 1))
        cases.append((4096, 1, 0, 200, 1))
        cases.append((1024, 64, 0, 50, 0))
        cases.append((256, 512, 0, 50, 0))
        for L, nr in ((64, 4096), (32, 8192), (128, 1024), (64, 100), (64, 1), (128, 100), (256, 256)):
            cases.append((L, nr, 0, 2000 if nr <= 100 else 50, 0, 0, (("resident", 0),)))
            cases.append((L, nr, 0, 2000 if nr <= 100 else 50, 0, 0, (("resident", 1),)))
        cases.append((64, 1, 0, 20000, 1, 0, (("resident", 1),)))
        cases.append((64, 100, 0, 2000, 1, 0, (("resident", 1),)))
    for c in cases:
        ups, ms, plan = run(*c)
        print(json.dumps({"lib": os.path.basename(tag), "L": c[0], "n_rep": c[1], "rows": c[2], "sweeps": c[3],
                          "measure_every": c[4] if len(c) > 4 else 0, "opts": c[6] if len(c) > 6 else (),
                          "updates_per_ns": round(ups, 1), "ms": round(ms, 3), "plan": plan}), flush=True)
