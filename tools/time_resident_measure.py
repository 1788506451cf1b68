import json, os, sys, time
sys.path.insert(0, os.getcwd())
import blume_capel_b200 as bc
for L, n_rep in ((64, 1), (32, 1), (8, 1), (64, 100), (128, 1)):
    row = {"L": L, "n_rep": n_rep}
    for me in (0, 1, 10):
        e = bc.Engine(L, bc.PERIODIC, n_rep, 0, 5)
        e.set_params(0.608, 1.966, 1.0); e.init_random(); e.sweep(100); e.sync()
        n = 50000
        t0 = time.perf_counter(); e.sweep(n, me, fetch=False); e.sync(); dt = time.perf_counter() - t0
        row[f"me{me}_us"] = round(dt / n * 1e6, 3)
        e.close()
    print(json.dumps(row), flush=True)
