"""Streaming-kernel variants side by side (development aid): one tie test per row instead of one per 8 sites, and the
pipelined kernel at 7 CTAs / SM.  BC_ENGINE_LIB selects the build."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc

def rate(L, n_rep, opts, sweeps=60, measure_every=0, boundary=0):
    e = bc.Engine(L, boundary, n_rep, 0, 11)
    for k, v in opts.items():
        e.set_option(k, v)
    e.set_params(0.608, 1.966, 1.0)
    e.init_random()
    e.sweep(32)
    e.sync()
    best = 0.0
    for _ in range(3):
        e.timer_start()
        e.sweep(sweeps, measure_every, fetch=False)
        ms = e.timer_stop()
        best = max(best, float(L) * L * n_rep * sweeps / (ms * 1e6))
    plan = e.plan()
    e.close()
    return best, plan

lib = os.path.basename(os.environ.get("BC_ENGINE_LIB", "default"))
cases = [(4096, 64, {}, 0, 0), (4096, 64, {"swp_min_rows": 16}, 0, 0), (4096, 64, {"swp_min_rows": 0}, 0, 0),
         (4096, 64, {}, 1, 0), (4096, 64, {}, 10, 0), (4096, 64, {"rows_per_thread": 32, "swp_min_rows": 16}, 0, 0),
         (4096, 64, {"rows_per_thread": 16}, 0, 0), (4096, 64, {"rows_per_thread": 8}, 0, 0),
         (4096, 64, {"rows_per_thread": 4}, 0, 0), (4096, 64, {}, 0, 1), (8192, 8, {"swp_min_rows": 16}, 0, 0),
         (8192, 8, {}, 0, 0), (1024, 64, {}, 0, 0), (4096, 1, {}, 0, 0)]
for L, n_rep, opts, me, bnd in cases:
    r, plan = rate(L, n_rep, opts, 60 if n_rep > 1 else 400, me, bnd)
    print(json.dumps({"lib": lib, "L": L, "n_rep": n_rep, "opts": opts, "measure_every": me, "boundary": bnd,
                      "updates_per_ns": round(r, 1), "rows": plan["rows_per_thread"], "grid": plan["grid"],
                      "pipelined": plan["pipelined"], "pair": plan.get("pair")}), flush=True)
