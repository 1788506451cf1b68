"""Streaming-kernel timing across strip heights and variants (development aid): plain / measured sweeps, open boundary,
per-site-table kernels (pair_min_rows 0), launch-bound single lattices.  BC_ENGINE_LIB selects the build; a second
argument "table" prints libraries side by side from a jsonl written by earlier runs."""
import json, os, sys
from collections import OrderedDict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 2 and sys.argv[1] == "table":
    t = OrderedDict()
    for r in (json.loads(l) for l in open(sys.argv[2])):
        t.setdefault((r["L"], r["n_rep"], json.dumps(r["opts"]), r["measure_every"], r["boundary"]), {})[r["lib"]] = r["updates_per_ns"]
    for k, v in t.items():
        print(k, {a.replace("libbc_engine", "").replace(".so", ""): b for a, b in v.items()})
    sys.exit(0)

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
cases = [(4096, 64, {}, 0, 0), (4096, 64, {}, 1, 0), (4096, 64, {}, 10, 0), (4096, 64, {}, 0, 1), (4096, 64, {}, 1, 1)]
cases += [(4096, 64, {"rows_per_thread": r}, 0, 0) for r in (128, 64, 16, 8, 4, 2)]
cases += [(4096, 64, {"rows_per_thread": r, "pair_min_rows": 0}, 0, 0) for r in (64, 32, 16)]
cases += [(8192, 8, {}, 0, 0), (8192, 8, {"rows_per_thread": 64}, 0, 0), (8192, 8, {"rows_per_thread": 16}, 0, 0),
          (1024, 64, {}, 0, 0), (4096, 4, {}, 0, 0), (4096, 1, {}, 0, 0), (2048, 1, {}, 0, 0), (8192, 1, {}, 0, 0),
          (4096, 1, {}, 0, 1)]
for L, n_rep, opts, me, bnd in cases:
    r, plan = rate(L, n_rep, opts, 60 if n_rep > 1 else 400, me, bnd)
    print(json.dumps({"lib": lib, "L": L, "n_rep": n_rep, "opts": opts, "measure_every": me, "boundary": bnd,
                      "updates_per_ns": round(r, 1), "rows": plan["rows_per_thread"], "grid": plan["grid"],
                      "pair": plan.get("pair")}), flush=True)
