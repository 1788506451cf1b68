"""Wave quantisation of the streaming kernel (development aid): the bench batch (64 x L=4096 = 4096 CTAs of 64-row strips)
on 6 CTAs/SM (pipelined kernel, 888 slots: 4.61 waves) against 7 CTAs/SM (pair-table kernel without pipelined draws,
1036 slots: 3.95 waves), and the per-GPU share of the L=65536 slab run at 8 GPUs (8 x L=8192 has the same CTA count
and strip shape as 8192 rows x 65536 columns) for several strip heights."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc

def rate(L, n_rep, opts, sweeps=60):
    e = bc.Engine(L, bc.PERIODIC, n_rep, 0, 11)
    for k, v in opts.items():
        e.set_option(k, v)
    e.set_params(0.608, 1.966, 1.0)
    e.init_random()
    e.sweep(32)
    e.sync()
    best = 0.0
    for _ in range(3):
        e.timer_start()
        e.sweep(sweeps)
        ms = e.timer_stop()
        best = max(best, float(L) * L * n_rep * sweeps / (ms * 1e6))
    plan = e.plan()
    e.close()
    return best, plan

for L, n_rep, opts in ((4096, 64, {}), (4096, 64, {"swp_min_rows": 0}), (4096, 64, {"rows_per_thread": 32}),
                       (4096, 64, {"rows_per_thread": 32, "swp_min_rows": 0}), (4096, 64, {"rows_per_thread": 128}),
                       (4096, 56, {}), (4096, 55, {}), (4096, 69, {}),
                       (8192, 8, {"rows_per_thread": 64}), (8192, 8, {"rows_per_thread": 32}), (8192, 8, {"rows_per_thread": 16}),
                       (8192, 8, {"rows_per_thread": 64, "swp_min_rows": 0}), (8192, 8, {"rows_per_thread": 32, "swp_min_rows": 0})):
    r, plan = rate(L, n_rep, opts)
    print(json.dumps({"L": L, "n_rep": n_rep, "opts": opts, "updates_per_ns": round(r, 1), "rows": plan["rows_per_thread"],
                      "grid": plan["grid"], "pipelined": plan["pipelined"], "waves_6": round(plan["grid"] / 888, 2),
                      "waves_7": round(plan["grid"] / 1036, 2)}), flush=True)
