"""Development aid: where does the pair-table kernel differ from the oracle (tie repair)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import blume_capel_b200 as bc
from oracle import oracle
L, n_rep, rows = 1024, 1, 32
params = (0.3, 1.0, 1.0)
seed = 0x9A1 + L + rows
tab = oracle.build_table(*params)
for swp, me, n in ((32, 0, 1), (0, 0, 1), (32, 1, 1), (32, 0, 2), (32, 2, 6)):
    e = bc.Engine(L, 0, n_rep, 0, seed)
    for k, v in (("rows_per_thread", rows), ("resident", 0), ("pair_min_rows", 8), ("swp_min_rows", swp)):
        e.set_option(k, v)
    e.set_params(*params)
    e.init_random()
    ref = oracle.init_random(L, seed, 0)
    e.sweep(n, me)
    got = e.download()[0]
    ties_gpu = e.tie_count
    plan = e.plan()
    e.close()
    oracle.sweep(ref, tab, seed, 0, 0, n, me, 0)
    bad = np.argwhere(got != ref)
    print("swp_min_rows", swp, "measure_every", me, "sweeps", n, "pipelined", plan["pipelined"], "ties", ties_gpu, "differ", len(bad),
          [(int(y), int(x), int(y) % 32, (int(x) // 2) % 16, int(got[y, x]), int(ref[y, x])) for y, x in bad[:12]], flush=True)
