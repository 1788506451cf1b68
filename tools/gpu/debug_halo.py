"""Development aid: where does a one-rank fused-halo slab run differ from the oracle?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import blume_capel_b200 as bc
from oracle import oracle
TC = (0.608, 1.966, 1.0)
L, n_rep, rows, boundary = 4096, 1, 64, 0
seed = 0x4A10
for slab in (True, False):
    for n in (1, 2):
        e = bc.Engine(L, boundary, n_rep, 0, seed, slab=(1, 0, None)) if slab else bc.Engine(L, boundary, n_rep, 0, seed)
        if slab: e.ipc_connect(None, None)
        e.set_option("rows_per_thread", rows)
        e.set_params(*TC)
        e.init_random()
        e.sweep(n, 0)
        got = e.download()
        plan = e.plan()
        e.close()
        tab = oracle.build_table(*TC)
        ref = oracle.init_random(L, seed, 0)
        oracle.sweep(ref, tab, seed, 0, 0, n, 0, boundary)
        bad = np.argwhere(got[0] != ref)
        print("slab", slab, "sweeps", n, "plan", plan, "differ", len(bad), bad[:20].tolist(), flush=True)
