"""Small fixed workload for compute-sanitizer: every kernel variant once (streaming, resident, generic,
open/periodic, measured, tie path via many sweeps, slab self-halo)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blume_capel_b200 as bc
for L, b, nrep, opts, slab in ((256, 0, 2, (("resident", 0),), None), (256, 1, 1, (("resident", 0),), None), (64, 0, 3, (), None),
                               (32, 1, 5, (), None), (24, 0, 2, (), None), (9, 1, 2, (), None), (128, 0, 1, (("resident", 0),), (1, 0, None))):
    e = bc.Engine(L, b, nrep, 0, 5, slab=slab)
    for k, v in opts:
        e.set_option(k, v)
    e.set_params(0.608, 1.966, 1.0)
    e.init_random()
    obs = e.sweep(40, 4)
    s = e.download()
    m = e.measure()
    e.upload(s)
    e.close()
    print(L, b, nrep, int(obs["B"].sum()))
print("done")
