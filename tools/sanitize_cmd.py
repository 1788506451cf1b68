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
# round 2: HALO kernels on a self-connected slab (pair table at 16-row strips, graph replay), the cluster move (label /
# border / flip), cluster labels -> radix sort -> gap histogram and cluster lists, the 2-bit wire format
for L, b, rows in ((512, 0, 16), (256, 1, 8), (128, 0, 0)):
    e = bc.Engine(L, b, 2, 0, 7, slab=(1, 0, None))
    e.ipc_connect(None, None)
    e.set_option("rows_per_thread", rows)
    e.set_option("graph_block", 4)
    e.set_params(0.608, 1.966, 1.0)
    e.init_random()
    obs = e.sweep(12, 5)
    e.cluster_update(1)
    e.sweep(2)
    e.close()
    print("halo", L, b, rows, int(obs["B"].sum()))
for L, b in ((192, 0), (100, 0), (33, 1), (8, 0)):
    e = bc.Engine(L, b, 3, 0, 9)
    for r in range(3):
        e.set_params(0.58 + 0.02 * r, 1.966, 1.0, replica=r)
    e.init_random()
    e.sweep(6)
    e.cluster_update(2)
    mom = e.cluster_moments(1, 3)
    h0 = e.gap_histogram(0, 0, 0)
    h1 = e.gap_histogram(1, 2, 1)
    sites, starts, signs = e.cluster_lists(1, 1, 2)
    p = e.download_packed()
    e.upload_packed(p)
    e.close()
    print("cluster", L, b, int(mom["n_clusters"].sum()), int(h0.sum()), int(h1.sum()), len(signs))
print("done")
