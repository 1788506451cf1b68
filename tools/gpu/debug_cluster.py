"""Development aid: where does the GPU cluster move differ from the oracle?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import blume_capel_b200 as bc
from oracle import oracle

L, T, D = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 0.5, 1.0
n_rep, seed = 2, 0xC1 + L
for trial in range(3):
    e = bc.Engine(L, 0, n_rep, 0, seed)
    e.set_params(T, D, 1.0)
    e.init_random()
    ref = np.stack([oracle.init_random(L, seed, r) for r in range(n_rep)])
    tab = oracle.build_table(T, D, 1.0)
    for step in range(3):
        e.sweep(2)
        for r in range(n_rep):
            oracle.sweep(ref[r], tab, seed, r, 2 * step, 2, 0, 0)
        pre = e.download()
        assert np.array_equal(pre, ref), "sweep mismatch"
        e.cluster_update(1)
        got = e.download()
        for r in range(n_rep):
            lab = oracle.cluster_labels(ref[r], 1, T, 1.0, seed, r, 0, 0)  # (labels at another counter: only for sizes)
            oracle.cluster_update(ref[r], T, 1.0, seed, r, step, 0)
            bad = np.argwhere(got[r] != ref[r])
            print(f"trial {trial} step {step} rep {r}: {len(bad)} sites differ", end="")
            if len(bad):
                ys, xs = bad[:, 0], bad[:, 1]
                print(f"  y {ys.min()}..{ys.max()} x {xs.min()}..{xs.max()} first {bad[:5].tolist()} tiles {sorted(set((int(y)//64, int(x)//64) for y, x in bad))[:12]}", end="")
            print()
        ref = got.copy()  # continue from the GPU's state
    e.close()
