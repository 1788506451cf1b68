"""Golden observables for BASELINE config 4 (L = 65536 periodic, one lattice of 4 GiB) from the CPU ORACLE.

bench.py's `slab` record (N > 1 GPUs) and tests/test_gpu_fullsize.py compare the (M, Q, B) a slab-decomposed GPU run
reports after SWEEPS sweeps with these numbers, so the multi-GPU path is pinned by the oracle at the full size and
not only by a single-GPU run of the same kernels.  The oracle's half-sweep is run over row ranges on all host cores
(ctypes releases the GIL; sites of one colour never read each other, so the split cannot change the result).

    python tests/golden/make_slab_golden.py [L] [sweeps]      # about 3 minutes on 8 cores for L = 65536, 2 sweeps
"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
SWEEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
SEED = 0x5EED0004
T, D, J = 0.608, 1.966, 1.0


def main():
    workers = os.cpu_count() or 1
    tab = oracle.build_table(T, D, J)
    t0 = time.time()
    spins = oracle.init_random(L, SEED, 0)
    print(f"init {time.time() - t0:.1f} s", flush=True)
    bands = [(i * L // (4 * workers), (i + 1) * L // (4 * workers)) for i in range(4 * workers)]
    out = {"L": L, "boundary": "periodic", "seed": SEED, "T": T, "D": D, "J": J, "source": "oracle/bc_oracle.c", "after": {}}
    out["after"]["0"] = list(oracle.observables(spins))
    with ThreadPoolExecutor(workers) as ex:
        for t in range(SWEEPS):
            for colour in (0, 1):
                list(ex.map(lambda b: oracle.half_sweep_rows(spins, tab, SEED, 0, t, colour, b[0], b[1]), bands))
            out["after"][str(t + 1)] = list(oracle.observables(spins))
            print(f"sweep {t + 1}: {out['after'][str(t + 1)]}  ({time.time() - t0:.1f} s)", flush=True)
    path = os.path.join(ROOT, "tests", "golden", f"slab_L{L}.json")
    prev = {}
    if os.path.exists(path):
        prev = json.load(open(path))
    prev.update(out)
    json.dump(prev, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
