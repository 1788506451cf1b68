"""One lattice (BASELINE config 2 taken literally: L = 4096, L2-resident, launch-bound) and other launches too short
to amortise a kernel boundary: persistent kernel (one cooperative launch, grid barriers) vs CUDA-graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.tune import run
for L, n_rep in ((4096, 1), (2048, 1), (1024, 1), (8192, 1), (2048, 8), (1024, 64), (512, 64)):
    for persist in (0, 1):
        for rows in ((0,) if not persist else (0, 2, 4, 8, 16)):
            for b in (0, 1):
                ups, ms, plan = run(L, n_rep, rows, 400, boundary=b, opts=(("persist", persist),))
                print(L, n_rep, "open" if b else "periodic", "persist", persist, "rows", rows, round(ups, 1), flush=True)
