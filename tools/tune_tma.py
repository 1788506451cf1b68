"""Pair table staged by cp.async.bulk (BC_PAIR_TMA=1) vs cp.async from all threads (0): bench batch and short strips."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.tune import run
for c in ((4096, 64, 0, 20), (4096, 64, 8, 20), (4096, 64, 16, 20), (2048, 8, 8, 200), (1024, 64, 16, 50), (1024, 64, 8, 50), (8192, 1, 16, 100)):
    ups, ms, plan = run(*c)
    print(os.environ.get("BC_ENGINE_LIB", "default"), c, round(ups, 1), plan["pair"], plan["rows_per_thread"], flush=True)
