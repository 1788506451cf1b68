"""Pair-table kernel vs per-site-lookup kernel (development aid): python tools/tune_pair.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.tune import run

if __name__ == "__main__":
    cases = []
    for me in (0, 1):
        for pair in (0, 1):
            cases.append((4096, 64, 32, 20, me, 0, (("pair_min_rows", pair),)))
    for rows in (16, 8):
        for pair in (0, 1):
            cases.append((4096, 64, rows, 20, 0, 0, (("pair_min_rows", pair),)))
    for L, nr, rows, sw in ((4096, 1, 4, 400), (4096, 1, 8, 400), (2048, 8, 8, 200), (1024, 8, 4, 400), (1024, 64, 16, 50),
                            (8192, 1, 8, 100), (512, 64, 8, 200)):
        for pair in (0, 1):
            cases.append((L, nr, rows, sw, 0, 0, (("pair_min_rows", pair),)))
    cases.append((4096, 64, 32, 20, 0, 1, (("pair_min_rows", 0),)))
    cases.append((4096, 64, 32, 20, 0, 1, (("pair_min_rows", 1),)))
    for c in cases:
        ups, ms, plan = run(*c)
        print(json.dumps({"L": c[0], "n_rep": c[1], "rows": c[2], "sweeps": c[3], "measure_every": c[4], "boundary": c[5],
                          "opts": c[6], "updates_per_ns": round(ups, 1), "ms": round(ms, 3), "pair": plan["pair"]}), flush=True)
