"""BASELINE configs 3 and 5 through bc_sweep_group for several strip heights (development aid): "rows_per_thread" of
the first member is the group's target strip height (0 = the plan's own choice)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc
T, D, J = 0.608, 1.966, 1.0

def series(sizes, boundary, n_rep, points, sweeps, seed, rows):
    engines = []
    for Ls in sizes:
        e = bc.Engine(Ls, boundary, n_rep, 0, seed + Ls)
        for r in range(n_rep):
            t, d = points[r % len(points)]
            e.set_params(t, d, J, replica=r)
        e.init_random(); e.sweep(4)
        engines.append(e)
    group = sorted((e for e in engines if e.L >= 128), key=lambda e: -e.L)
    small = [e for e in engines if e.L < 128]
    if rows: group[0].set_option("rows_per_thread", rows)
    bc.sweep_group(group, 16)
    for e in engines: e.sync()
    best = 0
    for _ in range(3):
        t0 = time.perf_counter()
        for e in small: e.sweep(sweeps, 0, fetch=False)
        bc.sweep_group(group, sweeps, 0)
        for e in engines: e.sync()
        wall = time.perf_counter() - t0
        best = max(best, sum(e.L * e.L * n_rep for e in engines) * sweeps / (wall * 1e9))
    # the group alone
    t0 = time.perf_counter(); bc.sweep_group(group, sweeps, 0)
    for e in group: e.sync()
    g_only = sum(e.L * e.L * n_rep for e in group) * sweeps / ((time.perf_counter() - t0) * 1e9)
    for e in engines: e.close()
    return best, g_only

pts = [(0.58 + 0.06 * i / 7, 1.94 + 0.05 * (i % 8) / 7) for i in range(8)]
for rows in (0, 4, 8, 16, 32):
    a = series([32, 64, 128, 256, 512, 1024, 2048], bc.PERIODIC, 8, pts, 400, 0x5EED0003, rows)
    b = series([64, 128, 256, 512, 1024], bc.OPEN, 64, [(T, D)], 200, 0x5EED0005, rows)
    print(json.dumps({"rows": rows, "config3": round(a[0], 1), "config3_group_only": round(a[1], 1), "config5": round(b[0], 1),
                      "config5_group_only": round(b[1], 1)}), flush=True)
