#!/bin/bash
# bench-batch timing of every libbc_engine*.so variant (development aid)
for f in blume_capel_b200/libbc_engine*.so; do BC_ENGINE_LIB=$PWD/$f python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
from tools.tune import run
for case in ((4096, 64, 32, 20, 0), (4096, 64, 32, 20, 1), (4096, 64, 16, 20, 0), (4096, 64, 32, 20, 0, 1), (2048, 8, 8, 200, 0), (1024, 64, 16, 50, 0)):
    ups, ms, plan = run(*case)
    print(os.path.basename("$f"), case, round(ups, 1), plan.get("pair"), flush=True)
PY
done
