#!/bin/bash
# run tools/tune.py quick for every libbc_engine_*.so variant (development aid)
for f in blume_capel_b200/libbc_engine_*.so; do BC_ENGINE_LIB=$PWD/$f python tools/tune.py ${1:-quick} | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'], 'L', d['L'], 'rep', d['n_rep'], 'rows', d['plan']['rows_per_thread'], 'meas', d['measure_every'], '->', d['updates_per_ns'])
"; done
