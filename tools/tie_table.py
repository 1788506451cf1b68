import json, sys
from collections import OrderedDict
rows = [json.loads(l) for l in open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/s2_tune_tie.jsonl')]
t = OrderedDict()
for r in rows:
    k = (r['L'], r['n_rep'], json.dumps(r['opts']), r['measure_every'], r['boundary'])
    t.setdefault(k, {})[r['lib']] = (r['updates_per_ns'], r['pipelined'])
for k, v in t.items():
    print(k, {a.replace('libbc_engine', '').replace('.so', ''): b[0] for a, b in v.items()})
