"""Run one Bivium predict case on the GPU (profiling helper): python tools/run_case.py K N [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pdsat_b200 import api
k = int(sys.argv[1]); n = int(sys.argv[2]); iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nv, cl, offs, lits, stream = bench.instance()
db = api.ClauseDb(nv, offs, lits, device=0)
db.set_fixed_assumptions(stream)
sv = [177 - i for i in range(k)]
b = api.Batch(db, [api.Point(sv, n, api.MODE_MT19937, seed=1)])
for it in range(iters):
    c = b.run()[0]
f, p = b.last_ms()
print("k=%d n=%d fill %.3f ms propagate %.3f ms conflicts %d pops/group %.1f implied/group %.1f"
      % (k, n, f, p, c.n_conflict, c.queue_pops / (n / 64), c.implied_lits / (n / 64)))
