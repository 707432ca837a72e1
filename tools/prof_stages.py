"""Device-time breakdown of the batch stages (fill = mt19937 jump + bits + transpose; propagate)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pdsat_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nv, cl, offs, lits, stream = bench.instance()
db = api.ClauseDb(nv, offs, lits, device=0)
db.set_fixed_assumptions(stream)
for k, sv in (("ending2_k30", bench.SET_VARS), ("beginning1_k30", list(range(1, 31))),
              ("ending2_k86", [177 - i for i in range(86)]), ("ending2_k120", [177 - i for i in range(120)])):
    b = api.Batch(db, [api.Point(sv, n, api.MODE_MT19937, seed=1)])
    for it in range(3):
        b.reseed(0, 1 + it, 0)
        t0 = time.perf_counter()
        c = b.run()[0]
        wall = time.perf_counter() - t0
    f, p = b.last_ms()
    print("%-16s n=%d fill %.3f ms propagate %.3f ms wall %.3f ms -> %.3g samples/s (kernel) conflicts %d implied_any %d pops/group %.1f implied_lits/group %.1f"
          % (k, n, f, p, wall * 1e3, n / p * 1e3, c.n_conflict, c.n_implied_any, c.queue_pops / (n / 64), c.implied_lits / (n / 64)))
    b.close()
