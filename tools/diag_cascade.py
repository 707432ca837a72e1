"""Per-group propagation statistics of bcp_kernel on cascade sets (diagnostics only)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pdsat_b200 import api

nv, cl, offs, lits, stream = bench.instance()
db = api.ClauseDb(nv, offs, lits, device=0)
db.set_fixed_assumptions(stream)
i = db.info()
print("live", i.n_live_vars, "clauses", i.n_clauses_live, "lits", i.n_lits_live, "wpc", i.warps_per_cta)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
for k in (30, 60, 80, 86, 90, 100, 120, 177):
    sv = [177 - j for j in range(k)]
    b = api.Batch(db, [api.Point(sv, n, api.MODE_MT19937, seed=1)])
    for _ in range(3):
        c = b.run()[0]
    f, p = b.last_ms()
    g = (n + 63) // 64
    print(json.dumps({"k": k, "ms": p, "samples_per_s": n / p * 1e3, "conflict": c.n_conflict / n, "full": c.n_full / n,
                      "mean_trail_ok": c.sum_trail / max(1, n - c.n_conflict), "pops_per_group": c.queue_pops / g,
                      "touched_per_group": c.implied_lits / g, "implied_any": c.n_implied_any / n}))
    b.close()
