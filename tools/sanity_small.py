"""Small end-to-end run touching every kernel (for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pdsat_b200 import api, fixtures as fx
nv, cl, offs, lits, stream = bench.instance()
db = api.ClauseDb(nv, offs, lits, device=0)
db.set_fixed_assumptions(stream)
pts = [api.Point(bench.SET_VARS, 3000, api.MODE_MT19937, seed=1, first_sample=700),
       api.Point([177 - i for i in range(86)], 2000, api.MODE_MT19937, seed=2),
       api.Point(list(range(1, 178)), 300, api.MODE_COUNTER, seed=3),
       api.Point(list(range(160, 172)), 1 << 12, api.MODE_ENUM)]
b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL | api.OUT_ASSIGN, max_models=8)
print([(c.n, c.n_conflict, c.n_full) for c in b.run()])
db2 = api.ClauseDb(nv, offs, lits, device=0)   # no keystream: full models
b2 = api.Batch(db2, [api.Point(list(range(1, 178)), 200, api.MODE_MT19937, seed=5)], api.OUT_FLAGS, max_models=4)
print([(c.n, c.n_conflict, c.n_full) for c in b2.run()], b2.models()[0])
onv, ocl = fx.odls_pair_cnf(10)
ooffs, olits = fx.to_csr(ocl)
odb = api.ClauseDb(onv, ooffs, olits, device=0)
rng = np.random.default_rng(1)
sv = sorted(int(v) for v in rng.choice(np.arange(1, onv + 1), size=30, replace=False))
vals = (rng.random((200, 30)) < 0.2).astype(np.uint8)
b3 = api.Batch(odb, [api.Point(sv, 200, api.MODE_EXPLICIT, explicit_values=vals)], api.OUT_FLAGS)
print([(c.n, c.n_conflict, c.n_full) for c in b3.run()])
