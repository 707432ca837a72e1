"""Host-side cost of one pdsat_batch_propagate call: python tools/host_cost.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pdsat_b200 import api
nv, cl, offs, lits, stream = bench.instance()
db = api.ClauseDb(nv, offs, lits, device=0)
db.set_fixed_assumptions(stream)
for n in (64, 10000000):
    b = api.Batch(db, [api.Point([177 - i for i in range(30)], n, api.MODE_MT19937, seed=1)])
    b.fill(); b.propagate(); b.sync()
    for reps in (200, 2000):
        t0 = time.perf_counter()
        for _ in range(reps):
            b.propagate()
        t1 = time.perf_counter()
        b.sync()
        t2 = time.perf_counter()
        print("n=%d reps=%d: host %.2f us/call, with sync %.2f us/call" % (n, reps, (t1 - t0) / reps * 1e6, (t2 - t0) / reps * 1e6))
    b.close()
