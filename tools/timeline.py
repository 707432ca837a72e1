"""Phase timestamps of bcp_kernel (library built with -DPDSAT_TIMELINE): python tools/timeline.py K N"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pdsat_b200 import api
k = int(sys.argv[1]); n = int(sys.argv[2])
nv, cl, offs, lits, stream = bench.instance()
db = api.ClauseDb(nv, offs, lits, device=0)
db.set_fixed_assumptions(stream)
b = api.Batch(db, [api.Point([177 - i for i in range(k)], n, api.MODE_MT19937, seed=1)])
lib = api.lib()
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(4):
    b.fill(); b.sync()
    flush.fill_(it); torch.cuda.synchronize()
    b.propagate(); b.sync()
    out = (ctypes.c_ulonglong * (148 * 8))()
    lib.pdsat_debug_timeline(out)
    t = np.array(out, dtype=np.int64).reshape(148, 8)
    t0 = t[:, 0].min()
    names = ["entry", "prologue", "first tile", "loop end", "blob waited", "flushed", "ticket", "published"]
    print("iter", it, "propagate_ms", b.last_ms()[1])
    for i, nm in enumerate(names):
        col = t[:, i] - t0
        col = col[t[:, i] > 0]
        if len(col): print("  %-12s min %7.2f  median %7.2f  max %7.2f us (%d CTAs)" % (nm, col.min() / 1e3, np.median(col) / 1e3, col.max() / 1e3, len(col)))
