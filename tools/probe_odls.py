"""ODLS order-10 pair CNF (config 5): predict over 8 cells of square A with random one-hot values."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pdsat_b200 import api, fixtures as fx

def odls_batch(db, n_o, seed=3):
    cells = [(1, 1), (1, 2), (2, 3), (3, 4), (4, 6), (5, 7), (6, 8), (8, 9)]
    ovars = [fx.odls_var(10, 0, i, j, z) for (i, j) in cells for z in range(10)]
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, 10, size=(n_o, len(cells)))
    vals = np.zeros((n_o, len(ovars)), dtype=np.uint8)
    for ci in range(len(cells)):
        vals[np.arange(n_o), ci * 10 + pick[:, ci]] = 1
    return ovars, vals

if __name__ == "__main__":
    n_o = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    onv, ocl = fx.odls_pair_cnf(10)
    ooffs, olits = fx.to_csr(ocl)
    t = time.time()
    odb = api.ClauseDb(onv, ooffs, olits, device=0)
    oi = odb.info()
    print("upload %.2fs live %d clauses %d wpc %d smem %d blob %d in_smem %d" % (time.time() - t, oi.n_live_vars, oi.n_clauses_live, oi.warps_per_cta, oi.smem_bytes, oi.blob_bytes, oi.db_in_smem))
    ovars, vals = odls_batch(odb, n_o)
    bo = api.Batch(odb, [api.Point(ovars, n_o, api.MODE_EXPLICIT, explicit_values=vals)])
    for _ in range(3):
        co = bo.run()[0]
    fo, po = bo.last_ms()
    g = (n_o + 63) // 64
    print("odls predict n=%d: propagate %.3f ms -> %.3g samples/s conflicts %.3f mean_trail %.1f pops/group %.1f touched/group %.1f" % (
        n_o, po, n_o / po * 1e3, co.n_conflict / n_o, co.sum_trail / max(1, n_o - co.n_conflict), co.queue_pops / g, co.implied_lits / g))
