"""A5/2 weakened instance (config 4): layout, brute-force rate on a slice, pruned enumeration of 2^k."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pdsat_b200 import api, fixtures as fx

def a52_instance(n_known=32, seed=9, stream_len=114):
    nv, cl, _, ks = fx.a5_2_cnf(stream_len)
    offs, lits = fx.to_csr(cl)
    rng = np.random.default_rng(seed)
    secret = rng.integers(0, 2, 81)
    stream = fx.A52Model(secret).stream(stream_len)
    fixed = [2 * (v - 1) + (0 if bit else 1) for v, bit in zip(ks, stream)]
    fixed += [2 * v + (0 if secret[v] else 1) for v in range(64, 81)]
    known = sorted(int(v) for v in rng.choice(np.arange(64), size=n_known, replace=False))
    fixed += [2 * v + (0 if secret[v] else 1) for v in known]
    dset = [v + 1 for v in range(64) if v not in known]
    return nv, offs, lits, fixed, dset, secret

if __name__ == "__main__":
    nv, offs, lits, fixed, dset, secret = a52_instance()
    t = time.time(); db = api.ClauseDb(nv, offs, lits, device=0); db.set_fixed_assumptions(fixed)
    i = db.info()
    print("upload %.2fs vars %d live %d clauses %d gates %d gate_clauses %d wpc %d smem %d blob %d in_smem %d" % (time.time() - t, i.n_vars, i.n_live_vars, i.n_clauses_live, i.n_gates, i.n_gate_clauses, i.warps_per_cta, i.smem_bytes, i.blob_bytes, i.db_in_smem))
    k = len(dset)
    idx = sum(int(secret[v - 1]) << r for r, v in enumerate(dset))
    n = 1 << 22
    b = api.Batch(db, [api.Point(dset, n, api.MODE_ENUM, first_sample=(idx >> 22) << 22)], api.OUT_FLAGS)
    for _ in range(2):
        c = b.run()[0]
    f, p = b.last_ms()
    print("brute slice 2^22: fill %.3f ms propagate %.3f ms -> %.3g assignments/s conflicts %d full %d pops/group %.1f" % (f, p, n / (f + p) * 1e3, c.n_conflict, c.n_full, c.queue_pops / (n / 64)))
    b.close()
    for rep in range(2):
        t = time.time()
        surv, ism, st = db.enumerate(dset, max_survivors=1 << 16)
        dt = time.time() - t
        print("enumerate 2^%d: %.3f s  survivors %d models %d undecided %d propagated %d (%.3g of cube) batches %d secret_in %s" % (k, dt, len(surv), st.n_models, st.n_undecided, st.n_propagated, st.n_propagated / st.n_assignments, st.n_batches, idx in [int(x) for x in surv]))
