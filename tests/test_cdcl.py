"""The library's own host CDCL (pdsat_b200/host/cdcl.hpp) - the hand-off target for what BCP
leaves open.  Verdicts are properties of the formula: they must equal those of the reference's
solve() under the same assumptions (oracle/_ref = unmodified MiniSat, orc_solve_assumps replaying
src_mpi/mpi_solver.cpp:218-245); every model must satisfy every clause (CheckSATset,
src_mpi/mpi_base.cpp:717-752); where the solution is unique the model is the reference's."""
import os

import numpy as np
import pytest

from pdsat_b200 import api, fixtures as fx


def _ref_available(orc):
    return orc.have_ref()


def _check_model(clauses, model):
    for c in clauses:
        assert any((model[abs(l) - 1] == 1) == (l > 0) for l in c), c


def _rand_3sat(rng, nv, nc):
    return [[int(v + 1) * (1 if rng.random() < 0.5 else -1) for v in rng.choice(nv, size=3, replace=False)]
            for _ in range(nc)]


def test_cdcl_random_3sat_matches_reference(oracle_mod, cuda_lib):
    orc = oracle_mod
    if not _ref_available(orc):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(7)
    n_sat = n_unsat = 0
    for inst in range(40):
        nv = 50
        cls = _rand_3sat(rng, nv, int(nv * (3.6 + 1.4 * rng.random())))
        offs, lits = fx.to_csr(cls)
        db = api.ClauseDb(nv, offs, lits, device=-1)
        R = orc.Oracle(nv, offs, lits, "ref")
        A = np.array([[2 * int(v) + int(rng.integers(0, 2)) for v in rng.choice(nv, size=4, replace=False)]
                      for _ in range(6)], dtype=np.int32)
        status, props, mp, models, st = db.cdcl_solve_many(A, max_models=6, n_threads=1)
        for i in range(len(A)):
            r, state, model = R.solve_assumps(list(A[i]))
            assert status[i] == (api.SAT if r.ret == 0 else api.UNSAT), (inst, i)
        for j, i in enumerate(mp):
            _check_model(cls, models[j])
            for l in A[i]:
                assert models[j][l >> 1] == (1 if (l & 1) == 0 else 0)
        n_sat += int((status == api.SAT).sum())
        n_unsat += int((status == api.UNSAT).sum())
        assert st.n_sat + st.n_unsat == len(A) and st.n_unknown == 0
        db.close()
    assert n_sat > 20 and n_unsat > 20


def test_cdcl_weakened_bivium_unique_solution(bivium, oracle_mod, cuda_lib):
    """template + keystream + 125 known state bits: the remaining 52 bits have one solution (the
    secret).  Under assumptions on a few of them: the secret's values -> SAT with the secret as
    the model; one flipped bit -> UNSAT.  Verdicts also against the reference."""
    import bench
    orc = oracle_mod
    nv, cl, offs, lits = bivium
    stream = list(bench.instance()[4])
    secret = bench.mt19937_bool_rand_numpy(bench.SECRET_SEED, 177)
    n_fixed = 125
    fixed = stream + [2 * v + (0 if secret[v] else 1) for v in range(n_fixed)]
    db = api.ClauseDb(nv, offs, lits, device=-1)
    db.set_fixed_assumptions(fixed)
    rng = np.random.default_rng(3)
    rows = []
    for t in range(12):
        vs = sorted(int(v) for v in rng.choice(np.arange(n_fixed, 177), size=10, replace=False))
        row = [2 * v + (0 if secret[v] else 1) for v in vs]
        if t % 2:
            row[int(rng.integers(0, 10))] ^= 1
        rows.append(row)
    A = np.array(rows, dtype=np.int32)
    status, props, mp, models, st = db.cdcl_solve_many(A, max_models=12)
    want = [api.SAT if t % 2 == 0 else api.UNSAT for t in range(12)]
    assert list(status) == want
    for j, i in enumerate(mp):
        assert list(models[j][:177]) == [int(x) for x in secret]
        _check_model(cl, models[j])
    if _ref_available(orc):
        R = orc.Oracle(nv, offs, lits, "ref")
        R.add_units(fixed)
        for i in range(len(A)):
            r, state, model = R.solve_assumps(list(A[i]))
            assert status[i] == (api.SAT if r.ret == 0 else api.UNSAT)
    db.close()


def test_cdcl_limits_and_threads(bivium, cuda_lib):
    import bench
    nv, cl, offs, lits = bivium
    stream = list(bench.instance()[4])
    secret = bench.mt19937_bool_rand_numpy(bench.SECRET_SEED, 177)
    db = api.ClauseDb(nv, offs, lits, device=-1)
    db.set_fixed_assumptions(stream + [2 * v + (0 if secret[v] else 1) for v in range(110)])
    rng = np.random.default_rng(5)
    A = np.array([[2 * v + int(rng.integers(0, 2)) for v in range(110, 130)] for _ in range(64)], dtype=np.int32)
    s1, p1, _, _, st1 = db.cdcl_solve_many(A, n_threads=1)
    s4, p4, _, _, st4 = db.cdcl_solve_many(A, n_threads=4)
    assert (s1 == s4).all() and st4.threads == 4 and st1.threads == 1  # verdicts do not depend on the schedule
    assert st1.n_unknown == 0
    # a conflict budget of 1 on an instance that needs search: interrupted, never a wrong verdict
    db2 = api.ClauseDb(nv, offs, lits, device=-1)
    db2.set_fixed_assumptions(stream + [2 * v + (0 if secret[v] else 1) for v in range(60)])
    sL, _, _, _, stL = db2.cdcl_solve_many(A[:4], max_conflicts=1, n_threads=1)
    assert stL.n_unknown > 0 and set(sL.tolist()) <= {api.UNSAT, api.UNKNOWN}
    db.close()
    db2.close()


def test_cdcl_level0_unsat_and_empty(cuda_lib):
    db = api.ClauseDb.from_clauses(2, [[1], [-1, 2], [-2]], device=-1)
    status, *_ = db.cdcl_solve_many(np.zeros((3, 0), dtype=np.int32))
    assert list(status) == [api.UNSAT] * 3
    db.close()
    db = api.ClauseDb.from_clauses(3, [[1, 2, 3]], device=-1)
    status, props, mp, models, st = db.cdcl_solve_many(np.array([[1, 3, 4], [1, 3, 5]], dtype=np.int32), max_models=2)
    assert list(status) == [api.SAT, api.UNSAT]
    assert models[0][2] == 1
    db.close()


def _handoff_vs_reference(orc, nv, cls, offs, lits, fixed, sv, vals, secret_rows=()):
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(fixed)
    n = len(vals)
    b = api.Batch(db, [api.Point(sv, n, api.MODE_EXPLICIT, explicit_values=vals)], api.OUT_FLAGS | api.OUT_TRAIL)
    b.run()
    assert (b.sample_values(0) == vals).all()
    conflict, full = b.flags(0)
    status, cost, ms, models, st = b.handoff(0, max_models=8, want_cost=True)
    n_open = int((~conflict & ~full).sum())
    assert st.n_problems == n_open
    R = orc.Oracle(nv, offs, lits, "ref")
    R.add_units(fixed)
    for s in range(n):
        assum = [2 * (v - 1) + (0 if vals[s][j] else 1) for j, v in enumerate(sv)]
        r, state, model = R.solve_assumps(assum)
        assert status[s] == (api.SAT if r.ret == 0 else api.UNSAT), s
        if conflict[s]:
            assert r.ret == 1
    for j in range(len(ms)):
        for c in cls:
            assert any((models[j][abs(l) - 1] == 1) == (l > 0) for l in c)
        srow = int(ms[j])
        for jj, v in enumerate(sv):
            assert models[j][v - 1] == vals[srow][jj]
    info = db.info()
    assert (cost[conflict] == info.n_base_assigned + len(sv)).all()
    b.close()
    db.close()
    return n_open, status, st


@pytest.mark.gpu
def test_gpu_handoff_every_sample_gets_the_reference_verdict(bivium, oracle_mod, cuda_lib):
    """predict-style batches: BCP decides some samples, the host CDCL the rest; the per-sample
    verdict equals the reference's solve() under the same assumptions."""
    import bench
    orc = oracle_mod
    if not _ref_available(orc):
        pytest.skip("oracle/_ref not built")
    # random 3-SAT below the threshold: BCP decides little, search decides the rest
    rng = np.random.default_rng(21)
    nv = 60
    cls = _rand_3sat(rng, nv, 240)
    offs, lits = fx.to_csr(cls)
    sv = sorted(int(v) + 1 for v in rng.choice(nv, size=8, replace=False))
    vals = rng.integers(0, 2, size=(256, 8)).astype(np.uint8)
    n_open, status, st = _handoff_vs_reference(orc, nv, cls, offs, lits, [], sv, vals)
    assert n_open > 50 and st.n_sat > 0 and st.n_unsat > 0
    # weakened Bivium: 110 known state bits + keystream, 20 more assumed at random (+ the secret's)
    nv, cl, offs, lits = bivium
    stream = list(bench.instance()[4])
    secret = bench.mt19937_bool_rand_numpy(bench.SECRET_SEED, 177)
    fixed = stream + [2 * v + (0 if secret[v] else 1) for v in range(110)]
    sv = list(range(111, 131))
    vals = np.random.default_rng(8).integers(0, 2, size=(128, len(sv))).astype(np.uint8)
    vals[5] = [secret[v - 1] for v in sv]
    n_open, status, st = _handoff_vs_reference(orc, nv, cl, offs, lits, fixed, sv, vals)
    assert status[5] == api.SAT
