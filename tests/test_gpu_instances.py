"""GPU: the other instance families of BASELINE.json — order-10 ODLS pair CNF (clause DB far
larger than shared memory: 434 440 clauses, 10-literal clauses -> chained records) and the
generated A5/2 CNF (8 745 variables: one warp per CTA) — against the CPU oracle and the
reference's own known answers."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def odls(oracle_mod, cuda_lib):
    from pdsat_b200 import api, fixtures as fx
    nv, cl = fx.odls_pair_cnf(10)
    offs, lits = fx.to_csr(cl)
    db = api.ClauseDb(nv, offs, lits, device=0)
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    pairs = fx.parse_odls_pairs(open(os.path.join(ROOT, "tests", "golden", "odls_pairs_sample.txt")).read())
    return nv, db, P, pairs


def test_odls_known_pairs_kat(odls):
    """the 200 positive cell literals of a known pair BCP to a full,
    conflict-free assignment; changing one cell conflicts.  Assumptions are positive literals
    only (src_common/latin_squares.cpp:145-150)."""
    from pdsat_b200 import api, fixtures as fx
    nv, db, P, pairs = odls
    info = db.info()
    assert (info.n_vars, info.n_clauses_live, info.max_clause_len) == (2000, 434440, 10)
    pts = []
    for a, b in pairs:
        good = fx.odls_assumptions_from_squares(10, a, b)
        bad = list(good)
        bad[37] = bad[37] + 1 if (bad[37] - 1) % 10 != 9 else bad[37] - 1  # another value in the same cell
        for vars_ in (good, bad):
            pts.append(api.Point(vars_, 1, api.MODE_EXPLICIT, explicit_values=np.ones((1, 200), dtype=np.uint8)))
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL, max_models=16)
    costs = b.run()
    for i, c in enumerate(costs):
        if i % 2 == 0:
            assert (c.n_conflict, c.n_full, c.sum_trail) == (0, 1, 2000)
        else:
            assert c.n_conflict == 1
    found, pt_of, _, models = b.models()
    assert found == len(pairs)
    for i in range(found):
        a, bb = pairs[pt_of[i] // 2]
        for v in fx.odls_assumptions_from_squares(10, a, bb):
            assert models[i][v - 1] == 1
        assert models[i].sum() == 200


def test_odls_random_assumptions_vs_oracle(odls):
    from pdsat_b200 import api
    nv, db, P, pairs = odls
    rng = np.random.default_rng(4)
    pts, meta = [], []
    for k in (3, 10, 25, 40):
        sv = sorted(int(v) for v in rng.choice(np.arange(1, nv + 1), size=k, replace=False))
        n = 70
        vals = (rng.random((n, k)) < (0.15 if k > 10 else 0.5)).astype(np.uint8)  # mostly negative: fewer clashes
        pts.append(api.Point(sv, n, api.MODE_EXPLICIT, explicit_values=vals))
        meta.append((sv, vals))
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL | api.OUT_ASSIGN)
    b.run()
    seen_ok = seen_conf = 0
    for p, (sv, vals) in enumerate(meta):
        conflict, full = b.flags(p)
        trail = b.trail(p)
        asg = b.assign(p)
        for s in range(len(vals)):
            a = [2 * (v - 1) + (0 if vals[s][j] else 1) for j, v in enumerate(sv)]
            r, ea = P.bcp(a)
            assert bool(conflict[s]) == bool(r.conflict), (p, s)
            if not r.conflict:
                assert trail[s] == r.trail_size and (asg[s] == ea).all(), (p, s)
                seen_ok += 1
            else:
                seen_conf += 1
    assert seen_ok > 20 and seen_conf > 20


def test_a52_weakened_instance(oracle_mod, cuda_lib):
    """Generated A5/2 CNF (no A5/2 CNF ships in the reference): keystream + R4 + 32 of the 64
    R1|R2|R3 bits fixed; the other 32 form the decomposition set.  Random samples against the
    oracle, then solve-mode enumeration of a 10-variable sub-cube that must contain the secret."""
    from pdsat_b200 import api, fixtures as fx
    nv, cl, _, ks = fx.a5_2_cnf(114)
    offs, lits = fx.to_csr(cl)
    rng = np.random.default_rng(9)
    secret = rng.integers(0, 2, 81)
    stream = fx.A52Model(secret).stream(114)
    fixed = [2 * (v - 1) + (0 if bit else 1) for v, bit in zip(ks, stream)]
    fixed += [2 * v + (0 if secret[v] else 1) for v in range(64, 81)]          # R4
    known = sorted(int(v) for v in rng.choice(np.arange(64), size=32, replace=False))
    fixed += [2 * v + (0 if secret[v] else 1) for v in known]
    dset = [v + 1 for v in range(64) if v not in known]
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(fixed)
    assert db.info().status == api.PDSAT_OK
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    P.add_units(fixed)
    n = 96
    b = api.Batch(db, [api.Point(dset, n, api.MODE_MT19937, seed=4)], api.OUT_FLAGS | api.OUT_TRAIL | api.OUT_ASSIGN)
    b.run()
    conflict, full = b.flags(0)
    trail = b.trail(0)
    asg = b.assign(0)
    A = oracle_mod.sample_assumptions(4, dset, 0, n)
    for s in range(n):
        r, ea = P.bcp(list(A[s]))
        assert bool(conflict[s]) == bool(r.conflict), s
        if not r.conflict:
            assert trail[s] == r.trail_size and (asg[s] == ea).all(), s
    b.close()
    # solve mode: fix 22 more set variables to the secret, enumerate the remaining 10
    extra = [2 * (v - 1) + (0 if secret[v - 1] else 1) for v in dset[:22]]
    db.set_fixed_assumptions(fixed + extra)
    sub = dset[22:]
    b = api.Batch(db, [api.Point(sub, 1 << 10, api.MODE_ENUM)], api.OUT_FLAGS, max_models=64)
    c = b.run()[0]
    conflict, full = b.flags(0)
    idx = sum(int(secret[v - 1]) << r for r, v in enumerate(sub))
    assert not conflict[idx]
    found, _, smp, models = b.models()
    if full[idx]:
        assert idx in [int(x) for x in smp]
        m = models[[int(x) for x in smp].index(idx)]
        assert [int(x) for x in m[:81]] == [int(x) for x in secret]
    assert c.n == 1024 and c.n_conflict >= 1


def _clauses_satisfied(cl, model):
    m = np.asarray(model, dtype=np.int8)
    for c in cl:
        if not any((m[abs(l) - 1] == 1) == (l > 0) for l in c):
            return False
    return True


def test_odls_solve_one_problem_semantics(odls, oracle_mod):
    """latin_square::SolveOneProblem (src_common/latin_squares.cpp:133-233): positive-literal
    assumptions on one persistent solver.  (1) the 11 rows of the reference's own sample input
    (VS_latin_conseq/in:2-12, 26 literals each): the device's BCP equals the reference BCP, the
    hand-off under a small conflict budget never contradicts it;  (2) a known pair with the last
    rows of square B left open: BCP + host CDCL complete it to a valid ODLS pair;  (3) one wrong
    cell: refuted."""
    from pdsat_b200 import api, fixtures as fx
    nv, db, P, pairs = odls
    _, cl = fx.odls_pair_cnf(10)
    rows = [[int(x) for x in l.split()] for l in open(os.path.join(ROOT, "tests", "golden", "latin_conseq_rows.txt"))
            if l.strip() and not l.startswith("#")]
    assert len(rows) == 11 and all(len(r) == 26 for r in rows)
    pts = [api.Point(r, 1, api.MODE_EXPLICIT, explicit_values=np.ones((1, 26), dtype=np.uint8)) for r in rows]
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL | api.OUT_ASSIGN)
    b.run()
    for p, r in enumerate(rows):
        res, a = P.bcp([2 * (v - 1) for v in r])
        conflict, full = b.flags(p)
        assert bool(conflict[0]) == bool(res.conflict) and not full[0]
        if not res.conflict:
            assert int(b.trail(p)[0]) == res.trail_size and (b.assign(p)[0] == a).all()
        status, _, ms, models, st = b.handoff(p, max_models=1, max_conflicts=300, n_threads=1)
        if res.conflict:
            assert status[0] == api.UNSAT and st.n_problems == 0
        else:
            assert st.n_problems == 1 and status[0] in (api.SAT, api.UNSAT, api.UNKNOWN)
            if status[0] == api.SAT:
                assert _clauses_satisfied(cl, models[0]) and all(models[0][v - 1] == 1 for v in r)
    b.close()
    # (2) + (3): square A and the first 7 rows of B of a known pair
    a_sq, b_sq = pairs[0]
    allv = fx.odls_assumptions_from_squares(10, a_sq, b_sq)
    part = allv[:100] + allv[100:170]
    bad = list(part)
    bad[150] = bad[150] + 1 if (bad[150] - 1) % 10 != 9 else bad[150] - 1
    pts = [api.Point(vs, 1, api.MODE_EXPLICIT, explicit_values=np.ones((1, len(vs)), dtype=np.uint8)) for vs in (part, bad)]
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL, max_models=4)
    b.run()
    status, _, ms, models, st = b.handoff(0, max_models=1, max_conflicts=200000, n_threads=1)
    assert status[0] == api.SAT
    if st.n_problems:  # completed by search: a valid pair that extends the assumptions
        assert _clauses_satisfied(cl, models[0]) and all(models[0][v - 1] == 1 for v in part)
    else:              # BCP alone completed it
        found, _, _, bm = b.models()
        assert found >= 1 and _clauses_satisfied(cl, bm[0])
    status, *_ = b.handoff(1, max_conflicts=200000, n_threads=1)
    assert status[0] == api.UNSAT
    b.close()
