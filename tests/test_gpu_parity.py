"""GPU: the CUDA path, called through the C-ABI, against (a) the golden vectors produced by
the unmodified reference MiniSat and (b) the CPU oracle on fresh seeded inputs.
Bit-exact: conflict flag, full flag, implied-literal set, trail size, models, integer costs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = ["ending2_k30", "beginning1_k30", "ending2_k85", "ending2_k86", "ending2_k90", "ending2_k100",
         "ending2_k120", "ending2_k177", "nostream_k177", "nostream_k150"]
TE_CASES = ["te_k80", "te_k86", "te_k88", "te_k100"]


def _golden_assign(golden, case, nv):
    t = np.unpackbits(golden[case + "_assign"], axis=1)[:, :nv].astype(bool)
    f = np.unpackbits(golden[case + "_assign_f"], axis=1)[:, :nv].astype(bool)
    a = np.full(t.shape, 2, dtype=np.uint8)
    a[t] = 0
    a[f] = 1
    return a


def _check_against_rows(api, batch, rows, gold_assign, n_base_check=True):
    n = len(rows)
    costs = batch.costs()[0]
    conflict, full = batch.flags(0)
    trail = batch.trail(0)
    asg = batch.assign(0, 0, n)
    ok = rows[:, 0] == 0
    assert (conflict == (rows[:, 0] == 1)).all()
    assert (full == (rows[:, 1] == 1)).all()
    assert (trail[ok] == rows[ok, 4]).all()
    # Solver::propagations of a conflict-free sample == its trail size
    assert (trail[ok] == rows[ok, 5]).all()
    assert (asg[ok] == gold_assign[ok]).all()
    assert costs.n == n and costs.n_conflict == int(rows[:, 0].sum()) and costs.n_full == int(rows[:, 1].sum())
    assert costs.sum_trail == int(rows[ok, 4].sum())
    assert costs.sumsq_trail == int((rows[ok, 4] ** 2).sum())


@pytest.mark.parametrize("fold", [True, False])
@pytest.mark.parametrize("case", CASES)
def test_gpu_matches_reference_golden(case, fold, bivium, golden, cuda_lib):
    from pdsat_b200 import api
    nv, _, offs, lits = bivium
    sv = golden[case + "_set"]
    seed, with_ks = (int(x) for x in golden[case + "_seed"])
    rows = golden[case + "_rows"]
    n = len(rows)
    db = api.ClauseDb(nv, offs, lits, device=0)
    stream = golden["stream_lits"]
    if with_ks and fold:
        db.set_fixed_assumptions(stream)
        pts = [api.Point(list(sv), n, api.MODE_MT19937, seed)]
    elif with_ks:
        # keystream literals as per-sample explicit assumptions on the unfolded template
        from oracle import oracle as orc
        bits = orc.bool_rand(seed, 0, n * len(sv)).reshape(n, len(sv))
        sbits = np.tile(((stream & 1) == 0).astype(np.uint8), (n, 1))
        vars_ = list(sv) + [int(l >> 1) + 1 for l in stream]
        pts = [api.Point(vars_, n, api.MODE_EXPLICIT, explicit_values=np.concatenate([bits, sbits], axis=1))]
    else:
        if not fold:
            pytest.skip("same as fold=True without keystream")
        pts = [api.Point(list(sv), n, api.MODE_MT19937, seed)]
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL | api.OUT_ASSIGN, max_models=256)
    b.fill()
    b.propagate()
    _check_against_rows(api, b, rows, _golden_assign(golden, case, nv))
    n_found, pt_of, smp_of, models = b.models()
    assert n_found == int(rows[:, 1].sum())
    if n_found:
        ga = _golden_assign(golden, case, nv)
        for i in range(len(smp_of)):
            assert (models[i] == (ga[int(smp_of[i])] == 0)).all()
    b.close()
    db.close()


@pytest.mark.parametrize("case", TE_CASES)
def test_gpu_matches_reference_golden_te(case, bivium, golden, cuda_lib):
    """te > 0 mode: known state + own keystream per sample (EXPLICIT values)."""
    from pdsat_b200 import api
    nv, _, offs, lits = bivium
    sv = golden[case + "_set"]
    rows = golden[case + "_rows"]
    n = len(rows)
    states = np.unpackbits(golden["te_states"], axis=1)[:, :177]
    streams = np.unpackbits(golden["te_streams"], axis=1)[:, :200]
    vals = np.concatenate([states[:, [v - 1 for v in sv]], streams], axis=1).astype(np.uint8)
    vars_ = list(sv) + list(range(578, 778))
    db = api.ClauseDb(nv, offs, lits, device=0)
    b = api.Batch(db, [api.Point(vars_, n, api.MODE_EXPLICIT, explicit_values=vals)],
                  api.OUT_FLAGS | api.OUT_TRAIL | api.OUT_ASSIGN, max_models=256)
    b.fill()
    b.propagate()
    _check_against_rows(api, b, rows, _golden_assign(golden, case, nv))
    b.close()
    db.close()


def test_gpu_vs_oracle_random_sets(bivium, golden, oracle_mod, cuda_lib):
    """Fresh seeded inputs, ragged sample counts, several points in one batch."""
    from pdsat_b200 import api
    nv, _, offs, lits = bivium
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    stream = list(golden["stream_lits"])
    P.add_units(stream)
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    rng = np.random.default_rng(11)
    pts, sets = [], []
    for i, (k, n) in enumerate([(1, 1), (5, 63), (40, 65), (84, 130), (86, 200), (87, 64), (177, 70), (60, 129)]):
        sv = sorted(int(v) for v in rng.choice(np.arange(1, 178), size=k, replace=False))
        sets.append(sv)
        pts.append(api.Point(sv, n, api.MODE_MT19937, seed=i + 1, first_sample=i * 7))
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL | api.OUT_ASSIGN, max_models=64)
    costs = b.run()
    for p, (pt, sv) in enumerate(zip(pts, sets)):
        A = oracle_mod.sample_assumptions(pt.seed, sv, pt.first_sample, pt.n_samples)
        conflict, full = b.flags(p)
        trail = b.trail(p)
        asg = b.assign(p)
        exp_sum = exp_sq = exp_conf = 0
        for s in range(pt.n_samples):
            r, a = P.bcp(list(A[s]))
            assert bool(conflict[s]) == bool(r.conflict), (p, s)
            exp_conf += r.conflict
            if not r.conflict:
                assert trail[s] == r.trail_size and bool(full[s]) == bool(r.full)
                assert (asg[s] == a).all(), (p, s)
                exp_sum += r.trail_size
                exp_sq += r.trail_size ** 2
        c = costs[p]
        assert (c.n, c.n_conflict, c.sum_trail, c.sumsq_trail) == (pt.n_samples, exp_conf, exp_sum, exp_sq)
    b.close()
    db.close()


def test_gpu_set_var_fixed_at_level0(bivium, golden, oracle_mod, cuda_lib):
    """A set variable that level 0 already assigned: equal value is a no-op, the opposite
    value is a conflict (addClause_ drops the false literal, Solver.cc:288-295)."""
    from pdsat_b200 import api
    nv, _, offs, lits = bivium
    stream = list(golden["stream_lits"])
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    P.add_units(stream)
    sv = [170, 578, 171, 600]  # 578 and 600 are keystream vars (fixed)
    n = 100
    b = api.Batch(db, [api.Point(sv, n, api.MODE_MT19937, seed=9)], api.OUT_FLAGS | api.OUT_TRAIL)
    b.run()
    conflict, _ = b.flags(0)
    trail = b.trail(0)
    A = oracle_mod.sample_assumptions(9, sv, 0, n)
    for s in range(n):
        r, _ = P.bcp(list(A[s]))
        assert bool(conflict[s]) == bool(r.conflict)
        if not r.conflict:
            assert trail[s] == r.trail_size
    assert 0 < conflict.sum() < n


def test_gpu_unsat_database(cuda_lib):
    from pdsat_b200 import api, fixtures as fx
    offs, lits = fx.to_csr([[1], [-1, 2], [-2, -1]])
    db = api.ClauseDb(3, offs, lits, device=0)
    assert db.info().status == api.PDSAT_ERR_UNSAT
    c = db.eval_batch([api.Point([3], 70, api.MODE_ENUM)])[0]
    assert c.n == 70 and c.n_conflict == 70


def test_gpu_enum_mode_small_cnf(oracle_mod, cuda_lib):
    """Solve-mode enumeration order (src_mpi/mpi_base.cpp:148-169) on a small random 3-CNF:
    every one of the 2^k assignments against the oracle, models included."""
    from pdsat_b200 import api, fixtures as fx
    rng = np.random.default_rng(3)
    nv = 24
    clauses = []
    for _ in range(70):
        vs = rng.choice(np.arange(1, nv + 1), size=3, replace=False)
        clauses.append([int(v) * (1 if rng.integers(0, 2) else -1) for v in vs])
    offs, lits = fx.to_csr(clauses)
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    db = api.ClauseDb(nv, offs, lits, device=0)
    sv = list(range(1, 13))
    n = 1 << len(sv)
    b = api.Batch(db, [api.Point(sv, n - 5, api.MODE_ENUM, first_sample=5)],
                  api.OUT_FLAGS | api.OUT_TRAIL | api.OUT_ASSIGN, max_models=n)
    b.run()
    conflict, full = b.flags(0)
    asg = b.assign(0)
    n_full = 0
    for s in range(n - 5):
        a = oracle_mod.enum_assumptions(sv, s + 5)
        r, ea = P.bcp(list(a))
        assert bool(conflict[s]) == bool(r.conflict), s
        if not r.conflict:
            assert (asg[s] == ea).all(), s
            assert bool(full[s]) == bool(r.full)
            n_full += r.full
    found, _, smp, models = b.models()
    assert found == n_full
    for i in range(len(smp)):
        m = models[i]
        for cl in clauses:  # CheckSATset (src_mpi/mpi_base.cpp:717-752): model satisfies every clause
            assert any((m[abs(l) - 1] == 1) == (l > 0) for l in cl)


def test_gpu_mt19937_stream(oracle_mod, cuda_lib):
    from pdsat_b200 import api
    for seed, first, n in [(1, 0, 5000), (5489, 9990, 20), (7, 623, 1300), (0, 1248, 1248),
                           (3, 624 * 1000 + 17, 70000), (11, 150_000_007, 4000)]:
        got = api.mt19937_bool_rand(seed, first, n)
        exp = oracle_mod.bool_rand(seed, first, n)
        assert (got == exp).all(), (seed, first, n)


def test_gpu_counter_mode_declared_generator(bivium, golden, oracle_mod, cuda_lib):
    """The counter generator is a declared deviation from the reference stream; the oracle is
    fed the same bits (include/pdsat_cuda.h PDSAT_MODE_RANDOM_COUNTER)."""
    from pdsat_b200 import api
    nv, _, offs, lits = bivium
    stream = list(golden["stream_lits"])
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    P.add_units(stream)
    sv = [177 - i for i in range(86)]
    n, seed, first = 130, 12345, 128

    def splitmix64(x):
        x = (x + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        return x ^ (x >> 31)

    b = api.Batch(db, [api.Point(sv, n, api.MODE_COUNTER, seed=seed, first_sample=first)], api.OUT_FLAGS | api.OUT_TRAIL)
    b.run()
    conflict, _ = b.flags(0)
    trail = b.trail(0)
    for s in range(n):
        g = (first + s) // 64
        a = []
        for j, v in enumerate(sv):
            w = splitmix64((seed + 0x9E3779B97F4A7C15 * (g * 65536 + j + 1)) & (2 ** 64 - 1))
            bit = (w >> ((first + s) % 64)) & 1
            a.append(2 * (v - 1) + (0 if bit else 1))
        r, _ = P.bcp(a)
        assert bool(conflict[s]) == bool(r.conflict)
        if not r.conflict:
            assert trail[s] == r.trail_size


def test_gpu_mt19937_multi_segment_batch(bivium, golden, oracle_mod, cuda_lib):
    """A long single-point stream is cut into jump-ahead segments; the sample values must
    still be the sequential mt19937 stream (1-worker order, mpi_predicter.cpp:3236-3242)."""
    from pdsat_b200 import api
    nv, _, offs, lits = bivium
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(list(golden["stream_lits"]))
    sv = list(range(148, 178))
    n, first = 300_000, 12_345
    b = api.Batch(db, [api.Point(sv, n, api.MODE_MT19937, seed=2, first_sample=first),
                       api.Point(sv[:7], 5000, api.MODE_MT19937, seed=3)], api.OUT_ASSIGN)
    b.run()
    for lo in (0, 63, 150_001, n - 40):
        asg = b.assign(0, lo, 40)
        A = oracle_mod.sample_assumptions(2, sv, first + lo, 40)
        for s in range(40):
            for j, v in enumerate(sv):
                assert asg[s][v - 1] == (A[s][j] & 1), (lo, s, j)
    # reseed to another window of the same stream without reallocating
    b.reseed(0, 5, 777)
    b.run()
    asg = b.assign(0, n - 10, 10)
    A = oracle_mod.sample_assumptions(5, sv, 777 + n - 10, 10)
    for s in range(10):
        for j, v in enumerate(sv):
            assert asg[s][v - 1] == (A[s][j] & 1)
    # the same offset again with other seeds: from its second appearance the offset is reached
    # with ONE jump by the composite polynomial x^(624*R0) instead of one jump per set bit
    for seed in (6, 7, 8):
        b.reseed(0, seed, 777)
        b.run()
        for lo in (0, n - 10):
            asg = b.assign(0, lo, 10)
            A = oracle_mod.sample_assumptions(seed, sv, 777 + lo, 10)
            for s in range(10):
                for j, v in enumerate(sv):
                    assert asg[s][v - 1] == (A[s][j] & 1), (seed, lo, s, j)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_gpu_random_cnfs(seed, oracle_mod, cuda_lib):
    """Random CNFs with mixed clause lengths (binary ... 12 literals, i.e. chained records),
    unit clauses at level 0, random sets: every enumerated assignment against the oracle.
    Exercises the first-wave candidate analysis and the queue-driven propagation on
    structures unlike Bivium."""
    from pdsat_b200 import api, fixtures as fx
    rng = np.random.default_rng(100 + seed)
    nv = int(rng.integers(20, 60))
    clauses = []
    for _ in range(int(nv * rng.uniform(1.5, 3.5))):
        ln = int(rng.choice([2, 2, 3, 3, 3, 4, 5, 9, 12]))
        ln = min(ln, nv)
        vs = rng.choice(np.arange(1, nv + 1), size=ln, replace=False)
        clauses.append([int(v) * (1 if rng.integers(0, 2) else -1) for v in vs])
    for _ in range(int(rng.integers(0, 3))):
        clauses.append([int(rng.integers(1, nv + 1)) * (1 if rng.integers(0, 2) else -1)])
    offs, lits = fx.to_csr(clauses)
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    db = api.ClauseDb(nv, offs, lits, device=0)
    if db.info().status != api.PDSAT_OK:
        assert not P.ok()
        return
    pts, sets = [], []
    for _ in range(4):
        k = int(rng.integers(1, 11))
        sv = sorted(int(v) for v in rng.choice(np.arange(1, nv + 1), size=k, replace=False))
        sets.append(sv)
        pts.append(api.Point(sv, 1 << k, api.MODE_ENUM))
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL | api.OUT_ASSIGN, max_models=4096)
    costs = b.run()
    total_full = 0
    for p, sv in enumerate(sets):
        conflict, full = b.flags(p)
        trail = b.trail(p)
        asg = b.assign(p)
        for s in range(1 << len(sv)):
            r, a = P.bcp(list(oracle_mod.enum_assumptions(sv, s)))
            assert bool(conflict[s]) == bool(r.conflict), (p, s)
            if not r.conflict:
                assert trail[s] == r.trail_size and bool(full[s]) == bool(r.full), (p, s)
                assert (asg[s] == a).all(), (p, s)
                total_full += r.full
    found, _, _, models = b.models()
    assert found == total_full
    for m in models:
        for cl in clauses:
            assert any((m[abs(l) - 1] == 1) == (l > 0) for l in cl)


@pytest.mark.parametrize("case", ["n161_k16", "n100_k14", "n60_k12", "n150_k10", "n85_k12", "n90_k12"])
def test_gpu_enumeration_vs_reference_cube(case, bivium, golden, cuda_lib):
    """solve-mode enumeration of a cube.  Golden vectors come from the
    reference itself (tests/golden/make_golden.py): (1) every assignment of the cube through the
    reference's propagate() - the device must leave exactly those conflict-free; (2) the list
    reported by Solver::gen_valid_assumptions_rc2.  rc2 equals (1) on the cubes SURVEY.md checked,
    but on cubes with deep conflicts it OVER-reports (after a conflict below the last level it
    flips a deeper assumption without undoing the conflicting level, Solver.cc:1399-1423, so the
    pending conflict is never seen again): there it must still contain every true survivor.
    rc2 orders with d_set[0] most significant; ENUM mode has bit r <-> set_vars[r]."""
    import os
    from pdsat_b200 import api
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rc2_golden.npz"))
    n_fixed, k = (int(x[1:]) for x in case.split("_"))
    rc2_valid = int(g[case + "_valid"][0])
    rows = np.unpackbits(g[case + "_rows"], axis=1)[:, :k]
    bcp_surv = set(int(x) for x in g[case + "_bcp_survivors"])
    nv, _, offs, lits = bivium
    secret = golden["secret"]
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(list(golden["stream_lits"]) + [2 * v + (0 if secret[v] else 1) for v in range(n_fixed)])
    d_set = list(range(178 - k, 178))
    b = api.Batch(db, [api.Point(d_set, 1 << k, api.MODE_ENUM)], api.OUT_FLAGS)
    c = b.run()[0]
    conflict, full = b.flags(0)
    survivors = set(int(i) for i in np.nonzero(~conflict)[0])
    assert survivors == bcp_surv and c.n - c.n_conflict == len(bcp_surv)
    rc2 = set(sum(int(r[j]) << j for j in range(k)) for r in rows)
    assert len(rc2) == rc2_valid and survivors <= rc2
    if case in ("n161_k16", "n100_k14", "n60_k12", "n150_k10"):
        assert survivors == rc2


def test_gpu_everything_fixed_at_level0(oracle_mod, cuda_lib):
    """No live variable at all: set variables are all level-0 fixed; a matching assignment is a
    (trivially) full model, a contradicting one a conflict."""
    from pdsat_b200 import api, fixtures as fx
    clauses = [[1], [2], [-1, 3]]
    offs, lits = fx.to_csr(clauses)
    db = api.ClauseDb(3, offs, lits, device=0)
    assert db.info().n_live_vars == 0 and db.info().status == api.PDSAT_OK
    P = oracle_mod.Oracle(3, offs, lits, "port")
    b = api.Batch(db, [api.Point([1, 3], 4, api.MODE_ENUM), api.Point([2], 1, api.MODE_ENUM, first_sample=1)],
                  api.OUT_FLAGS | api.OUT_TRAIL, max_models=8)
    costs = b.run()
    conflict, full = b.flags(0)
    for s in range(4):
        r, _ = P.bcp(list(oracle_mod.enum_assumptions([1, 3], s)))
        assert bool(conflict[s]) == bool(r.conflict) and bool(full[s]) == bool(r.full)
    assert (costs[0].n_conflict, costs[0].n_full, costs[0].sum_trail) == (3, 1, 3)
    assert (costs[1].n, costs[1].n_conflict, costs[1].n_full) == (1, 0, 1)
    found, _, _, models = b.models()
    assert found == 2 and all(list(m) == [1, 1, 1] for m in models)


@pytest.mark.parametrize("k", [30, 85])
def test_config0_predict_estimate_matches_reference_path(k, bivium, golden, oracle_mod, cuda_lib):
    """BASELINE.json configs[0]: 1 000-sample Monte Carlo predict of a start-state set.  The
    reference path (fresh solver per sample, cost = Solver::propagations, GetPredict arithmetic in
    long double) and the device path (integer sums -> host estimator) must give the same runtime
    estimate within 1e-12 relative (north star); observed: identical."""
    from pdsat_b200 import api, predicter as pr
    nv, _, offs, lits = bivium
    stream = list(golden["stream_lits"])
    sv = pr.schema_vars("bivium_Ending2", k)
    n = 1000
    P = oracle_mod.Oracle(nv, offs, lits, "ref" if oracle_mod.have_ref() else "port")
    A = oracle_mod.sample_assumptions(1, sv, 0, n)
    cost = np.empty(n)
    for s in range(n):
        # eval "prep": the reference stops after BCP (Solver.cc:874-877); with "propagation" it
        # would run the full CDCL search.  Solver::propagations is the same BCP count either way.
        r, _ = P.predict_sample(list(A[s]) + stream, 0, want_assigns=False)
        assert not r.conflict
        cost[s] = float(r.propagations)
    ref_sum, ref_med, ref_pred = oracle_mod.predict_mean(cost, k, proc_count=4)
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    res = pr.Predicter(db, cnf_in_set_count=n, evaluation_type="bcp_trail", proc_count=4, base_seed=1).evaluate([sv])[0]
    assert res.sum_cost == ref_sum
    assert abs(float((res.predict - np.longdouble(ref_pred)) / np.longdouble(ref_pred))) <= 1e-12
    assert abs(pr.sample_variance(res.cost) - (oracle_mod.sample_variance(cost) if cost.std() > 0 else 0.0)) <= \
        1e-12 * max(1.0, oracle_mod.sample_variance(cost))


@pytest.mark.parametrize("k", [86, 88, 100])
def test_estimates_on_sets_with_conflicts(k, bivium, golden, oracle_mod, cuda_lib):
    """What the two estimates are on sets where BCP hits conflicts (round 1 only checked
    conflict-free sets).  "prep": getCurPredictTime (src_mpi/mpi_predicter.cpp:1846-1901) with
    solved = samples BCP decides - reference arithmetic on the reference's own per-sample prep flags.
    "bcp_trail": GetPredict (:1976-1996) on the costs  trail(sample) for conflict-free samples, 0 for
    conflicts (stated deviation: the reference's count at a conflict is order dependent)."""
    from pdsat_b200 import api, predicter as pr
    nv, _, offs, lits = bivium
    stream = list(golden["stream_lits"])
    sv = pr.schema_vars("bivium_Ending2", k)
    n = 2000
    P = oracle_mod.Oracle(nv, offs, lits, "ref" if oracle_mod.have_ref() else "port")
    P.add_units(stream)
    conflict, full, trail, sums = P.batch_mt(0, 1, sv, n)
    cost = np.where(conflict, 0.0, trail.astype(np.float64))
    ref_sum, ref_med, ref_pred = oracle_mod.predict_mean(cost, k, proc_count=2)
    solved = int(conflict.sum() + full.sum())
    ref_prep = oracle_mod.predict_prep(solved, n, k)
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    res = pr.Predicter(db, cnf_in_set_count=n, evaluation_type="bcp_trail", proc_count=2, base_seed=1).evaluate([sv])[0]
    assert res.cost.n_conflict == int(conflict.sum()) > 0
    assert res.sum_cost == ref_sum
    if ref_pred > 0:
        assert abs(float((res.predict - np.longdouble(ref_pred)) / np.longdouble(ref_pred))) <= 1e-12
    else:
        assert res.predict == 0
    res = pr.Predicter(db, cnf_in_set_count=n, evaluation_type="prep", base_seed=1).evaluate([sv])[0]
    assert res.solved == solved
    assert abs(float((res.predict - np.longdouble(ref_prep)) / np.longdouble(ref_prep))) <= 1e-12
