"""Solve-mode enumeration with BCP pruning (pdsat_enumerate) against brute-force enumeration on the
device (PDSAT_MODE_ENUM, itself pinned to the reference's propagate() survivors in
tests/test_gpu_parity.py) and against the CPU oracle on small cubes."""
import numpy as np
import pytest

from pdsat_b200 import api, fixtures as fx

pytestmark = pytest.mark.gpu


def _weakened_bivium(bivium, n_fixed):
    import bench
    nv, cl, offs, lits = bivium
    stream = bench.instance()[4]
    secret = bench.mt19937_bool_rand_numpy(bench.SECRET_SEED, 177)
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(list(stream) + [2 * v + (0 if secret[v] else 1) for v in range(n_fixed)])
    return db, secret


def _brute(db, sv):
    b = api.Batch(db, [api.Point(sv, 1 << len(sv), api.MODE_ENUM)], api.OUT_FLAGS)
    c = b.run()[0]
    conflict, full = b.flags(0)
    b.close()
    return np.nonzero(~conflict)[0].astype(np.uint64), full[~conflict], c


@pytest.mark.parametrize("n_fixed,k", [(157, 20), (161, 16), (100, 14), (60, 12), (165, 12), (172, 5), (170, 7)])
def test_enumerate_equals_brute_force_on_weakened_bivium(n_fixed, k, bivium, cuda_lib):
    db, secret = _weakened_bivium(bivium, n_fixed)
    sv = list(range(177 - k + 1, 178))
    surv, ism, st = db.enumerate(sv)
    want, want_full, c = _brute(db, sv)
    assert (surv == want).all() and (ism == want_full).all()
    assert st.n_assignments == 1 << k
    assert st.n_conflict == c.n_conflict and st.n_models == c.n_full
    assert st.n_undecided == (1 << k) - c.n_conflict - c.n_full
    if n_fixed + k == 177 and n_fixed >= 100:
        # the single survivor is the secret state's bits
        expect = sum(int(secret[n_fixed + r]) << r for r in range(k))
        assert list(surv) == [expect] and ism[0]
    db.close()


def test_enumerate_prunes_and_splits_by_top_bits(bivium, cuda_lib):
    """k = 26 over a heavily weakened instance: far fewer samples are propagated than 2^26; the
    union of the 8 top-3-bit shares (the 8-GPU task split, mpi_solver.cpp:447-473) is the whole."""
    db, secret = _weakened_bivium(bivium, 151)
    sv = list(range(152, 178))
    surv, ism, st = db.enumerate(sv)
    expect = sum(int(secret[151 + r]) << r for r in range(26))
    assert list(surv) == [expect] and ism[0]
    assert st.n_assignments == 1 << 26 and st.n_conflict == (1 << 26) - 1
    assert st.n_propagated < (1 << 26) // 4
    parts, n_conf = [], 0
    for g in range(8):
        s_g, m_g, st_g = db.enumerate(sv, n_fixed_top=3, fixed_top_value=g)
        assert st_g.n_assignments == 1 << 23
        parts += list(s_g)
        n_conf += st_g.n_conflict
    assert parts == [expect] and n_conf == (1 << 26) - 1
    # first-model stop
    s1, m1, st1 = db.enumerate(sv, stop_at_first_model=True)
    assert list(s1) == [expect]
    db.close()


@pytest.mark.parametrize("seed", [0, 1])
def test_enumerate_random_cnf_vs_oracle(seed, oracle_mod, cuda_lib):
    orc = oracle_mod
    rng = np.random.default_rng(40 + seed)
    nv = 40
    cls = []
    for _ in range(90):
        ln = int(rng.integers(2, 4))
        vs = rng.choice(nv, size=ln, replace=False)
        cls.append([int(v + 1) * (1 if rng.random() < 0.5 else -1) for v in vs])
    offs, lits = fx.to_csr(cls)
    db = api.ClauseDb(nv, offs, lits, device=0)
    P = orc.Oracle(nv, offs, lits, "port")
    sv = sorted(int(v) + 1 for v in rng.choice(nv, size=13, replace=False))
    surv, ism, st = db.enumerate(sv)
    want = []
    for lint in range(1 << 13):
        r, _ = P.bcp(list(orc.enum_assumptions(sv, lint)), want_assigns=False)
        if not r.conflict:
            want.append((lint, bool(r.full)))
    assert [(int(a), bool(b)) for a, b in zip(surv, ism)] == want
    assert st.n_conflict == (1 << 13) - len(want)
    db.close()
