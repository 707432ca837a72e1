"""CPU: pins the plain-C restatement (oracle/bcp_oracle.c) against the golden vectors that
were produced by the unmodified reference MiniSat (tests/golden/make_golden.py), and
directly against oracle/_ref when it is present."""
import numpy as np
import pytest

CASES = ["ending2_k30", "beginning1_k30", "ending2_k85", "ending2_k86", "ending2_k90", "ending2_k100",
         "ending2_k120", "ending2_k177", "nostream_k177", "nostream_k150"]
TE_CASES = ["te_k80", "te_k86", "te_k88", "te_k100"]


def _assign_planes(asg):
    return np.packbits(asg == 0, axis=1), np.packbits(asg == 1, axis=1)


@pytest.mark.parametrize("case", CASES)
def test_port_matches_reference_golden(case, bivium, golden, oracle_mod):
    nv, _, offs, lits = bivium
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    sv = golden[case + "_set"]
    seed, with_ks = golden[case + "_seed"]
    rows = golden[case + "_rows"]
    stream = list(golden["stream_lits"]) if with_ks else []
    A = oracle_mod.sample_assumptions(int(seed), sv, 0, len(rows))
    asg = np.zeros((len(rows), nv), dtype=np.uint8)
    for s in range(len(rows)):
        r, a = P.predict_sample(list(A[s]) + stream, 0)
        got = (r.conflict, r.full, r.sat, r.prep, r.trail_size, r.propagations, r.watch_scans)
        assert got == tuple(rows[s]), (case, s)
        asg[s] = a
    t, f = _assign_planes(asg)
    assert (t == golden[case + "_assign"]).all() and (f == golden[case + "_assign_f"]).all()


@pytest.mark.parametrize("case", TE_CASES)
def test_port_matches_reference_golden_te(case, bivium, golden, oracle_mod):
    nv, _, offs, lits = bivium
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    sv = golden[case + "_set"]
    rows = golden[case + "_rows"]
    states = np.unpackbits(golden["te_states"], axis=1)[:, :177]
    streams = np.unpackbits(golden["te_streams"], axis=1)[:, :200]
    for s in range(len(rows)):
        a = [2 * (v - 1) + (0 if states[s][v - 1] else 1) for v in sv] + \
            [2 * (577 + i) + (0 if streams[s][i] else 1) for i in range(200)]
        r, _ = P.predict_sample(a, 0)
        got = (r.conflict, r.full, r.sat, r.prep, r.trail_size, r.propagations, r.watch_scans)
        assert got == tuple(rows[s]), (case, s)


def test_port_vs_ref_live(bivium, oracle_mod):
    """Direct comparison with the reference build (only where oracle/_ref exists)."""
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref not built")
    nv, _, offs, lits = bivium
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    R = oracle_mod.Oracle(nv, offs, lits, "ref")
    rng = np.random.default_rng(5)
    for trial in range(60):
        k = int(rng.integers(1, 178))
        sv = rng.choice(np.arange(1, 778), size=k, replace=False)
        a = [2 * (int(v) - 1) + int(rng.integers(0, 2)) for v in sv]
        for fn in ("predict_sample", "bcp"):
            rp, ap = getattr(P, fn)(a)
            rr, ar = getattr(R, fn)(a)
            for fld in ("conflict", "full", "trail_size", "propagations", "watch_scans"):
                assert getattr(rp, fld) == getattr(rr, fld), (trial, fn, fld)
            assert (ap == ar).all()


def test_bivium_cipher_kat(bivium, oracle_mod):
    """BCP of the template under a full 177-bit state must reproduce the
    independent cipher model on every variable (keystream = vars 578..777)."""
    from pdsat_b200 import fixtures as fx
    nv, _, offs, lits = bivium
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    for seed in range(5):
        st = oracle_mod.bool_rand(100 + seed, 0, 177)
        a = [2 * v + (0 if st[v] else 1) for v in range(177)]
        r, asg = P.bcp(a)
        assert not r.conflict and r.full and r.trail_size == 777
        trace = fx.BiviumModel(st).trace(200)
        assert ((asg == 0).astype(int) == np.asarray(trace)).all()


def test_mt19937_known_answers(oracle_mod):
    """mt19937 published known answers: seed 5489 -> first draw 3499211612, 10000th draw
    4123659995 (the C++11 standard's check value)."""
    d = oracle_mod.mt19937_draws(5489, 0, 10000)
    assert int(d[0]) == 3499211612
    assert int(d[9999]) == 4123659995
    b = oracle_mod.bool_rand(5489, 0, 16)
    assert (b == ((d[:16] >> 31) == 0)).all()


def test_enum_order(oracle_mod):
    """src_mpi/mpi_base.cpp:148-169: bit r of lint <-> r-th variable ascending, set => positive."""
    a = oracle_mod.enum_assumptions([3, 7, 20], 0b101)
    assert list(a) == [2 * 2 + 0, 2 * 6 + 1, 2 * 19 + 0]


def test_estimator_restatement(oracle_mod):
    cost = np.asarray([230.0] * 1000)
    s, med, pred = oracle_mod.predict_mean(cost, 30, 1)
    assert s == 230000.0 and float(med) == 230.0 and float(pred) == 230.0 * 2 ** 30
    assert float(oracle_mod.predict_prep(5, 1000, 30)) == 1e308       # 0.5 % is not enough
    assert float(oracle_mod.predict_prep(6, 1000, 30)) == 2.0 ** 30 * 3.0 / 0.006


@pytest.mark.parametrize("case", ["n161_k16", "n100_k14", "n60_k12", "n85_k12", "n90_k12"])
def test_port_cube_enumeration_matches_reference_propagate(case, bivium, golden, oracle_mod):
    """Solve-mode cube, assignment by assignment (order of src_mpi/mpi_base.cpp:148-169): the C
    restatement leaves exactly the assignments the reference's propagate() leaves
    (tests/golden/rc2_golden.npz, *_bcp_survivors)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rc2_golden.npz"))
    n_fixed, k = (int(x[1:]) for x in case.split("_"))
    nv, _, offs, lits = bivium
    secret = golden["secret"]
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    P.add_units(list(golden["stream_lits"]) + [2 * v + (0 if secret[v] else 1) for v in range(n_fixed)])
    d_set = list(range(178 - k, 178))
    surv = []
    for idx in range(1 << k):
        r, _ = P.bcp(oracle_mod.enum_assumptions(d_set, idx), want_assigns=False)
        if not r.conflict:
            surv.append(idx)
    assert surv == [int(x) for x in g[case + "_bcp_survivors"]]
