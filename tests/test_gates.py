"""Gate records (pdsat_b200/csrc/host_db.hpp, kernels.cuh eval_gate) are equivalent to unit
propagation over the clauses they replace.

CPU part: a Python restatement of the device's closed forms is compared, on EVERY partial
assignment (3^n), with clause-level unit propagation over the canonical clause set - same
conflict verdict, same fixpoint.  The recogniser is checked on the fixtures: every gate it reports
regenerates clauses that are all present in the CNF.
GPU part: single-gate and mixed CNFs through the C ABI against the oracle (clause-level BCP of the
reference, Solver.cc:599-662), with every subset of variables as the decomposition set and every
assignment of it (ENUM mode) - i.e. again all 3^n partial assignments, now on the device code."""
import itertools
import os

import numpy as np
import pytest

from pdsat_b200 import api, fixtures as fx

T, F, U = 1, 0, None


# ----------------------------------------------------------------------------- canonical CNFs
def xor_clauses(xs, p):
    """x_1 ^ ... ^ x_m = p: forbid every point of parity p ^ 1."""
    out = []
    for pt in itertools.product((0, 1), repeat=len(xs)):
        if sum(pt) % 2 != p:
            out.append([-x if v else x for x, v in zip(xs, pt)])  # literal false <=> x == v
    return out


def andxor_clauses(xs, lb, lc, p):
    """xor(xs) ^ (lb & lc) = p, lb / lc DIMACS literals: (lb | [xor = p]), (lc | [xor = p]),
    (-lb | -lc | [xor = p ^ 1])."""
    out = []
    for cl in xor_clauses(xs, p):
        out.append(cl + [lb])
        out.append(cl + [lc])
    for cl in xor_clauses(xs, p ^ 1):
        out.append(cl + [-lb, -lc])
    return out


# ----------------------------------------------------------------------------- clause-level UP
def lit_val(a, l):
    v = a.get(abs(l))
    if v is None:
        return None
    return v if l > 0 else 1 - v


def up_fixpoint(clauses, a):
    """Unit propagation to the fixpoint; returns (conflict, assignment)."""
    a = dict(a)
    changed = True
    while changed:
        changed = False
        for cl in clauses:
            vals = [lit_val(a, l) for l in cl]
            if any(v == 1 for v in vals):
                continue
            und = [l for l, v in zip(cl, vals) if v is None]
            if not und:
                return True, a
            if len(und) == 1:
                a[abs(und[0])] = 1 if und[0] > 0 else 0
                changed = True
    return False, a


# ----------------------------------------------------------------------------- device closed form
def gate_step(kind, xs, lb, lc, p, a):
    """One evaluation of eval_gate (kernels.cuh) on one sample; returns (conflict, implied dict)."""
    par = p
    unknown = []
    for x in xs:
        v = a.get(x)
        if v is None:
            unknown.append(("x", x))
        else:
            par ^= v
    if kind == "andxor":
        vb, vc = lit_val(a, lb), lit_val(a, lc)
        q1 = vb == 1 and vc == 1
        q0 = vb == 0 or vc == 0
        if q1:
            par ^= 1
        if not (q1 or q0):
            unknown.append(("q", None))
    if not unknown:
        return par == 1, {}
    if len(unknown) != 1:
        return False, {}
    imp = {}
    what, x = unknown[0]
    if what == "x":
        imp[x] = par
    else:
        vb, vc = lit_val(a, lb), lit_val(a, lc)
        if par == 1:  # product must be 1
            if vb is None:
                imp[abs(lb)] = 1 if lb > 0 else 0
            if vc is None:
                imp[abs(lc)] = 1 if lc > 0 else 0
        else:  # product must be 0
            if vb == 1:
                imp[abs(lc)] = 0 if lc > 0 else 1
            if vc == 1:
                imp[abs(lb)] = 0 if lb > 0 else 1
    return False, imp


def gate_fixpoint(kind, xs, lb, lc, p, a):
    a = dict(a)
    while True:
        conflict, imp = gate_step(kind, xs, lb, lc, p, a)
        if conflict:
            return True, a
        new = {v: val for v, val in imp.items() if a.get(v) is None}
        if not new:
            return False, a
        a.update(new)


def all_partial(vars_):
    for vals in itertools.product((None, 0, 1), repeat=len(vars_)):
        yield {v: x for v, x in zip(vars_, vals) if x is not None}


@pytest.mark.parametrize("m", [2, 3, 4, 5, 6, 7])
def test_xor_closed_form_equals_unit_propagation(m):
    xs = list(range(1, m + 1))
    for p in (0, 1):
        cls = xor_clauses(xs, p)
        assert len(cls) == 2 ** (m - 1)
        for a in all_partial(xs):
            assert up_fixpoint(cls, a) == gate_fixpoint("xor", xs, None, None, p, a), (m, p, a)


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5])
def test_andxor_closed_form_equals_unit_propagation(m):
    xs = list(range(1, m + 1))
    b, c = m + 1, m + 2
    for p in (0, 1):
        for lb in (b, -b):
            for lc in (c, -c):
                cls = andxor_clauses(xs, lb, lc, p)
                assert len(cls) == 3 * 2 ** (m - 1)
                for a in all_partial(xs + [b, c]):
                    cu, au = up_fixpoint(cls, a)
                    cg, ag = gate_fixpoint("andxor", xs, lb, lc, p, a)
                    assert cu == cg, (m, p, lb, lc, a)
                    if not cu:
                        assert au == ag, (m, p, lb, lc, a)


def test_bivium_and_xor_gate_is_canonical():
    """The template's 24-clause block y = a ^ b&c ^ d ^ e is ANDXOR-4 with parity 0."""
    got = sorted(sorted(c) for c in fx._and_xor_gate(6, 1, 2, 3, 4, 5))
    want = sorted(sorted(c) for c in andxor_clauses([6, 1, 4, 5], 2, 3, 0))
    assert got == want


# ----------------------------------------------------------------------------- recogniser (host)
def _gate_clauses(g):
    xs = [g.x[i] for i in range(g.m)]
    if g.type == 0:
        return xor_clauses(xs, g.parity)
    lb = (g.lb >> 1) + 1 if (g.lb & 1) == 0 else -((g.lb >> 1) + 1)
    lc = (g.lc >> 1) + 1 if (g.lc & 1) == 0 else -((g.lc >> 1) + 1)
    return andxor_clauses(xs, lb, lc, g.parity)


def _simplify(clauses, base):
    """clause set after level-0 simplification with lbool bytes `base` (0 true, 1 false, 2 undef)"""
    out = set()
    for cl in clauses:
        keep = []
        sat = False
        for l in cl:
            v = base[abs(l) - 1]
            if v == 2:
                keep.append(l)
            elif (v == 0) == (l > 0):
                sat = True
                break
        if not sat:
            out.add(tuple(sorted(keep)))
    return out


@pytest.mark.parametrize("fold_keystream", [False, True])
def test_recogniser_on_bivium(fold_keystream, bivium, cuda_lib):
    nv, cl, offs, lits = bivium
    db = api.ClauseDb(nv, offs, lits, device=-1)
    if fold_keystream:
        import bench
        db.set_fixed_assumptions(bench.instance()[4])
    info = db.info()
    gates = db.gates()
    assert info.n_gates == 600 == len(gates)
    assert info.n_gate_clauses == info.n_clauses_live  # no plain clause left
    live = _simplify(cl, db.base_assignment())
    covered = set()
    for g in gates:
        gc = [tuple(sorted(c)) for c in _gate_clauses(g)]
        assert len(gc) == g.n_clauses
        for c in gc:
            assert c in live and c not in covered
            covered.add(c)
    assert covered == live
    db.close()


def test_recogniser_on_a52_and_odls(cuda_lib):
    nv, cl, _, _ = fx.a5_2_cnf(20)
    db = api.ClauseDb.from_clauses(nv, cl, device=-1)
    live = _simplify(cl, db.base_assignment())
    covered = set()
    for g in db.gates():
        for c in _gate_clauses(g):
            c = tuple(sorted(c))
            assert c in live and c not in covered
            covered.add(c)
    assert len(covered) == db.info().n_gate_clauses > 0
    db.close()
    onv, ocl = fx.odls_pair_cnf(4)
    odb = api.ClauseDb.from_clauses(onv, ocl, device=-1)
    assert odb.info().n_gates == 0  # ALO / AMO / orthogonality clauses: nothing xor-like
    odb.close()


def test_no_gates_switch(bivium, cuda_lib, monkeypatch):
    nv, cl, offs, lits = bivium
    monkeypatch.setenv("PDSAT_NO_GATES", "1")
    db = api.ClauseDb(nv, offs, lits, device=-1)
    assert db.info().n_gates == 0 and db.info().n_gate_clauses == 0
    db.close()


# ----------------------------------------------------------------------------- device (GPU)
def _enum_all_subsets(api, orc, n_vars, clauses, max_set=None):
    """every non-empty subset of the variables as the set, every assignment of it: device vs oracle"""
    offs, lits = fx.to_csr(clauses)
    db = api.ClauseDb(n_vars, offs, lits, device=0)
    P = orc.Oracle(n_vars, offs, lits, "port")
    subsets = [s for r in range(1, n_vars + 1) for s in itertools.combinations(range(1, n_vars + 1), r)]
    if max_set:
        subsets = subsets[:max_set]
    pts = [api.Point(list(s), 1 << len(s), api.MODE_ENUM) for s in subsets]
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_ASSIGN | api.OUT_TRAIL)
    b.run()
    n_checked = 0
    for p, s in enumerate(subsets):
        conflict, full = b.flags(p)
        asg = b.assign(p)
        trail = b.trail(p)
        for lint in range(1 << len(s)):
            assum = [2 * (v - 1) + (0 if (lint >> r) & 1 else 1) for r, v in enumerate(s)]
            r, a = P.bcp(assum)
            assert bool(conflict[lint]) == bool(r.conflict), (s, lint)
            if not r.conflict:
                assert (asg[lint] == a).all(), (s, lint)
                assert trail[lint] == r.trail_size
            n_checked += 1
    b.close()
    db.close()
    return n_checked


@pytest.mark.gpu
@pytest.mark.parametrize("gate", ["xor3_0", "xor5_1", "xor7_0", "andxor1", "andxor3", "andxor4_neg", "andxor5"])
def test_gpu_single_gate_all_partial_assignments(gate, oracle_mod, cuda_lib):
    if gate.startswith("xor"):
        m, p = int(gate[3]), int(gate[-1])
        cls, nv = xor_clauses(list(range(1, m + 1)), p), m
    else:
        m = int(gate[6])
        neg = gate.endswith("neg")
        xs = list(range(1, m + 1))
        cls, nv = andxor_clauses(xs, -(m + 1) if neg else m + 1, m + 2, 1 if neg else 0), m + 2
    db = api.ClauseDb.from_clauses(nv, cls, device=-1)
    assert db.info().n_gates == 1 and db.info().n_gate_clauses == len(cls)
    db.close()
    assert _enum_all_subsets(api, oracle_mod, nv, cls) == 3 ** nv - 1


@pytest.mark.gpu
def test_gpu_two_gates_imply_opposite_values(oracle_mod, cuda_lib):
    """x2 = x1 (XOR) and x3 ^ (-x2 & x4) = 0 (ANDXOR reading x2 through its negative literal):
    with x1 = 1, x3 = 1 the first gate makes x2 true and the second makes it false in the SAME
    level; each gate alone then sees a value that suits it.  The reference hits a conflict (the
    second reason clause is falsified) - so must the device (explicit true-and-false check)."""
    cls = xor_clauses([1, 2], 0) + andxor_clauses([3], -2, 4, 0)
    db = api.ClauseDb.from_clauses(4, cls, device=-1)
    assert db.info().n_gates == 2 and db.info().n_gate_clauses == len(cls)
    db.close()
    for _ in range(20):  # the order of the two implications is a race: repeat
        assert _enum_all_subsets(api, oracle_mod, 4, cls) == 3 ** 4 - 1


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_gpu_mixed_gates_and_clauses_vs_oracle(seed, oracle_mod, cuda_lib):
    """random circuits: xor / and-xor gates sharing variables + random plain clauses"""
    orc = oracle_mod
    rng = np.random.default_rng(100 + seed)
    nv = 60
    cls = []
    for _ in range(40):
        m = int(rng.integers(2, 6))
        vs = [int(v) + 1 for v in rng.choice(nv, size=m + 2, replace=False)]
        if rng.random() < 0.5:
            cls += xor_clauses(vs[:m], int(rng.integers(0, 2)))
        else:
            lb = vs[m] * (1 if rng.random() < 0.5 else -1)
            lc = vs[m + 1] * (1 if rng.random() < 0.5 else -1)
            cls += andxor_clauses(vs[:m], lb, lc, int(rng.integers(0, 2)))
    for _ in range(60):
        ln = int(rng.integers(2, 5))
        vs = rng.choice(nv, size=ln, replace=False)
        cls.append([int(v + 1) * (1 if rng.random() < 0.5 else -1) for v in vs])
    offs, lits = fx.to_csr(cls)
    db = api.ClauseDb(nv, offs, lits, device=0)
    info = db.info()
    assert info.n_gates > 0 and info.n_gate_clauses < info.n_clauses_live  # both paths in one kernel
    P = orc.Oracle(nv, offs, lits, "port")
    pts = []
    for _ in range(24):
        k = int(rng.integers(4, 40))
        sv = sorted(int(v) + 1 for v in rng.choice(nv, size=k, replace=False))
        pts.append(api.Point(sv, 200, api.MODE_MT19937, seed=int(rng.integers(1, 1000))))
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_ASSIGN | api.OUT_TRAIL)
    costs = b.run()
    for p, pt in enumerate(pts):
        A = orc.sample_assumptions(pt.seed, pt.set_vars, 0, pt.n_samples)
        conflict, full = b.flags(p)
        asg = b.assign(p)
        nconf = 0
        for s in range(pt.n_samples):
            r, a = P.bcp(list(A[s]))
            assert bool(conflict[s]) == bool(r.conflict), (p, s)
            nconf += r.conflict
            if not r.conflict:
                assert (asg[s] == a).all(), (p, s)
                assert bool(full[s]) == bool(r.full)
        assert costs[p].n_conflict == nconf
    b.close()
    db.close()


@pytest.mark.gpu
def test_gpu_gate_path_equals_clause_path(bivium, cuda_lib, monkeypatch):
    """same cascade-heavy Bivium batch with and without gate extraction: identical accumulators"""
    import bench
    nv, cl, offs, lits = bivium
    stream = bench.instance()[4]
    res = []
    for no_gates in (False, True):
        if no_gates:
            monkeypatch.setenv("PDSAT_NO_GATES", "1")
        else:
            monkeypatch.delenv("PDSAT_NO_GATES", raising=False)
        db = api.ClauseDb(nv, offs, lits, device=0)
        db.set_fixed_assumptions(stream)
        assert (db.info().n_gates == 0) == no_gates
        pts = [api.Point([177 - i for i in range(k)], 20000, api.MODE_MT19937, seed=3) for k in (84, 86, 88, 92, 120, 177)]
        b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL)
        costs = b.run()
        res.append(([(c.n, c.n_conflict, c.n_full, c.sum_trail, c.sumsq_trail, c.n_implied_any) for c in costs],
                    [b.flags(p)[0].copy() for p in range(len(pts))], [b.trail(p).copy() for p in range(len(pts))]))
        b.close()
        db.close()
    assert res[0][0] == res[1][0]
    for a, bb in zip(res[0][1], res[1][1]):
        assert (a == bb).all()
    for a, bb in zip(res[0][2], res[1][2]):
        assert (a == bb).all()
