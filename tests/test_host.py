"""CPU: fixtures, the C-ABI surface, and the host half of the clause-database upload
(no kernel is launched here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_bivium_template_shape(bivium):
    nv, cl, offs, lits = bivium
    assert nv == 777 and len(cl) == 12800 and int(offs[-1]) == 67200
    assert sum(1 for c in cl if len(c) == 5) == 9600 and sum(1 for c in cl if len(c) == 6) == 3200


def test_bivium_template_equals_reference(bivium, tmp_path):
    """The generated template must be the reference's bivium_template.cnf, line for line."""
    path = os.path.join(REF, "src_common", "bivium_template.cnf")
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    from pdsat_b200 import fixtures as fx
    nv, cl, comments = fx.bivium_template()
    out = tmp_path / "b.cnf"
    fx.write_dimacs(str(out), nv, cl, comments)
    mine = [l.strip() for l in open(out) if l.strip()]
    ref = [l.strip() for l in open(path, encoding="latin-1") if l.strip()]
    assert mine == ref
    nv2, cl2, meta = fx.read_dimacs(path)
    assert nv2 == 777 and cl2 == cl and meta == {"input_variables": 177, "output_variables": 200}


def test_odls_counts():
    from pdsat_b200 import fixtures as fx
    nv, cl = fx.odls_pair_cnf(10)
    assert nv == 2000 and len(cl) == 434440 and sum(map(len, cl)) == 1684000


def test_abi_exports_every_declared_symbol(cuda_lib):
    hdr = open(os.path.join(ROOT, "include", "pdsat_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pdsat_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    from pdsat_b200 import api
    assert declared == set(api.EXPORTS)
    for name in declared:
        assert hasattr(cuda_lib, name), name


def test_upload_host_only_matches_oracle_level0(bivium, golden, oracle_mod, cuda_lib):
    from pdsat_b200 import api
    nv, _, offs, lits = bivium
    db = api.ClauseDb(nv, offs, lits, device=-1)
    i = db.info()
    assert (i.n_vars, i.n_live_vars, i.n_clauses_live, i.n_lits_live, i.status) == (777, 777, 12800, 67200, 0)
    stream = golden["stream_lits"]
    db.set_fixed_assumptions(stream)
    i = db.info()
    assert i.n_base_assigned == 200 and i.n_live_vars == 577 and i.status == api.PDSAT_OK
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    r, a = P.bcp(list(stream))
    assert (a == db.base_assignment()).all()
    # contradictory fixed assumptions: level-0 refutation (Solver::ok == false)
    db.set_fixed_assumptions([0, 1])
    assert db.info().status == api.PDSAT_ERR_UNSAT
    db.set_fixed_assumptions([])
    assert db.info().n_live_vars == 777


def test_upload_normalisation(cuda_lib, oracle_mod):
    """addClause_ rules (Solver.cc:279-307): duplicates, tautologies, units, empty clause."""
    from pdsat_b200 import api, fixtures as fx
    clauses = [[1, 1, 2], [3, -3, 4], [5], [-5, 6], [-6, 7, 8], [2, 9, 9, -10]]
    offs, lits = fx.to_csr(clauses)
    db = api.ClauseDb(10, offs, lits, device=-1)
    i = db.info()
    base = db.base_assignment()
    P = oracle_mod.Oracle(10, offs, lits, "port")
    r, a = P.bcp([])
    assert (a == base).all()
    assert list(base[[4, 5]]) == [0, 0] and i.n_base_assigned == 2
    assert i.n_clauses_live == 3  # [1 2], [7 8], [2 9 -10]
    offs2, lits2 = fx.to_csr([[1], [-1]])
    db2 = api.ClauseDb(1, offs2, lits2, device=-1)
    assert db2.info().status == api.PDSAT_ERR_UNSAT


def test_no_cpu_fallback(bivium, cuda_lib):
    """Compute entry points must fail loudly without a device."""
    from pdsat_b200 import api
    nv, _, offs, lits = bivium
    db = api.ClauseDb(nv, offs, lits, device=-1)
    with pytest.raises(api.PdsatError) as e:
        db.eval_batch([api.Point(list(range(148, 178)), 64)])
    assert e.value.code == api.PDSAT_ERR_CUDA
    if api.device_count() == 0:
        with pytest.raises(api.PdsatError):
            api.ClauseDb(nv, offs, lits, device=0)
        with pytest.raises(api.PdsatError):
            api.mt19937_bool_rand(1, 0, 10)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pdsat_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU oracle", "").replace("the CPU oracle", ""), f
