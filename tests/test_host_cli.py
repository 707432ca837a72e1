"""The C++ host driver (pdsat_b200/host/): CPU self-tests of the estimator / parser, and on the
GPU the predict and solve modes against the Python mirror and the reference's own KATs."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "pdsat_b200", "pd-sat-b200")


@pytest.fixture(scope="module")
def cli(cuda_lib):
    from pdsat_b200 import build
    build.build_host()
    return BIN


def _instance_file(tmp_path, fixed_state_bits=0):
    """Bivium template + keystream units (+ the first N state bits as units) as a DIMACS file."""
    import bench
    from pdsat_b200 import fixtures as fx
    nv, cl, offs, lits, stream = bench.instance()
    secret = bench.mt19937_bool_rand_numpy(bench.SECRET_SEED, 177)
    units = [[(int(l >> 1) + 1) * (1 if (l & 1) == 0 else -1)] for l in stream]
    units += [[(v + 1) if secret[v] else -(v + 1)] for v in range(fixed_state_bits)]
    path = str(tmp_path / "inst.cnf")
    fx.write_dimacs(path, nv, list(cl) + units, ["input variables 177", "output variables 200"])
    return path, secret


def test_cli_estimator_matches_oracle(cli, oracle_mod):
    cost = np.asarray([230.0] * 999 + [777.0])
    s, med, pred = oracle_mod.predict_mean(cost, 31, 3)
    out = subprocess.check_output([cli, "--estimate", "1000", "0", "0", str(int(cost.sum())), "31", "3", "propagation"]).split()
    assert float(out[0]) == s
    assert abs(float(np.longdouble(out[1].decode()) - np.longdouble(med))) <= 1e-12 * float(med)
    assert abs(float((np.longdouble(out[2].decode()) - np.longdouble(pred)) / np.longdouble(pred))) <= 1e-12
    out = subprocess.check_output([cli, "--estimate", "1000", "4", "2", "0", "30", "1", "prep"]).split()
    ep = oracle_mod.predict_prep(6, 1000, 30)
    assert abs(float((np.longdouble(out[2].decode()) - np.longdouble(ep)) / np.longdouble(ep))) <= 1e-12
    out = subprocess.check_output([cli, "--estimate", "1000", "5", "0", "0", "30", "1", "prep"]).split()
    assert float(out[2]) == 1e308


def test_cli_parse_only(cli, tmp_path):
    path, _ = _instance_file(tmp_path)
    out = subprocess.check_output([cli, "--parse-only", path]).split()
    assert [int(x) for x in out] == [777, 13000, 67400, 177, 200, 0, 0]


@pytest.mark.gpu
def test_cli_predict_matches_python_mirror(cli, tmp_path):
    import bench
    from pdsat_b200 import api, predicter as pr
    path, _ = _instance_file(tmp_path)
    n = 20000
    for k, ev in ((30, "propagation"), (86, "propagation"), (87, "prep")):
        out = subprocess.check_output([cli, path, str(k), "-cnf_in_set=%d" % n, "-schema=bivium_Ending2", "-eval=" + ev],
                                      cwd=str(tmp_path)).decode()
        vals = dict(l.split(" ", 1) for l in out.strip().splitlines())
        nv, cl, offs, lits, stream = bench.instance()
        db = api.ClauseDb(nv, offs, lits, device=0)
        db.set_fixed_assumptions(stream)
        res = pr.Predicter(db, cnf_in_set_count=n, evaluation_type=ev).evaluate([pr.schema_vars("bivium_Ending2", k)])[0]
        got = np.longdouble(vals["best_predict"])
        assert abs(float((got - res.predict) / res.predict)) <= 1e-12, (k, ev, vals, res.predict)
        row = open(tmp_path / "predict").read().splitlines()[1].split()
        assert int(row[0]) == k and int(row[8]) == n
        assert int(row[6]) == res.cost.n_conflict and int(row[7]) == res.cost.n_full


@pytest.mark.gpu
def test_cli_deep_predict_step(cli, tmp_path):
    """One tabu step: the centre + its 177 Hamming-1 neighbours in one batch."""
    path, _ = _instance_file(tmp_path)
    out = subprocess.check_output([cli, path, "86", "-cnf_in_set=2000", "-schema=bivium_Ending2", "-eval=prep",
                                   "-deep_predict=1", "-max_steps=1"], cwd=str(tmp_path)).decode()
    vals = dict(l.split(" ", 1) for l in out.strip().splitlines())
    assert int(vals["points_evaluated"]) == 178
    assert float(vals["best_predict"]) < 1e308


@pytest.mark.gpu
def test_cli_solve_weakened_bivium_kat(cli, tmp_path):
    """template + keystream + the first 161 state bits as units, enumerate the
    last 16 state variables: BCP leaves exactly one assignment = the secret state's bits."""
    path, secret = _instance_file(tmp_path, fixed_state_bits=161)
    out = subprocess.check_output([cli, path, "16", "-schema=bivium_Ending2", "-solve_all"], cwd=str(tmp_path)).decode()
    vals = {}
    sat_idx = []
    for l in out.strip().splitlines():
        a, b = l.split(" ", 1)
        if a == "SAT":
            sat_idx.append(int(b.split()[1]))
        else:
            vals[a] = b
    assert int(vals["assignments"]) == 1 << 16
    assert int(vals["models"]) == 1 and int(vals["undecided"]) == 0 and int(vals["conflicts"]) == (1 << 16) - 1
    expect = sum(int(secret[161 + r]) << r for r in range(16))  # bit r <-> r-th variable ascending
    assert sat_idx == [expect]
    model = open(tmp_path / "satisfying_assignments").read().split()[0]
    assert [int(c) for c in model[:177]] == [int(b) for b in secret]
