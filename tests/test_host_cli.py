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


def _predict_rows(path):
    """rows of a `predict` file (format of src_mpi/mpi_predicter.cpp:2147-2267) as dicts"""
    rows = []
    for line in open(path).read().splitlines():
        t = line.split()
        if len(t) > 10 and t[2] == "sum" and t[4] == "pred":
            d = {"size": int(t[0]), "status": int(t[1]), "sum": float(t[3]), "pred": float(t[5])}
            for key in ("min", "max", "med", "s^2", "stopped", "skipped", "solved", "prepr:"):
                d[key] = float(t[t.index(key) + 1])
            rows.append(d)
    return rows


@pytest.mark.parametrize("deep,ts", [(6, 0), (5, 0), (6, 4), (6, 2), (3, 0)])
def test_search_driver_selftest_matches_python_mirror(deep, ts, cli):
    """The C++ search driver (host/search.hpp) on a synthetic integer objective, no GPU: every
    step's centre, the records and the final set equal the Python mirror's (tests/search_mirror.py)
    - tabu lists, L2 jumps with the mt19937 draws, annealing acceptance, swap / window strategies."""
    from oracle import oracle as orc
    from search_mirror import Mirror
    orc.build()
    n, seed, steps, init = 24, 9, 25, 12

    def evaluate(sets):
        return [(float(1000 + 3 * (len(s_) % 5) + sum((37 * v) % 11 - 4 for v in s_)), 0) for s_ in sets]

    def draws(sd):
        first = 0
        while True:
            for x in orc.mt19937_draws(sd, first, 256):
                yield int(x)
            first += 256

    out = subprocess.check_output([cli, "--search-selftest", str(n), str(deep), str(ts), str(seed), str(steps), str(init)]).decode()
    centres = [[int(x) for x in l.split(":")[1].split()] for l in out.splitlines() if l.startswith("centre ")]
    vals = dict(l.split(" ", 1) for l in out.splitlines() if not l.startswith("centre "))
    M = Mirror(list(range(1, n + 1)), evaluate, deep_predict=deep, ts_strategy=ts, seed=seed, max_steps=steps, draws=draws,
               min_temperature=0.5).run(list(range(1, init + 1)))
    assert centres == M.centres
    assert [int(x) for x in vals["best_set"].split()] == M.best_set
    assert float(vals["best_predict"]) == M.best and int(vals["records"]) == len(M.records)
    assert len(centres) >= 3


@pytest.mark.gpu
def test_cli_predict_matches_python_mirror(cli, tmp_path):
    import bench
    from pdsat_b200 import api, predicter as pr
    path, _ = _instance_file(tmp_path)
    n = 20000
    for k, ev in ((30, "bcp_trail"), (86, "bcp_trail"), (87, "prep")):
        out = subprocess.check_output([cli, path, str(k), "-cnf_in_set=%d" % n, "-schema=bivium_Ending2", "-eval=" + ev],
                                      cwd=str(tmp_path)).decode()
        vals = dict(l.split(" ", 1) for l in out.strip().splitlines())
        nv, cl, offs, lits, stream = bench.instance()
        db = api.ClauseDb(nv, offs, lits, device=0)
        db.set_fixed_assumptions(stream)
        res = pr.Predicter(db, cnf_in_set_count=n, evaluation_type=ev).evaluate([pr.schema_vars("bivium_Ending2", k)])[0]
        got = np.longdouble(vals["best_predict"])
        assert abs(float((got - res.predict) / res.predict)) <= 1e-12, (k, ev, vals, res.predict)
        text = open(tmp_path / "predict").read()
        assert text.startswith("Processor count 1\nCount of CNF in set %d\n" % n) and "Best var num %d" % k in text
        row = _predict_rows(tmp_path / "predict")[0]
        assert row["size"] == k and row["solved"] == n and row["status"] == 2
        assert row["prepr:"] == res.cost.n_conflict + res.cost.n_full
        assert abs(row["pred"] - float(res.predict)) <= 0.006 * float(res.predict)  # %.2e column
        g = open(tmp_path / "graph_file").read().splitlines()
        assert g[0].startswith("# best_var_num best_predict_time") and g[1].split()[1] == str(k)


@pytest.mark.gpu
def test_cli_eval_modes(cli, tmp_path):
    """-eval=time / watch_scans are refused; -eval=propagation hands the undecided samples to the
    host CDCL under a conflict budget and marks the set STOPPED when samples are interrupted."""
    path, _ = _instance_file(tmp_path)
    for ev in ("time", "watch_scans"):
        r = subprocess.run([cli, path, "30", "-cnf_in_set=100", "-schema=bivium_Ending2", "-eval=" + ev], cwd=str(tmp_path),
                           capture_output=True)
        assert r.returncode != 0 and b"not supported" in r.stderr
    path2, _ = _instance_file(tmp_path, fixed_state_bits=110)
    out = subprocess.check_output([cli, path2, "20", "-cnf_in_set=256", "-schema=bivium_Ending2", "-eval=propagation",
                                   "-max_conflicts=100000", "-threads=2"], cwd=str(tmp_path)).decode()
    vals = dict(l.split(" ", 1) for l in out.strip().splitlines())
    assert float(vals["best_predict"]) > 0
    row = _predict_rows(tmp_path / "predict")[0]
    assert row["status"] == 2 and row["stopped"] == 0 and row["solved"] == 256


@pytest.mark.gpu
def test_cli_deep_predict_step(cli, tmp_path):
    """One tabu step: the centre + its 177 Hamming-1 neighbours in one batch."""
    path, _ = _instance_file(tmp_path)
    out = subprocess.check_output([cli, path, "86", "-cnf_in_set=2000", "-schema=bivium_Ending2", "-eval=prep",
                                   "-deep_predict=6", "-max_steps=1"], cwd=str(tmp_path)).decode()
    vals = dict(l.split(" ", 1) for l in out.strip().splitlines() if not l.startswith("centre"))
    assert int(vals["points_evaluated"]) == 178
    assert float(vals["best_predict"]) < 1e308
    assert len(_predict_rows(tmp_path / "predict")) == 177
    assert "L2 point # : set_count : size" in open(tmp_path / "L2_list").read()


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


@pytest.mark.gpu
def test_cli_interval_estimation(cli, tmp_path, oracle_mod):
    """-interval_predict (calculateIntervalEstimation, src_mpi/mpi_predicter.cpp:3058-3181, rc2):
    `size` consecutive assignments in the binary order of Solver::gen_valid_assumptions_rc2
    (d_set[0] most significant, Solver.cc:1380-1423) from a random corner; the number solved on
    preprocessing / surviving BCP equals the reference BCP's count over the same interval."""
    import bench
    path, secret = _instance_file(tmp_path, fixed_state_bits=140)
    k, size, seed = 30, 1 << 15, 5
    out = subprocess.check_output([cli, path, str(k), "-schema=bivium_Ending2", "-interval_predict", "-interval_size=%d" % size,
                                   "-interval_required=50", "-seed=%d" % seed, "-max_conflicts=50000"], cwd=str(tmp_path)).decode()
    vals = dict(l.split(" ", 1) for l in out.strip().splitlines())
    sv = sorted(177 - i for i in range(k))
    bits = oracle_mod.bool_rand(seed, 0, k)
    start = 0
    for b in bits:
        start = (start << 1) | int(b)
    start = min(start, (1 << k) - size)
    assert int(vals["interval_start"]) == start and int(vals["interval_size"]) == size
    nv, cl, offs, lits, stream = bench.instance()
    O = oracle_mod.Oracle(nv, offs, lits, "port")
    O.add_units(list(stream) + [2 * v + (0 if secret[v] else 1) for v in range(140)])
    conflict, full, sums = O.batch_enum(sv[::-1], start, size)   # bit 0 of the index = last variable
    n_surv = int((~conflict & ~full).sum())
    assert int(vals["interval_nonprepr_number"]) == n_surv
    assert int(vals["interval_prepr_number"]) == size - n_surv
    solved = vals["survivors_solved"].split()
    assert int(solved[0]) == min(50, n_surv)
    assert int(solved[0]) == int(solved[2]) + int(solved[4]) + int(solved[6])
    assert float(vals["interval_predict"]) > 0
