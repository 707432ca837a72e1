"""The C++ search driver (pdsat_b200/host/search.hpp: tabu lists L1 / L2, simulated annealing,
ts_strategy, record handling - src_mpi/mpi_predicter.cpp:1123-1408, 1687-1828, 2405-2510,
2697-3056) against a Python mirror of the same decisions whose point evaluations come from the CPU
oracle: same centre of every step, same records, same final set.  The C++ side evaluates on the
GPU, so this also checks the whole predict path end to end (worker stream running on across
points of different size, prep estimate in long double)."""
import os
import subprocess

import numpy as np
import pytest

from pdsat_b200 import api, fixtures as fx, predicter as pr
from search_mirror import Mirror

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "pdsat_b200", "pd-sat-b200")
N_FIXED = 100


def _instance(tmp_path):
    import bench
    nv, cl, offs, lits, stream = bench.instance()
    secret = bench.mt19937_bool_rand_numpy(bench.SECRET_SEED, 177)
    units = [[(int(l >> 1) + 1) * (1 if (l & 1) == 0 else -1)] for l in stream]
    units += [[(v + 1) if secret[v] else -(v + 1)] for v in range(N_FIXED)]
    cand = list(range(N_FIXED + 1, 178))
    path = str(tmp_path / "weak.cnf")
    fx.write_dimacs(path, nv, list(cl) + units,
                    ["input variables 177", "output variables 200", "var_set " + " ".join(map(str, cand))])
    fixed = list(stream) + [2 * v + (0 if secret[v] else 1) for v in range(N_FIXED)]
    return path, cand, (nv, offs, lits, fixed)


def _oracle_evaluator(orc, inst, seed, n_samples):
    nv, offs, lits, fixed = inst
    O = orc.Oracle(nv, offs, lits, "port")
    O.add_units(fixed)
    state = {"draws": 0}

    def evaluate(sets):
        out = []
        for sv in sets:
            k = len(sv)
            n = pr.samples_for_set(k, n_samples)
            conflict, full, _, sums = O.batch_mt(0, seed, sv, n, first_draw=state["draws"], want_trail=False)
            state["draws"] += n * k
            c = api.Cost(n, sums[0], sums[1], 0, 0, 0)
            _, _, pred = pr.predict_from_cost(c, k, "prep")
            out.append((pred, 0))
        return out

    return evaluate


def _draws(orc):
    def gen(seed):
        first = 0
        while True:
            for x in orc.mt19937_draws(seed, first, 256):
                yield int(x)
            first += 256
    return gen


def _run_cli(path, cwd, extra):
    out = subprocess.check_output([BIN, path, "-eval=prep"] + extra, cwd=cwd).decode()
    centres, vals = [], {}
    for l in out.strip().splitlines():
        if l.startswith("centre "):
            centres.append([int(x) for x in l.split(":")[1].split()])
        else:
            a, b = (l.split(" ", 1) + [""])[:2]
            vals[a] = b
    return centres, vals


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["tabu", "annealing", "tabu_ts4"])
def test_search_driver_matches_python_mirror(mode, tmp_path, oracle_mod, cuda_lib):
    from pdsat_b200 import build
    build.build_host()
    path, cand, inst = _instance(tmp_path)
    seed, n, steps = 3, 300, 7
    deep = 5 if mode == "annealing" else 6
    ts = 4 if mode == "tabu_ts4" else 0
    centres, vals = _run_cli(path, str(tmp_path), ["-cnf_in_set=%d" % n, "-deep_predict=%d" % deep, "-ts_strategy=%d" % ts,
                                                   "-max_steps=%d" % steps, "-seed=%d" % seed])
    M = Mirror(cand, _oracle_evaluator(oracle_mod, inst, seed, n), deep_predict=deep, ts_strategy=ts, seed=seed,
               max_steps=steps, draws=_draws(oracle_mod)).run(cand)
    assert centres == M.centres
    assert [int(x) for x in vals["best_set"].split()] == M.best_set
    assert int(vals["records"]) == len(M.records)
    got = np.longdouble(vals["best_predict"])
    assert abs(float((got - M.best) / M.best)) <= 1e-12
    assert len(centres) >= 3 and len(M.records) >= 2
    # the files of the reference's search log
    for name in ("deep_predict", "predict", "graph_file", "L2_list", "predict_1"):
        assert os.path.exists(tmp_path / name), name
    log = open(tmp_path / "deep_predict").read()
    assert "NewRecordPoint()" in log and "*** GetDeepPredictTasks" in log and "best_var_choose_order" in log
    g = open(tmp_path / "graph_file").read().splitlines()
    assert len(g) == 1 + len(M.records)
    if mode == "annealing":
        assert "cur_temperature" in g[0] and "cur_temperature" in log


@pytest.mark.gpu
def test_search_driver_evolutionary_modes_run(tmp_path, cuda_lib):
    """FEA (algorithm_type 1) and GA (2): reproducible from -seed, records only ever improve."""
    from pdsat_b200 import build
    build.build_host()
    path, cand, inst = _instance(tmp_path)
    for alg in (1, 2):
        runs = []
        for rep in range(2):
            centres, vals = _run_cli(path, str(tmp_path), ["-cnf_in_set=200", "-deep_predict=6", "-algorithm_type=%d" % alg,
                                                           "-max_steps=40", "-seed=11", "-no_files"])
            runs.append((centres, vals["best_set"], vals["best_predict"]))
        assert runs[0] == runs[1]
        assert int(vals["points_evaluated"]) >= 20 and float(vals["best_predict"]) > 0


@pytest.mark.gpu
def test_cli_multi_gpu_flag_same_results(tmp_path, cuda_lib):
    """-gpus=N shards every point's sample range over N devices inside one host process; the sums
    are integers, so the search is identical for every N."""
    import torch
    from pdsat_b200 import build
    build.build_host()
    n_dev = torch.cuda.device_count()
    path, cand, inst = _instance(tmp_path)
    base = None
    for g in sorted({1, min(2, n_dev), n_dev}):
        centres, vals = _run_cli(path, str(tmp_path), ["-cnf_in_set=300", "-deep_predict=6", "-max_steps=3", "-seed=3",
                                                       "-gpus=%d" % g, "-no_files"])
        cur = (centres, vals["best_set"], vals["best_predict"])
        if base is None:
            base = cur
        assert cur == base
