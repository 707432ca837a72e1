"""CPU, world_size 2, gloo: the N>1 host path — contiguous sample sharding per point, one
all-reduce of the integer accumulators, estimator on the reduced sums.  The per-rank
"evaluation" is done by the CPU oracle here (no GPU in this container); the arithmetic and
the plumbing are the ones bench.py / Predicter use on the GPUs."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, sets, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    from pdsat_b200 import fixtures as fx, predicter as pr
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nv, cl, _ = fx.bivium_template()
    offs, lits = fx.to_csr(cl)
    P = orc.Oracle(nv, offs, lits, "port")
    rows = []
    for p, sv in enumerate(sets):
        lo, hi = pr.shard_range(n, rank, world)
        A = orc.sample_assumptions(1 + p, sv, lo, hi - lo)
        acc = [0] * 8
        for s in range(hi - lo):
            r, _ = P.bcp(list(A[s]), want_assigns=False)
            acc[0] += 1
            acc[1] += r.conflict
            acc[2] += r.full
            if not r.conflict:
                acc[3] += r.trail_size
                acc[4] += r.trail_size ** 2
        rows.append(acc)
    t = torch.tensor(rows, dtype=torch.int64)
    dist.all_reduce(t)  # the single exchange step of the path
    if rank == 0:
        q.put(t.numpy().tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce(oracle_mod):
    import torch.multiprocessing as mp
    from pdsat_b200 import api, fixtures as fx, predicter as pr
    sets = [list(range(1, 151)), [177 - i for i in range(30)], list(range(1, 178))]
    n = 37  # ragged on purpose
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, sets, q)) for r in range(2)]
    for p in procs:
        p.start()
    reduced = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference
    nv, cl, _ = fx.bivium_template()
    offs, lits = fx.to_csr(cl)
    P = oracle_mod.Oracle(nv, offs, lits, "port")
    for p, sv in enumerate(sets):
        A = oracle_mod.sample_assumptions(1 + p, sv, 0, n)
        costs = []
        acc = [0] * 5
        for s in range(n):
            r, _ = P.bcp(list(A[s]), want_assigns=False)
            acc[0] += 1
            acc[1] += r.conflict
            acc[2] += r.full
            if not r.conflict:
                acc[3] += r.trail_size
                acc[4] += r.trail_size ** 2
                costs.append(float(r.trail_size))
        assert reduced[p][:5] == acc
        c = api.Cost(*reduced[p][:6])
        s_, med, pred = pr.predict_from_cost(c, len(sv), "bcp_trail", proc_count=4)
        if costs and acc[1] == 0:
            es, emed, epred = oracle_mod.predict_mean(np.asarray(costs), len(sv), 4)
            assert s_ == es
            assert abs(float(med - np.longdouble(emed))) <= 1e-12 * abs(float(emed))
            assert abs(float((pred - np.longdouble(epred)) / np.longdouble(epred))) <= 1e-12
        solved = acc[1] + acc[2]
        _, _, pp = pr.predict_from_cost(c, len(sv), "prep")
        ep = oracle_mod.predict_prep(solved, n, len(sv))
        assert abs(float((pp - np.longdouble(ep)) / np.longdouble(ep))) <= 1e-12


def test_shard_ranges_partition():
    from pdsat_b200 import predicter as pr
    for n in (0, 1, 7, 64, 1000, 10**7 + 3):
        for world in (1, 2, 4, 8):
            ranges = [pr.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]


def test_samples_for_set_and_schemas():
    from pdsat_b200 import predicter as pr
    assert pr.samples_for_set(5, 1000) == 32 and pr.samples_for_set(30, 10**7) == 10**7
    assert pr.samples_for_set(31, 10**7) == 10**7 and pr.samples_for_set(10, 100) == 100
    assert pr.schema_vars("bivium_Ending2", 30) == list(range(148, 178))
    assert pr.schema_vars("bivium_Beginning1", 30) == list(range(1, 31))
    nb = pr.hamming1_neighbourhood(list(range(148, 178)), list(range(1, 178)))
    assert len(nb) == 177 and sorted(len(s) for s in nb) == [29] * 30 + [31] * 147


def test_sample_variance_from_integer_sums(oracle_mod):
    from pdsat_b200 import api, predicter as pr
    rng = np.random.default_rng(2)
    x = rng.integers(230, 778, size=5000)
    c = api.Cost(len(x), 0, 0, int(x.sum()), int((x.astype(np.int64) ** 2).sum()), 0)
    ref = oracle_mod.sample_variance(x.astype(np.float64))
    assert abs(pr.sample_variance(c) - ref) <= 1e-12 * ref
