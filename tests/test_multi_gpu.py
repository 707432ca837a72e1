"""GPU, N = 2 ranks over NCCL (skipped on a 1-GPU box): sample sharding + one all-reduce must
give exactly the single-GPU accumulators (integer sums are rank-count independent)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, sets, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import bench
    from pdsat_b200 import api, predicter as pr
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    nv, cl, offs, lits, stream = bench.instance()
    db = api.ClauseDb(nv, offs, lits, device=rank)
    db.set_fixed_assumptions(stream)

    def all_reduce(arr):
        t = torch.from_numpy(arr).cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()

    P = pr.Predicter(db, cnf_in_set_count=n, rank=rank, world=world, all_reduce=all_reduce)
    res = P.evaluate(sets)
    if rank == 0:
        q.put([[r.cost.n, r.cost.n_conflict, r.cost.n_full, r.cost.sum_trail, r.cost.sumsq_trail] for r in res])
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_matches_one_gpu(cuda_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import bench
    from pdsat_b200 import api, predicter as pr
    sets = [list(range(148, 178)), [177 - i for i in range(86)], [177 - i for i in range(87)], [3, 5, 9]]
    n = 100_003
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29650, sets, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    nv, cl, offs, lits, stream = bench.instance()
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    one = pr.Predicter(db, cnf_in_set_count=n).evaluate(sets)
    for g, r in zip(got, one):
        assert g == [r.cost.n, r.cost.n_conflict, r.cost.n_full, r.cost.sum_trail, r.cost.sumsq_trail]


BURSTS = [(5, False), (6, False), (7, True), (8, False), (9, True), (10, True), (11, False), (12, False),
          (13, False), (14, False), (15, True)]


def _peer_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import bench
    from pdsat_b200 import api, predicter as pr
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    nv, cl, offs, lits, stream = bench.instance()
    db = api.ClauseDb(nv, offs, lits, device=rank)
    db.set_fixed_assumptions(stream)
    sets = [list(range(148, 178)), [177 - i for i in range(86)]]
    pts = []
    for p, sv in enumerate(sets):
        lo, hi = pr.shard_range(n, rank, world)
        pts.append(api.Point(sv, hi - lo, api.MODE_MT19937, seed=1 + p, first_sample=lo))
    b = api.Batch(db, pts)
    mine = torch.frombuffer(bytearray(b.peer_export()), dtype=torch.uint8).cuda()
    allh = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine)
    b.peer_connect([h.cpu().numpy().tobytes() for h in allh], rank)
    dist.barrier()
    out = []
    for step in range(5):  # several steps: exercises the parity double-buffering
        for p in range(len(pts)):
            b.reseed(p, 1 + p + 10 * step, pts[p].first_sample)
        c = b.run()
        out.append([[x.n, x.n_conflict, x.n_full, x.sum_trail, x.sumsq_trail] for x in c])
    # the deferred exchange: launches whose costs are never asked for (their sums are pushed and
    # gathered by the launches that follow), then a read; bursts of different lengths
    for step, read in BURSTS:
        for p in range(len(pts)):
            b.reseed(p, 1 + p + 10 * step, pts[p].first_sample)
        b.fill()
        b.propagate()
        if read:
            out.append([[x.n, x.n_conflict, x.n_full, x.sum_trail, x.sumsq_trail] for x in b.costs()])
    q.put((rank, out))
    dist.barrier()
    b.close()
    dist.destroy_process_group()


def test_fused_peer_reduce_matches_one_gpu(cuda_lib):
    """The NVLink peer-memory reduce fused into the BCP kernel: every rank ends each step with the
    sums of ALL ranks, equal to a single-GPU evaluation of the whole sample range."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import bench
    from pdsat_b200 import api
    n = 50_001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, 29670, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0] == got[1]
    nv, cl, offs, lits, stream = bench.instance()
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    sets = [list(range(148, 178)), [177 - i for i in range(86)]]
    read_steps = list(range(5)) + [st for st, read in BURSTS if read]
    assert len(got[0]) == len(read_steps)
    for i, step in enumerate(read_steps):
        costs = db.eval_batch([api.Point(sv, n, api.MODE_MT19937, seed=1 + p + 10 * step) for p, sv in enumerate(sets)])
        exp = [[x.n, x.n_conflict, x.n_full, x.sum_trail, x.sumsq_trail] for x in costs]
        assert got[0][i] == exp, step
