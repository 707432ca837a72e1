"""torchrun --nproc-per-node 2 tools/peer_check.py : exercise the fused peer reduce with prints."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from pdsat_b200 import api

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
print(rank, "init ok", flush=True)
try:
    nv, cl, offs, lits, stream = bench.instance()
    db = api.ClauseDb(nv, offs, lits, device=rank)
    db.set_fixed_assumptions(stream)
    b = api.Batch(db, [api.Point(bench.SET_VARS, 100000, api.MODE_MT19937, seed=1, first_sample=rank * 100000)])
    h = b.peer_export()
    print(rank, "export ok", h[:8].hex(), flush=True)
    mine = torch.frombuffer(bytearray(h), dtype=torch.uint8).cuda()
    allh = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine)
    torch.cuda.synchronize()
    print(rank, "gather ok", flush=True)
    b.peer_connect([x.cpu().numpy().tobytes() for x in allh], rank)
    print(rank, "connect ok", flush=True)
    dist.barrier()
    for step in range(4):
        b.fill()
        b.propagate()
        print(rank, "launched", step, flush=True)
        c = b.costs()[0]
        print(rank, "step", step, c.n, c.sum_trail, flush=True)
    dist.barrier()
except Exception:
    traceback.print_exc()
    sys.stdout.flush()
    os._exit(1)
print(rank, "done", flush=True)
dist.destroy_process_group()
