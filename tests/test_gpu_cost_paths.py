"""The two ways bcp_kernel hands its per-point accumulators to the host must agree: the direct
path (library-owned double buffer, no ticket) and the ticket path (last CTA publishes into a
caller-owned buffer / the stable device pointer), across repeated launches and path changes."""
import pytest

from pdsat_b200 import api


def _key(costs):
    return [(c.n, c.n_conflict, c.n_full, c.sum_trail, c.sumsq_trail) for c in costs]


@pytest.mark.gpu
def test_gpu_direct_and_ticket_cost_paths_agree(bivium, cuda_lib):
    import torch
    import bench
    nv, cl, offs, lits = bivium
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(bench.instance()[4])
    # a streaming point, two propagating points, a ragged one: several points per CTA
    pts = [api.Point([177 - i for i in range(k)], n, api.MODE_MT19937, seed=3 + k)
           for k, n in ((30, 200000), (86, 30000), (100, 20001), (88, 777))]
    b = api.Batch(db, pts)
    b.fill()
    b.propagate()
    want = _key(b.costs())
    assert want[0][0] == 200000 and want[1][1] > 0
    for _ in range(5):  # direct path: the other buffer is cleared by the launch before
        b.propagate()
        assert _key(b.costs()) == want
    ext = torch.full((api.COST_WORDS * len(pts),), 12345, dtype=torch.int64, device="cuda")
    b.set_cost_buffer(ext.data_ptr())  # ticket path into a caller-owned (dirty) buffer
    for _ in range(3):
        b.propagate()
        assert _key(b.costs()) == want
    got = ext.cpu().numpy().reshape(len(pts), api.COST_WORDS)
    assert [int(r[0]) for r in got] == [w[0] for w in want]
    b.set_cost_buffer(0)  # back to the library's buffers
    for _ in range(3):
        b.propagate()
        assert _key(b.costs()) == want
    ptr = b.cost_device_ptr()  # pins the ticket path: the pointer stays valid from now on
    for _ in range(3):
        b.propagate()
        assert _key(b.costs()) == want
        assert b.cost_device_ptr() == ptr
    b.close()
    db.close()


@pytest.mark.gpu
def test_gpu_timing_switch(bivium, cuda_lib):
    import bench
    nv, cl, offs, lits = bivium
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(bench.instance()[4])
    b = api.Batch(db, [api.Point([177 - i for i in range(86)], 50000, api.MODE_MT19937, seed=9)])
    want = _key(b.run())
    f, p = b.last_ms()
    assert f > 0 and p > 0
    b.set_timing(False)
    assert _key(b.run()) == want
    assert b.last_ms() == (0.0, 0.0)
    b.set_timing(True)
    assert _key(b.run()) == want
    f, p = b.last_ms()
    assert f > 0 and p > 0
    b.close()
    db.close()
