"""Edges of the C ABI that round 1 left untested: 64-bit sum of squares / 32-bit trails on
instances with more than 8 192 (65 535) level-0 units, stale batches after
pdsat_set_fixed_assumptions, back-to-back reseed + fill, stream continuation (first_draw)."""
import numpy as np
import pytest

from pdsat_b200 import api, fixtures as fx


def _padded_bivium(bivium, n_pad):
    """Bivium template + n_pad extra variables, each fixed by a unit clause (level-0 trail)."""
    nv, cl, _, _ = bivium
    cls = [list(c) for c in cl] + [[nv + 1 + i] for i in range(n_pad)]
    offs, lits = fx.to_csr(cls)
    return nv + n_pad, offs, lits


@pytest.mark.gpu
@pytest.mark.parametrize("n_pad", [9000, 70000])
def test_gpu_sumsq_and_trail_beyond_16_bits(n_pad, bivium, oracle_mod, cuda_lib):
    """trail >= 8192 in every sample of a group: sum of 64 squares needs more than 32 bits
    (src_mpi/mpi_predicter.cpp:2219-2224 feeds on it); trail >= 65536 needs a 32-bit trail."""
    import bench
    orc = oracle_mod
    nv, offs, lits = _padded_bivium(bivium, n_pad)
    stream = bench.instance()[4]
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    assert db.info().n_base_assigned == n_pad + 200
    P = orc.Oracle(nv, offs, lits, "port")
    P.add_units(list(stream))
    n = 640
    pts = [api.Point([177 - i for i in range(k)], n, api.MODE_MT19937, seed=5) for k in (30, 86)]
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL)
    costs = b.run()
    for p, pt in enumerate(pts):
        A = orc.sample_assumptions(pt.seed, pt.set_vars, 0, n)
        tr = b.trail(p)
        conflict, _ = b.flags(p)
        s1 = s2 = 0
        for s in range(n):
            r, _ = P.bcp(list(A[s]), want_assigns=False)
            assert bool(conflict[s]) == bool(r.conflict)
            if not r.conflict:
                assert int(tr[s]) == r.trail_size > n_pad
                s1 += r.trail_size
                s2 += r.trail_size * r.trail_size
        assert costs[p].sum_trail == s1
        assert costs[p].sumsq_trail == s2  # exact integers
        if p == 1:
            assert costs[p].n_implied_any > 0  # the per-lane (implied) accumulation path ran
        assert s2 > 64 * 8192 * 8192 // 2
    b.close()
    db.close()


@pytest.mark.gpu
def test_gpu_stale_batch_is_refused(bivium, cuda_lib):
    import bench
    nv, cl, offs, lits = bivium
    db = api.ClauseDb(nv, offs, lits, device=0)
    b = api.Batch(db, [api.Point(list(range(148, 178)), 128, api.MODE_MT19937)])
    b.run()
    db.set_fixed_assumptions(bench.instance()[4])  # renumbers the live variables
    with pytest.raises(api.PdsatError) as e:
        b.fill()
    assert e.value.code == api.PDSAT_ERR_ARG and "stale" in str(e.value)
    b2 = api.Batch(db, [api.Point(list(range(148, 178)), 128, api.MODE_MT19937)])
    assert b2.run()[0].sum_trail == 128 * 230
    # destroying the database first is safe: the batches keep it alive
    db.close()
    b2.close()
    b.close()


@pytest.mark.gpu
def test_gpu_back_to_back_reseed_fill(bivium, oracle_mod, cuda_lib):
    """reseed + fill twice without a sync in between: the second stream must be the one used
    (the pinned staging buffers of the first fill are still being copied)."""
    import bench
    orc = oracle_mod
    nv, cl, offs, lits = bivium
    stream = bench.instance()[4]
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    P = orc.Oracle(nv, offs, lits, "port")
    P.add_units(list(stream))
    sv = [177 - i for i in range(88)]
    n = 4096
    b = api.Batch(db, [api.Point(sv, n, api.MODE_MT19937, seed=1)], api.OUT_FLAGS)
    for rep in range(5):
        b.reseed(0, 11 + rep, 0)
        b.fill()
        b.reseed(0, 900 + rep, 64 * rep)
        b.fill()
        b.propagate()
        conflict, _ = b.flags(0)
        A = orc.sample_assumptions(900 + rep, sv, 64 * rep, n)
        want = np.array([P.bcp(list(A[s]), want_assigns=False)[0].conflict for s in range(0, n, 37)], dtype=bool)
        assert (conflict[::37] == want).all()
    b.close()
    db.close()


@pytest.mark.gpu
def test_gpu_stream_continuation_first_draw(bivium, oracle_mod, cuda_lib):
    """one generator per worker, seeded once, running on across points of different k
    (src_mpi/mpi_predicter.cpp:87, 3236-3242): point i starts at the draw where point i-1 ended"""
    import bench
    orc = oracle_mod
    nv, cl, offs, lits = bivium
    stream = bench.instance()[4]
    db = api.ClauseDb(nv, offs, lits, device=0)
    db.set_fixed_assumptions(stream)
    P = orc.Oracle(nv, offs, lits, "port")
    P.add_units(list(stream))
    sets = [[177 - i for i in range(k)] for k in (85, 31, 90, 87)]
    ns = [300, 1000, 77, 500]
    pts, draw = [], 0
    for sv, n in zip(sets, ns):
        pts.append(api.Point(sv, n, api.MODE_MT19937, seed=4, first_draw=draw))
        draw += n * len(sv)
    b = api.Batch(db, pts, api.OUT_FLAGS | api.OUT_TRAIL)
    b.run()
    bits = orc.bool_rand(4, 0, draw)
    d = 0
    for p, (sv, n) in enumerate(zip(sets, ns)):
        conflict, _ = b.flags(p)
        tr = b.trail(p)
        for s in range(n):
            assum = [2 * (v - 1) + (0 if bits[d + j] else 1) for j, v in enumerate(sv)]
            d += len(sv)
            r, _ = P.bcp(assum, want_assigns=False)
            assert bool(conflict[s]) == bool(r.conflict), (p, s)
            if not r.conflict:
                assert int(tr[s]) == r.trail_size
    # moving a point along the stream afterwards
    b.set_first_draw(0, 12345)
    b.run()
    bits = orc.bool_rand(4, 12345, ns[0] * len(sets[0]))
    conflict, _ = b.flags(0)
    for s in range(ns[0]):
        assum = [2 * (v - 1) + (0 if bits[s * 85 + j] else 1) for j, v in enumerate(sets[0])]
        assert bool(conflict[s]) == bool(P.bcp(assum, want_assigns=False)[0].conflict)
    b.close()
    db.close()
