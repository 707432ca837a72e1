#!/usr/bin/env python
"""bench.py — PDSAT decomposition-set evaluation on B200 (BASELINE.json metric).

Workload (configs[1]): Bivium template + 200 keystream unit assumptions, fixed 30-variable
start-state decomposition set {148..177} (schema bivium_Ending2), N_SAMPLES random samples
per GPU per step, sample bits = bool_rand of mt19937 (the reference's generator).
A "step" = one pass of the hot path (BCP of every sample + integer cost reduce [+ NCCL
all-reduce of the per-point accumulators when N > 1]) over one batch.

  value   samples/s, device-timed, assumption words already resident in HBM
  e2e     samples/s through the public call sequence reseed -> fill (device mt19937 +
          transpose) -> propagate -> costs() (D2H), host wall clock
  --impl reference : the reference's own CPU path (oracle/_ref = unmodified MiniSat replaying
          solverProgramCalling; falls back to the C port) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SET_VARS = list(range(148, 178))  # bivium_Ending2, first 30 (src_mpi/mpi_base.cpp:319-322)
SECRET_SEED = 0


def instance():
    """Bivium template + keystream literals of the secret state."""
    from pdsat_b200 import fixtures as fx
    nv, cl, _ = fx.bivium_template()
    offs, lits = fx.to_csr(cl)
    # secret state: first 177 bool_rand bits of mt19937(0); computed with numpy's MT19937
    # (same generator) so that the product path does not need the oracle
    secret = mt19937_bool_rand_numpy(SECRET_SEED, 177)
    ks = fx.BiviumModel(secret).stream(200)
    stream = np.asarray([2 * (577 + i) + (0 if ks[i] else 1) for i in range(200)], dtype=np.int32)
    return nv, cl, offs, lits, stream


def mt19937_bool_rand_numpy(seed: int, n: int) -> np.ndarray:
    """bool_rand bits via numpy's MT19937 seeded the mt19937 way (init_genrand)."""
    mt = np.zeros(624, dtype=np.uint32)
    x = seed & 0xFFFFFFFF
    mt[0] = x
    for i in range(1, 624):
        x = (1812433253 * (x ^ (x >> 30)) + i) & 0xFFFFFFFF
        mt[i] = x
    bg = np.random.MT19937()
    st = bg.state
    st["state"]["key"] = mt
    st["state"]["pos"] = 624
    bg.state = st
    draws = bg.random_raw(n).astype(np.uint64)
    return ((draws >> 31) == 0).astype(np.uint8)


def algorithmic_bytes_per_sample(nv, clauses, true_lits_dimacs):
    """sum over true literals p of 4*|occ(~p)| + 4*sum |c| for c in occ(~p),
    + 4*|T| + 8."""
    occ = {}
    for c in clauses:
        for l in c:
            occ.setdefault(l, []).append(len(c))
    total = 0
    for p in true_lits_dimacs:
        lens = occ.get(-p, [])
        total += 4 * len(lens) + 4 * sum(lens)
    return total + 4 * len(true_lits_dimacs) + 8


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm
# Everything below that touches oracle/ is the CHECKER or the reported CPU baseline, never the
# product path.  One process per host core, static split of the sample range; the per-sample
# loop runs in C (orc_batch_mt), so the figures are the solver's, not Python's.
_CPU = {}


def _cpu_oracle(kind, with_units):
    key = (kind, with_units)
    if key not in _CPU:
        from oracle import oracle as orc
        nv, cl, offs, lits, stream = instance()
        O = orc.Oracle(nv, offs, lits, kind)
        if with_units:
            O.add_units(list(stream))
        _CPU[key] = (O, [int(x) for x in stream])
    return _CPU[key]


def _cpu_worker(args):
    """(kind, mode, seed, set_vars, first_sample, n, want) -> (seconds, sums, conflict bits, trail)"""
    kind, mode, seed, set_vars, first, n, want = args
    O, stream = _cpu_oracle(kind, mode == 0)
    t0 = time.perf_counter()
    conflict, full, trail, sums = O.batch_mt(mode, seed, set_vars, n, first_sample=first,
                                             extra=stream if mode == 1 else (), want_trail=want)
    dt = time.perf_counter() - t0
    if want:
        return dt, sums, np.packbits(conflict, bitorder="little"), trail
    return dt, sums, None, None


_POOL = None


def cpu_pool():
    global _POOL
    if _POOL is None:
        import multiprocessing as mp
        from oracle import oracle as orc
        orc.build()
        _POOL = mp.get_context("spawn").Pool(os.cpu_count() or 1)
    return _POOL


def cpu_kind():
    from oracle import oracle as orc
    orc.build()
    return "ref" if orc.have_ref() else "port"


def cpu_run(mode, seed, set_vars, n, first_sample=0, want=False, cores=None):
    """Reference CPU path over samples [first_sample, first_sample + n) on all host cores.
    mode 1 = solverProgramCalling (fresh Solver + addProblem per sample, eval=prep),
    mode 0 = BCP only on a persistent clause DB."""
    cores = cores or os.cpu_count() or 1
    kind = cpu_kind()
    cuts = [first_sample + n * i // cores for i in range(cores + 1)]
    jobs = [(kind, mode, seed, list(set_vars), cuts[i], cuts[i + 1] - cuts[i], want) for i in range(cores)
            if cuts[i + 1] > cuts[i]]
    t0 = time.perf_counter()
    res = cpu_pool().map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    sums = [sum(r[1][j] for r in res) for j in range(6)]
    out = {"n": n, "cores": len(jobs), "busy_s": busy, "wall_s": wall, "samples_per_s": n / busy, "sums": sums,
           "kind": "reference" if kind == "ref" else "port"}
    if want:
        out["conflict"] = np.concatenate([np.unpackbits(r[2], bitorder="little")[:j[5]].astype(bool)
                                          for r, j in zip(res, jobs)])
        out["trail"] = np.concatenate([r[3] for r in res])
    return out


def cpu_arm(n_per_core, mode="faithful", cores=None):
    cores = cores or os.cpu_count() or 1
    r = cpu_run(1 if mode == "faithful" else 0, 1000, SET_VARS, n_per_core * cores, cores=cores)
    return {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"],
            "sample": "%d samples/core x %d cores, %s" % (n_per_core, r["cores"],
                      "fresh Solver + addProblem per sample (solverProgramCalling, eval=prep)" if mode == "faithful"
                      else "persistent clause DB, BCP only")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_per_core = args.cpu_samples
    cores = os.cpu_count() or 1
    # one calibration pass (also the warm-up); the per-step sample is bounded so that the K steps
    # together stay near 90 s of wall clock whatever K the caller asks for
    cal = cpu_arm(max(20, n_per_core // 10))
    per_core_rate = cal["value"] / cores
    n_per_core = int(max(20, min(n_per_core, 90.0 * per_core_rate / max(1, args.steps))))
    t_all = 0.0
    last = None
    for _ in range(args.steps):
        last = cpu_arm(n_per_core)
        t_all += cores * n_per_core / last["value"]
    v = cores * n_per_core * args.steps / t_all
    line = {"impl": "reference", "metric": "decomp-set samples/sec", "value": v, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(cores * n_per_core),
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": last["kind"],
                             "sample": last["sample"]},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_samples):
    return {"workload": "Bivium bivium_template (777 vars, 12800 clauses) + 200 keystream units, "
                        "30-var start-state set {148..177}, mt19937 bool_rand samples, predict (eval=prep/propagation)",
            "samples_per_gpu_per_step": n_samples, "set_size": len(SET_VARS),
            "l2": "device-timed steps rotate over resident batches whose words total > 2x the 126 MB L2 "
                  "(every launch reads from HBM); cascade / isolated launches: 256 MiB flush before each"}


L2_BYTES = 126 * 1024 * 1024


# ----------------------------------------------------------------------------- parity + cascade legs
def gpu_vs_cpu_per_sample(api, db, set_vars, seed, n, first_sample=0):
    """EVERY sample of the batch: conflict flag and trail (= Solver::propagations of a
    conflict-free sample) of the device against the reference BCP on the host cores."""
    b = api.Batch(db, [api.Point(list(set_vars), n, api.MODE_MT19937, seed=seed, first_sample=first_sample)],
                  api.OUT_FLAGS | api.OUT_TRAIL)
    cost = b.run()[0]
    conflict, _ = b.flags(0)
    trail = b.trail(0)
    b.close()
    cpu = cpu_run(0, seed, set_vars, n, first_sample=first_sample, want=True)
    ok = ~cpu["conflict"]
    bad_c = int((conflict != cpu["conflict"]).sum())
    bad_t = int((trail[ok] != cpu["trail"][ok]).sum())
    sums_ok = (cost.n_conflict == cpu["sums"][0] and cost.sum_trail == cpu["sums"][2]
               and cost.sumsq_trail == cpu["sums"][3])
    return {"samples": n, "conflict_mismatches": bad_c, "trail_mismatches": bad_t, "integer_sums_equal": bool(sums_ok),
            "checker": "oracle/_ref (unmodified reference MiniSat)" if cpu["kind"] == "reference" else "oracle port",
            "cpu_bcp_only_samples_per_s": cpu["samples_per_s"], "cpu_cores": cpu["cores"],
            "cpu_propagations": cpu["sums"][4], "cpu_watch_scans": cpu["sums"][5]}, cost


def work_equivalent_bytes(nv, clauses, stream, set_vars, seed, n_probe=400):
    """SURVEY.md 8(d) byte formula on a probe of samples (what a per-sample watched-literal BCP
    touches): mean over conflict-free samples of sum_{p in T} [4|occ(~p)| + 4 sum|c|] + 4|T| + 8."""
    from oracle import oracle as orc
    O, st = _cpu_oracle(cpu_kind(), True)
    occ_cost = {}
    for c in clauses:
        for l in c:
            a = occ_cost.setdefault(l, [0, 0])
            a[0] += 1
            a[1] += len(c)
    A = orc.sample_assumptions(seed, set_vars, 0, n_probe)
    tot, cnt = 0, 0
    for s in range(n_probe):
        r, asg = O.bcp(list(A[s]))
        if r.conflict:
            continue
        b = 8
        for v in range(nv):
            if asg[v] == 2:
                continue
            p = (v + 1) if asg[v] == 0 else -(v + 1)
            oc = occ_cost.get(-p, (0, 0))
            b += 4 * oc[0] + 4 * oc[1] + 4
        tot += b
        cnt += 1
    return tot / cnt if cnt else None


def _cpu_enum_worker(args):
    kind, n_known, first, n = args
    from oracle import oracle as orc
    from pdsat_b200 import fixtures as fx
    key = ("a52", kind, n_known)
    if key not in _CPU:
        nv, cl, fixed, dset, secret = fx.a5_2_weakened_instance(n_known)
        offs, lits = fx.to_csr(cl)
        O = orc.Oracle(nv, offs, lits, kind)
        O.add_units(fixed)
        _CPU[key] = (O, dset)
    O, dset = _CPU[key]
    t0 = time.perf_counter()
    conflict, full, sums = O.batch_enum(dset, first, n)
    return time.perf_counter() - t0, sums, np.packbits(conflict, bitorder="little"), np.packbits(full, bitorder="little")


def leg_tabu(api, torch, dist, db, rank, world, args):
    """BASELINE config 3: one tabu step = the centre {148..177} and its 177 Hamming-1 neighbours
    (src_mpi/mpi_predicter.cpp:2770, 2926-2935), tabu_samples mt19937 samples per point (stream
    seed = 1 + point), every point's sample range sharded over the ranks, ONE reduce of the 178 x 8
    accumulators fused into the BCP kernel over NVLink.  Timed end to end through the public call
    sequence; the all-rank sums must equal an un-sharded evaluation on rank 0, integer for integer."""
    from pdsat_b200 import predicter as pr
    sets = [SET_VARS] + pr.hamming1_neighbourhood(SET_VARS, list(range(1, 178)))
    n_pt = args.tabu_samples
    lo, hi = pr.shard_range(n_pt, rank, world)
    batch = api.Batch(db, [api.Point(sv, hi - lo, api.MODE_MT19937, seed=1 + i, first_sample=lo)
                           for i, sv in enumerate(sets)])
    mode = "none"
    cost_t = None
    if dist is not None:
        cost_t = torch.zeros(api.COST_WORDS * len(sets), dtype=torch.int64, device="cuda")
        batch.set_cost_buffer(cost_t.data_ptr())
        mode = connect_peers(torch, dist, batch, rank, world, args.reduce)

    def step():
        batch.fill()
        batch.propagate()
        if mode == "nccl":
            dist.all_reduce(cost_t)
        return batch.costs()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        step()
    barrier()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        costs = step()
    barrier()
    dt = (time.perf_counter() - t0) / reps
    f_ms, p_ms = batch.last_ms()
    t = torch.tensor([dt, f_ms, p_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
    dt, f_ms, p_ms = (float(x) for x in t.tolist())
    batch.close()
    out = None
    if rank == 0:
        got = [[c.n, c.n_conflict, c.n_full, c.sum_trail, c.sumsq_trail, c.n_implied_any] for c in costs]
        identical = None
        if world > 1:  # the same step un-sharded on this GPU alone
            whole = db.eval_batch([api.Point(sv, n_pt, api.MODE_MT19937, seed=1 + i) for i, sv in enumerate(sets)])
            identical = got == [[c.n, c.n_conflict, c.n_full, c.sum_trail, c.sumsq_trail, c.n_implied_any] for c in whole]
        est = [pr.predict_from_cost(c, len(sv), "prep") for c, sv in zip(costs, sets)]
        out = {"points": len(sets), "samples_per_point": n_pt, "ranks": world, "reduce": mode,
               "e2e_ms_per_step": dt * 1e3, "fill_ms": f_ms, "propagate_ms": p_ms,
               "samples_per_s_e2e": len(sets) * n_pt / dt, "points_per_s": len(sets) / dt,
               "all_points_complete": all(g[0] == n_pt for g in got),
               "bit_identical_to_one_gpu": identical,
               "sum_check": [sum(g[j] for g in got) for j in range(6)],
               "best_prep_estimate_finite": bool(any(float(e[2]) < 1e300 for e in est))}
    return out


def leg_solve(api, torch, dist, rank, world, args):
    """BASELINE config 4: generated A5/2 CNF, keystream + R4 + 32 state bits known, solve-mode
    enumeration of all 2^32 assignments of the other 32 state variables.  Rank g owns the
    assignments whose top log2(N) index bits equal g (task split of mpi_solver.cpp:447-473);
    BCP-pruned enumeration on the device (pdsat_enumerate); then allreduce(max) of the found flag
    and allgather of the hit indices.  Rank 0 also checks a 2^20 slice around the secret against
    the reference BCP on the host cores (conflict / model flags of every assignment)."""
    from pdsat_b200 import fixtures as fx
    nv, cl, fixed, dset, secret = fx.a5_2_weakened_instance(args.a52_known)
    offs, lits = fx.to_csr(cl)
    t0 = time.perf_counter()
    db = api.ClauseDb(nv, offs, lits, device=torch.cuda.current_device())
    db.set_fixed_assumptions(fixed)
    upload_s = time.perf_counter() - t0
    k = len(dset)
    top = world.bit_length() - 1
    assert (1 << top) == world, "solve leg needs a power-of-two number of ranks"
    idx = sum(int(secret[v - 1]) << r for r, v in enumerate(dset))
    db.enumerate(dset[-8:], max_survivors=16)  # warm-up (module load, allocations)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    times = []
    for rep in range(3):  # the first repetition also pays the allocations of the level batches
        barrier()
        t0 = time.perf_counter()
        surv, ism, st = db.enumerate(dset, n_fixed_top=top, fixed_top_value=rank, max_survivors=1 << 12)
        barrier()
        times.append(time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    surv, ism, st = db.enumerate(dset, n_fixed_top=top, fixed_top_value=rank, max_survivors=1 << 12)
    hits = torch.full((16,), -1, dtype=torch.int64, device="cuda")
    models = [int(x) for x, m in zip(surv, ism) if m][:16]
    if models:
        hits[:len(models)] = torch.tensor(models, dtype=torch.int64, device="cuda")
    found = torch.tensor([1 if models else 0], device="cuda")
    stats = torch.tensor([st.n_conflict, st.n_models, st.n_undecided, st.n_propagated, st.n_batches], dtype=torch.int64,
                         device="cuda")
    all_hits = [hits]
    if dist is not None:
        dist.all_reduce(found, op=dist.ReduceOp.MAX)   # found flag
        all_hits = [torch.empty_like(hits) for _ in range(world)]
        dist.all_gather(all_hits, hits)                # the (rare) satisfying assignment indices
        dist.all_reduce(stats)
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    times.append(float(tt.item()))
    dt = statistics.median(times)
    out = None
    if rank == 0:
        gathered = sorted(int(x) for h in all_hits for x in h.tolist() if x >= 0)
        n_conf, n_mod, n_und, n_prop, n_bat = (int(x) for x in stats.tolist())
        info = db.info()
        # device brute force on a 2^20 slice around the secret vs the reference BCP on the host cores
        sl = args.solve_check_bits
        first = (idx >> sl) << sl
        b = api.Batch(db, [api.Point(dset, 1 << sl, api.MODE_ENUM, first_sample=first)], api.OUT_FLAGS)
        for _ in range(2):
            c = b.run()[0]
        f_ms, p_ms = b.last_ms()
        conflict, full = b.flags(0)
        b.close()
        cores = os.cpu_count() or 1
        cuts = [first + (1 << sl) * i // cores for i in range(cores + 1)]
        res = cpu_pool().map(_cpu_enum_worker, [(cpu_kind(), args.a52_known, cuts[i], cuts[i + 1] - cuts[i])
                                                for i in range(cores)])
        cpu_conf = np.concatenate([np.unpackbits(r[2], bitorder="little")[:cuts[i + 1] - cuts[i]] for i, r in enumerate(res)]).astype(bool)
        cpu_full = np.concatenate([np.unpackbits(r[3], bitorder="little")[:cuts[i + 1] - cuts[i]] for i, r in enumerate(res)]).astype(bool)
        cpu_rate = (1 << sl) / max(r[0] for r in res)
        out = {"instance": "A5/2 (generated), %d vars / %d clauses, keystream %d + R4 + %d state bits known" %
                           (nv, len(cl), 114, args.a52_known),
               "set_size": k, "assignments": 1 << k, "ranks": world, "upload_s": upload_s,
               "live_vars": info.n_live_vars, "gates": info.n_gates, "warps_per_cta": info.warps_per_cta,
               "seconds": dt, "seconds_all": times, "assignments_per_s": (1 << k) / dt,
               "found_flag": int(found.item()), "hit_indices": gathered, "secret_index": idx,
               "secret_found": idx in gathered,
               "conflicts": n_conf, "models": n_mod, "undecided": n_und,
               "all_decided": n_conf + n_mod + n_und == (1 << k),
               "samples_propagated": n_prop, "device_batches": n_bat,
               "pruning": "BCP after every block of assigned variables; %.2e of the cube is ever propagated" % (n_prop / (1 << k)),
               "slice_check": {"bits": sl, "first": first, "device_brute_force_assignments_per_s": (1 << sl) / ((f_ms + p_ms) * 1e-3),
                               "conflict_mismatches": int((conflict != cpu_conf).sum()),
                               "model_mismatches": int((full != cpu_full).sum()),
                               "cpu_assignments_per_s": cpu_rate, "cpu_cores": cores,
                               "cpu_brute_force_2^%d_hours" % k: (1 << k) / cpu_rate / 3600.0}}
    if dist is not None:
        dist.barrier()
    db.close()
    return out


def leg_odls(api, torch, args):
    """BASELINE config 5: ODLS order-10 pair CNF (434 440 clauses: the clause database is far
    larger than shared memory and is read through L2; 10-literal clauses are visited lazily).
    Predict over 8 cells of square A (80 variables, one random value per cell and sample); the
    first `check` samples are compared with the reference BCP sample by sample."""
    from pdsat_b200 import fixtures as fx
    from oracle import oracle as orc
    onv, ocl = fx.odls_pair_cnf(10)
    ooffs, olits = fx.to_csr(ocl)
    t0 = time.perf_counter()
    odb = api.ClauseDb(onv, ooffs, olits, device=torch.cuda.current_device())
    upload_s = time.perf_counter() - t0
    cells = [(1, 1), (1, 2), (2, 3), (3, 4), (4, 6), (5, 7), (6, 8), (8, 9)]
    ovars = [fx.odls_var(10, 0, i, j, z) for (i, j) in cells for z in range(10)]
    n_o = args.odls_samples
    rng = np.random.default_rng(3)
    pick = rng.integers(0, 10, size=(n_o, len(cells)))
    vals = np.zeros((n_o, len(ovars)), dtype=np.uint8)
    for ci in range(len(cells)):
        vals[np.arange(n_o), ci * 10 + pick[:, ci]] = 1
    bo = api.Batch(odb, [api.Point(ovars, n_o, api.MODE_EXPLICIT, explicit_values=vals)], api.OUT_FLAGS | api.OUT_TRAIL)
    bo.fill()
    ms = []
    for it in range(6):
        bo.propagate()
        if it:
            ms.append(bo.last_ms()[1])
    co = bo.costs()[0]
    conflict, full = bo.flags(0)
    trail = bo.trail(0)
    bo.close()
    oi = odb.info()
    n_chk = min(args.odls_check, n_o)
    O = orc.Oracle(onv, ooffs, olits, cpu_kind())
    bad = 0
    t0 = time.perf_counter()
    for s_ in range(n_chk):
        r, _ = O.bcp([2 * (v - 1) + (0 if vals[s_][j] else 1) for j, v in enumerate(ovars)], want_assigns=False)
        if bool(conflict[s_]) != bool(r.conflict) or (not r.conflict and int(trail[s_]) != r.trail_size):
            bad += 1
    cpu_s = time.perf_counter() - t0
    odb.close()
    med = statistics.median(ms)
    return {"instance": "ODLS order 10: %d vars, %d clauses (10-literal clauses lazily), %d gates" % (onv, len(ocl), oi.n_gates),
            "upload_s": upload_s, "samples": n_o, "set_size": len(ovars), "propagate_ms": med,
            "samples_per_s": n_o / (med * 1e-3), "conflict_frac": co.n_conflict / n_o,
            "mean_trail_conflict_free": co.sum_trail / max(1, n_o - co.n_conflict),
            "literal_visits_per_group": co.queue_pops / ((n_o + 63) // 64),
            "warps_per_cta": oi.warps_per_cta, "index_blob_in_smem": bool(oi.db_in_smem),
            "parity": {"samples": n_chk, "mismatches": bad, "cpu_bcp_only_samples_per_s_1core": n_chk / cpu_s}}


def load_profile_metrics():
    tp = os.path.join(ROOT, "profiles", "bcp_kernel_metrics.json")
    return json.load(open(tp)) if os.path.exists(tp) else {}


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    from pdsat_b200 import api, build
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    if rank == 0:
        build.build()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    api.lib()

    nv, cl, offs, lits, stream = instance()
    db = api.ClauseDb(nv, offs, lits, device=local_rank)
    db.set_fixed_assumptions(stream)
    # run the library on a torch-owned (non-default) stream so that torch.cuda.Event pairs
    # recorded below sit on the stream the kernels are launched on
    stream_t = torch.cuda.Stream()
    torch.cuda.set_stream(stream_t)
    db.set_stream(stream_t.cuda_stream)
    n = args.samples
    # One point, N*n samples of ONE mt19937 stream per step; rank r owns the contiguous slice
    # [r*n, (r+1)*n) and reaches it by jump-ahead (results do not depend on N).
    seed0 = 1
    first = rank * n
    batch = api.Batch(db, [api.Point(SET_VARS, n, api.MODE_MT19937, seed=seed0, first_sample=first)])
    cost_t = torch.zeros(api.COST_WORDS, dtype=torch.int64, device="cuda")
    if dist is not None:  # NCCL-mode all-reduce works on a caller-owned tensor; alone the library keeps its own
        batch.set_cost_buffer(cost_t.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # N > 1: the per-point cost all-reduce is fused into the BCP kernel over NVLink peer memory
    # (CUDA IPC handles exchanged once through torch.distributed); --reduce nccl keeps the
    # separate NCCL all-reduce for comparison
    reduce_mode = "none"
    if dist is not None:
        reduce_mode = connect_peers(torch, dist, batch, rank, world, args.reduce)
    use_nccl = reduce_mode == "nccl"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-timed: inputs resident in HBM.  The steps rotate over `n_rot` resident
    # batches (same set, different seeds) whose assumption words together are > 2x the L2, so every
    # launch reads its words from HBM; ONE event pair brackets all K launches (no flush inside).
    words_bytes = ((n + 63) // 64) * ((len(SET_VARS) + 1) & ~1) * 8
    n_rot = max(2, min(16, -(-(2 * L2_BYTES) // words_bytes) + 1))
    rot = [(batch, cost_t)]
    for r in range(1, n_rot):
        b_r = api.Batch(db, [api.Point(SET_VARS, n, api.MODE_MT19937, seed=seed0 + r, first_sample=first)])
        c_r = torch.zeros(api.COST_WORDS, dtype=torch.int64, device="cuda")
        if dist is not None:
            b_r.set_cost_buffer(c_r.data_ptr())
        if dist is not None:
            assert connect_peers(torch, dist, b_r, rank, world, args.reduce) == reduce_mode
        rot.append((b_r, c_r))
    for b_r, _ in rot:
        b_r.fill()
        b_r.set_timing(False)  # no event records between the back-to-back launches
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    for w in range(max(args.warmup, n_rot)):
        b_r, c_r = rot[w % n_rot]
        b_r.propagate()
        if use_nccl:
            dist.all_reduce(c_r)
    barrier()
    sampler.start()
    launches1 = sum(b_r.launch_count() for b_r, _ in rot)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    # a spin kernel in front of the timed region keeps the stream busy while the host queues the K
    # launches: the timed region then measures the device, not the launch rate of this process
    with torch.cuda.stream(stream_t):
        torch.cuda._sleep(int(2000 * (150 + 15 * min(args.steps, 600))))  # ~2000 cycles per us
    ev0.record()
    for s in range(args.steps):
        b_r, c_r = rot[s % n_rot]
        b_r.propagate()
        if use_nccl:
            dist.all_reduce(c_r)
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = ev0.elapsed_time(ev1)
    gpu_launches = sum(b_r.launch_count() for b_r, _ in rot) - launches1
    kernel_ms = dev_ms / args.steps  # launch-to-launch average over the timed region (includes the gaps)
    # one isolated launch after an L2 flush, bracketed by its own event pair (launch latency included)
    batch.set_timing(True)
    flush.zero_()
    batch.propagate()
    _, kernel_ms_isolated = batch.last_ms()
    costs_host = batch.costs()[0]
    n_total_check = costs_host.n
    for b_r, _ in rot[1:]:
        assert b_r.costs()[0].n == n_total_check
    if dist is not None:
        dist.barrier()
    for b_r, _ in rot[1:]:
        b_r.close()

    # ---------------- e2e: public call sequence, host descriptors in, costs out
    for w in range(max(2, args.warmup // 2)):  # 2: a rank's stream offset gets its composite jump polynomial on its second appearance
        batch.reseed(0, seed0 + 7919 * (w + 1), first)
        batch.fill()
        batch.propagate()
        if use_nccl:
            dist.all_reduce(cost_t)
        batch.costs()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        batch.reseed(0, seed0 + 104729 * (s + 1), first)
        batch.fill()
        batch.propagate()
        if use_nccl:
            dist.all_reduce(cost_t)
        c = batch.costs()[0]
    barrier()
    e2e_serial_s = time.perf_counter() - t0
    n_reduced_check = c.n
    fill_ms_last, _ = batch.last_ms()
    h2d_per_step = batch.h2d_bytes()

    # ---------------- e2e, pipelined: the same call sequence with `depth` batches in flight (one
    # database handle + stream each, round robin); every step still uploads its own descriptors and
    # reads its own 64-byte result back, the read of step s waits only for step s.  The stages of a
    # step (jump tree, bit generator, transpose, BCP) are latency-bound kernels with small grids:
    # steps that overlap fill each other's idle issue slots.
    depth = max(1, args.pipeline if dist is None else min(args.pipeline, int(os.environ.get("PDSAT_BENCH_MAX_DEPTH", "8"))))
    lanes = []
    for d in range(depth):
        if d == 0:
            lanes.append((db, batch, cost_t))
            continue
        db_d = api.ClauseDb(nv, offs, lits, device=local_rank)
        db_d.set_fixed_assumptions(stream)
        st_d = torch.cuda.Stream()
        db_d.set_stream(st_d.cuda_stream)
        b_d = api.Batch(db_d, [api.Point(SET_VARS, n, api.MODE_MT19937, seed=seed0, first_sample=first)])
        c_d = torch.zeros(api.COST_WORDS, dtype=torch.int64, device="cuda")
        if dist is not None:
            b_d.set_cost_buffer(c_d.data_ptr())
        if dist is not None:
            m_d = connect_peers(torch, dist, b_d, rank, world, args.reduce)
            assert m_d == reduce_mode
        lanes.append((db_d, b_d, c_d, st_d))

    def pipelined(n_steps, seed_mul):
        last = None
        for s in range(n_steps):
            ln = lanes[s % depth]
            if s >= depth:
                last = ln[1].costs()[0]  # result of the step that used this lane `depth` steps ago
            ln[1].reseed(0, seed0 + seed_mul * (s + 1), first)
            ln[1].fill()
            ln[1].propagate()
            if use_nccl:
                with torch.cuda.stream(ln[3] if len(ln) > 3 else stream_t):
                    dist.all_reduce(ln[2])
        for s in range(max(0, n_steps - depth), n_steps):
            last = lanes[s % depth][1].costs()[0]
        return last

    pipelined(2 * depth, 7919)
    barrier()
    t0 = time.perf_counter()
    c_p = pipelined(args.steps, 15485863)
    barrier()
    e2e_s = time.perf_counter() - t0
    n_pipe_check = c_p.n
    # the timed regions above last milliseconds; keep the same fill+propagate load running for
    # ~1.5 s more so that nvidia-smi (100 ms period) sees the clocks under this load
    t_probe = time.perf_counter()
    probe_iters = torch.tensor([0], device="cuda")
    while True:  # same number of (collective) propagate calls on every rank
        for _ in range(20):
            batch.fill()
            batch.propagate()
        torch.cuda.synchronize()
        probe_iters[0] = 1 if time.perf_counter() - t_probe < 1.5 else 0
        if dist is not None:
            dist.all_reduce(probe_iters, op=dist.ReduceOp.MIN)
        if int(probe_iters.item()) == 0:
            break
    clocks = sampler.stop()
    clocks["note"] = "sampled across the timed regions plus 1.5 s of the same fill+propagate load"
    if dist is not None:
        dist.barrier()
    for ln in lanes[1:]:
        ln[1].close()
    if dist is not None:
        dist.barrier()
    for ln in lanes[1:]:
        ln[0].close()
    batch.close()

    # ---------------- the same e2e sequence with the declared counter-based generator
    # (PDSAT_MODE_RANDOM_COUNTER: NOT the reference's stream; for callers that only need i.i.d. bits)
    bc = api.Batch(db, [api.Point(SET_VARS, n, api.MODE_COUNTER, seed=seed0, first_sample=(first + 63) // 64 * 64)])
    cost_c = torch.zeros(api.COST_WORDS, dtype=torch.int64, device="cuda")
    if dist is not None:
        bc.set_cost_buffer(cost_c.data_ptr())
    mode_c = "none"
    if dist is not None:
        mode_c = connect_peers(torch, dist, bc, rank, world, args.reduce)
    for w in range(3):
        bc.reseed(0, seed0 + w, (first + 63) // 64 * 64)
        bc.fill()
        bc.propagate()
        if mode_c == "nccl":
            dist.all_reduce(cost_c)
        bc.costs()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        bc.reseed(0, seed0 + 104729 * (s + 1), (first + 63) // 64 * 64)
        bc.fill()
        bc.propagate()
        if mode_c == "nccl":
            dist.all_reduce(cost_c)
        cc = bc.costs()[0]
    barrier()
    e2e_counter_s = time.perf_counter() - t0
    fill_c_ms, _ = bc.last_ms()
    if dist is not None:
        dist.barrier()
    bc.close()

    prof = load_profile_metrics()
    info = db.info()
    extra = {}
    parity = None
    cascade = {}
    if rank == 0 and not args.no_extra:
        # ---------------- per-sample parity of the WHOLE headline batch
        parity, _ = gpu_vs_cpu_per_sample(api, db, SET_VARS, seed0, n, first_sample=first)
        # ---------------- cascade sets: where BCP actually propagates.  k = 86: ~40 % conflicts,
        # ~60 % of the samples imply literals; k >= 90: every sample conflicts (at different depths)
        for k in args.cascade:
            sv = [177 - i for i in range(k)]
            nk = args.cascade_samples
            bk = api.Batch(db, [api.Point(sv, nk, api.MODE_MT19937, seed=seed0)])
            bk.fill()
            for _ in range(3):
                flush.zero_()
                bk.propagate()
            ms = []
            for _ in range(5):
                flush.zero_()
                bk.propagate()
                ms.append(bk.last_ms()[1])
            ck = bk.costs()[0]
            fk, _ = bk.last_ms()
            bk.close()
            par_k, _ = gpu_vs_cpu_per_sample(api, db, sv, seed0, nk)
            med = statistics.median(ms)
            n_ok = nk - ck.n_conflict
            # literals BCP handled per sample: the k assumptions + what it implied (the level-0
            # literals folded at upload are NOT counted); conflict samples: assumptions only
            lits = (ck.sum_trail - n_ok * info.n_base_assigned) + ck.n_conflict * k
            weq = work_equivalent_bytes(nv, cl, stream, sv, seed0) if n_ok else None
            entry = {"samples": nk, "propagate_ms": med, "propagate_ms_all": ms, "fill_ms": fk,
                     "samples_per_s": nk / (med * 1e-3), "conflict_frac": ck.n_conflict / nk,
                     "implied_any_frac": ck.n_implied_any / nk,
                     "mean_implied_per_ok_sample": ((ck.sum_trail / n_ok) - info.n_base_assigned - k) if n_ok else None,
                     "propagated_literals_per_s": lits / (med * 1e-3),
                     "literal_visits_per_group": ck.queue_pops / ((nk + 63) // 64),
                     "parity": par_k,
                     "speedup_vs_cpu_bcp_only": (nk / (med * 1e-3)) / par_k["cpu_bcp_only_samples_per_s"],
                     "work_equivalent": None if weq is None else {
                         "bytes_per_ok_sample": weq, "gbs": weq * n_ok / (med * 1e-3) / 1e9,
                         "note": "SURVEY.md 8(d) formula (what a per-sample watched-literal BCP touches); "
                                 "a work measure, not a bandwidth"},
                     "ncu": prof.get("k%d" % k)}
            cascade["k%d" % k] = entry
        extra["cascade"] = cascade
        extra["odls10_predict"] = leg_odls(api, torch, args)
    # ---------------- configs 3 and 4 at every N (collective legs: all ranks take part)
    if not args.no_extra:
        extra_tabu = leg_tabu(api, torch, dist, db, rank, world, args)
        extra_solve = leg_solve(api, torch, dist, rank, world, args)
        torch.cuda.set_stream(stream_t)
        if rank == 0:
            extra["tabu_step"] = extra_tabu
            extra["solve_a52"] = extra_solve

    t = torch.tensor([dev_ms, e2e_s * 1e3, float(kernel_ms), e2e_counter_s * 1e3, e2e_serial_s * 1e3], dtype=torch.float64,
                     device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, kernel_ms, e2e_counter_ms, e2e_serial_ms = (float(x) for x in t.tolist())

    if rank == 0:
        total = n * world * args.steps
        value = total / (dev_ms * 1e-3)
        e2e_value = total / (e2e_ms * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = json.load(open(peaks_path))["hbm_gbs"]
            peak_src = "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # Roofline of the timed kernel on the bytes it must move: the bit-sliced assumption words
        # (k rounded up to even, 8 bytes per 64 samples each) in, 64 bytes of accumulators out.
        kp = (len(SET_VARS) + 1) & ~1
        groups = (n + 63) // 64
        alg_bytes = groups * kp * 8 + 64
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        ph = prof.get("headline_k30", {})
        traffic = ph.get("dram_bytes_per_sample")
        traffic = traffic * n if traffic else None
        true_lits = [int(l >> 1) + 1 if (l & 1) == 0 else -(int(l >> 1) + 1) for l in stream] + SET_VARS
        weq = algorithmic_bytes_per_sample(nv, cl, true_lits)
        line = {
            "metric": "decomp-set samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(n),
            # per sample BCP handles the 30 assumed literals and implies nothing on this set; the 200
            # keystream literals are folded once at upload and are not counted
            "propagated_literals_per_s": value * len(SET_VARS),
            "e2e": {"value": e2e_value, "unit": "samples/s",
                    "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": 8 * api.COST_WORDS,
                    "ms_per_step": e2e_ms / args.steps, "fill_ms": fill_ms_last,
                    "pipeline_depth": depth, "n_check": int(n_pipe_check),
                    "serial": {"value": total / (e2e_serial_ms * 1e-3), "ms_per_step": e2e_serial_ms / args.steps,
                               "note": "one step at a time: reseed -> fill -> propagate -> costs() before the next"},
                    "generator": "mt19937 + bool_rand in the reference's 1-worker order (device jump-ahead)",
                    "note": "inputs are descriptors (set, seed, window): sample values are generated on the "
                            "device; H2D = point descriptors + generator seed state + jump plan.  `pipeline_depth` "
                            "steps are in flight (one handle + stream each); every step uploads its own descriptors "
                            "and reads its own result",
                    "counter_generator": {"value": total / (e2e_counter_ms * 1e-3), "ms_per_step": e2e_counter_ms / args.steps,
                                          "fill_ms": fill_c_ms, "n_check": int(cc.n),
                                          "note": "same call sequence with PDSAT_MODE_RANDOM_COUNTER (declared "
                                                  "splitmix64 generator, not the reference stream)"}},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": alg_bytes, "bytes_per_sample": alg_bytes / n,
                         "kernel": "bcp_kernel", "kernel_ms": kernel_ms, "kernel_ms_isolated": float(kernel_ms_isolated),
                         "ncu": ph or None,
                         "work_equivalent": {"bytes_per_sample": weq, "gbs": weq * n / (kernel_ms * 1e-3) / 1e9,
                                             "note": "SURVEY.md 8(d) formula for T = 230 true literals: what a "
                                                     "per-sample watched-literal BCP touches - a work measure"},
                         "note": "achieved = bytes the kernel must read (bit-sliced assumption words) / average "
                                 "launch-to-launch time over the timed region (one CUDA-event pair around the K "
                                 "back-to-back launches, inputs rotating over batches larger than L2); "
                                 "kernel_ms_isolated = one launch after an L2 flush inside its own event pair.  On this set the first-wave candidate list is empty: the kernel "
                                 "streams the words and accounts 64 samples per group; the sets on which BCP "
                                 "propagates are in extra.cascade (DESIGN.md section 5)"},
            "clocks": clocks,
            "kernel_ms_last": kernel_ms, "fill_ms_last": fill_ms_last, "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "check": {"n": int(n_total_check), "n_conflict": int(costs_host.n_conflict),
                      "sum_trail": int(costs_host.sum_trail), "n_after_reduce": int(n_reduced_check)},
            "parity": parity,
            "reduce": reduce_mode,
            "extra": extra,
            "launch": {"grid": info.grid, "warps_per_cta": info.warps_per_cta, "smem_bytes": info.smem_bytes,
                       "live_vars": info.n_live_vars, "live_clauses": info.n_clauses_live, "gates": info.n_gates,
                       "index_blob_bytes": info.blob_bytes, "index_blob_in_smem": bool(info.db_in_smem)},
        }
        if world == 1 and not args.no_cpu:
            cb = cpu_arm(args.cpu_samples, "faithful")
            cb2 = cpu_arm(args.cpu_samples * 400, "bcp")
            cb["bcp_only_value"] = cb2["value"]
            cb["bcp_only_sample"] = cb2["sample"]
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()  # nobody frees a peer-mapped buffer while another rank may still use it
    db.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def connect_peers(torch, dist, batch, rank, world, want):
    """Exchange the CUDA IPC handles of the ranks' reduce buffers; returns "peer" or "nccl"."""
    mode = "nccl"
    if want in ("peer", "auto"):
        try:
            mine = torch.frombuffer(bytearray(batch.peer_export()), dtype=torch.uint8).cuda()
            allh = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            batch.peer_connect([bytes(h.cpu().numpy().tobytes()) for h in allh], rank)
            mode = "peer"
        except Exception as e:  # IPC unavailable: fall back to NCCL, and say so
            if rank == 0:
                print("peer reduce unavailable (%s); using NCCL all-reduce" % e, file=sys.stderr)
    ok = torch.tensor([1 if mode == "peer" else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if mode == "peer" and int(ok.item()) == 0:
        batch.peer_disconnect()
        mode = "nccl"
    dist.barrier()
    return mode


def close_pool():
    global _POOL
    if _POOL is not None:
        _POOL.close()
        _POOL.join()
        _POOL = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=10_000_000, help="samples per GPU per step")
    ap.add_argument("--cpu-samples", type=int, default=1000, help="CPU-arm samples per core per step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the parity / cascade / multi-GPU workload legs")
    ap.add_argument("--cascade", type=lambda v: [int(x) for x in v.split(",") if x], default=[86, 90, 100, 120, 177],
                    help="set sizes (bivium_Ending2 prefixes) of the timed + parity-checked cascade legs")
    ap.add_argument("--cascade-samples", type=int, default=1_000_000)
    ap.add_argument("--pipeline", type=int, default=8, help="steps in flight in the e2e measurement (1 GPU; same for N > 1; PDSAT_BENCH_MAX_DEPTH caps it)")
    ap.add_argument("--tabu-samples", type=int, default=1_000_000, help="samples per point of the tabu-step leg")
    ap.add_argument("--odls-samples", type=int, default=1_000_000)
    ap.add_argument("--odls-check", type=int, default=20_000, help="ODLS samples compared with the reference BCP")
    ap.add_argument("--a52-known", type=int, default=32, help="known R1|R2|R3 bits of the A5/2 solve leg (set = 64 - known)")
    ap.add_argument("--solve-check-bits", type=int, default=18, help="log2 size of the slice checked against the CPU")
    ap.add_argument("--reduce", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: fused NVLink peer-memory reduce (auto/peer; falls back to NCCL if CUDA IPC "
                         "is unavailable) or the separate NCCL all-reduce (DESIGN.md section 6)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        close_pool()


if __name__ == "__main__":
    main()
