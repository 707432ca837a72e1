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
def _cpu_worker(args):
    kind, mode, seed, n = args
    from oracle import oracle as orc
    nv, cl, offs, lits, stream = instance()
    O = orc.Oracle(nv, offs, lits, kind)
    A = orc.sample_assumptions(seed, SET_VARS, 0, n)
    st = list(stream)
    t0 = time.perf_counter()
    conf = 0
    if mode == "faithful":
        for s in range(n):
            r, _ = O.predict_sample(list(A[s]) + st, 0, want_assigns=False)
            conf += r.conflict
    else:
        O.add_units(st)
        for s in range(n):
            r, _ = O.bcp(A[s], want_assigns=False)
            conf += r.conflict
    return time.perf_counter() - t0, conf


def cpu_arm(n_per_core: int, mode: str = "faithful", cores: int | None = None):
    """Reference CPU path on all host cores (one process per core, static split — the
    reference's master/worker MPI only adds dispatch overhead)."""
    import multiprocessing as mp
    from oracle import oracle as orc
    orc.build()
    kind = "ref" if orc.have_ref() else "port"
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, mode, 1000 + i, n_per_core) for i in range(cores)])
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    return {"value": cores * n_per_core / busy, "unit": "samples/s", "cores": cores,
            "kind": "reference" if kind == "ref" else "port", "wall_s": wall,
            "sample": "%d samples/core x %d cores, %s" % (n_per_core, cores,
                      "fresh Solver + addProblem per sample (solverProgramCalling, eval=prep)" if mode == "faithful"
                      else "persistent clause DB, BCP only")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_per_core = args.cpu_samples
    for _ in range(args.warmup):
        cpu_arm(max(50, n_per_core // 10))
    vals = []
    t_all = 0.0
    cores = os.cpu_count() or 1
    last = None
    for _ in range(args.steps):
        last = cpu_arm(n_per_core)
        vals.append(last["value"])
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
            "l2": "flushed between timed steps (256 MiB write)"}


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
    batch.set_cost_buffer(cost_t.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # N > 1: the per-point cost all-reduce is fused into the BCP kernel over NVLink peer memory
    # (CUDA IPC handles exchanged once through torch.distributed); --reduce nccl keeps the
    # separate NCCL all-reduce for comparison
    reduce_mode = "none"
    if dist is not None:
        reduce_mode = "nccl"
        if args.reduce in ("peer", "auto"):
            try:
                mine = torch.frombuffer(bytearray(batch.peer_export()), dtype=torch.uint8).cuda()
                allh = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
                dist.all_gather(allh, mine)
                batch.peer_connect([bytes(h.cpu().numpy().tobytes()) for h in allh], rank)
                reduce_mode = "peer"
            except Exception as e:  # IPC unavailable: fall back to NCCL, and say so
                if rank == 0:
                    print("peer reduce unavailable (%s); using NCCL all-reduce" % e, file=sys.stderr)
        ok = torch.tensor([1 if reduce_mode == "peer" else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if reduce_mode == "peer" and int(ok.item()) == 0:
            batch.peer_disconnect()
            reduce_mode = "nccl"
        dist.barrier()
    use_nccl = reduce_mode == "nccl"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-timed: inputs resident in HBM
    batch.fill()
    torch.cuda.synchronize()
    launches0 = batch.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    for w in range(args.warmup):
        flush.zero_()
        batch.propagate()
        if use_nccl:
            dist.all_reduce(cost_t)
    barrier()
    sampler.start()
    launches1 = batch.launch_count()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()
        ev[s][0].record()
        batch.propagate()
        if use_nccl:
            dist.all_reduce(cost_t)
        ev[s][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(step_ms)
    gpu_launches = batch.launch_count() - launches1
    _, kernel_ms = batch.last_ms()
    costs_host = batch.costs()[0]
    n_total_check = costs_host.n

    # ---------------- e2e: public call sequence, host descriptors in, costs out
    for w in range(max(1, args.warmup // 2)):
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
    e2e_s = time.perf_counter() - t0
    n_reduced_check = c.n
    fill_ms_last, _ = batch.last_ms()
    h2d_per_step = batch.h2d_bytes()
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

    # ---------------- secondary workloads (rank 0 only, not part of the headline): a cascade-
    # heavy set (k = 86: ~40 % conflicts, ~60 % of the samples imply literals) and solve-mode
    # enumeration of 2^24 assignments of the weakened-Bivium stand-in
    extra = {}
    if rank == 0 and not args.no_extra:
        sv86 = [177 - i for i in range(86)]
        n86 = 2_000_000
        b86 = api.Batch(db, [api.Point(sv86, n86, api.MODE_MT19937, seed=1)])
        for _ in range(3):
            c86 = b86.run()[0]
        f86, p86 = b86.last_ms()
        extra["cascade_k86"] = {"samples": n86, "propagate_ms": p86, "fill_ms": f86,
                                "samples_per_s": n86 / (p86 * 1e-3), "conflict_frac": c86.n_conflict / n86,
                                "mean_trail_conflict_free": c86.sum_trail / max(1, n86 - c86.n_conflict),
                                "propagated_literals_per_s": c86.sum_trail / (p86 * 1e-3)}
        b86.close()
        secret = mt19937_bool_rand_numpy(SECRET_SEED, 177)
        db2 = api.ClauseDb(nv, offs, lits, device=local_rank)
        db2.set_stream(stream_t.cuda_stream)
        db2.set_fixed_assumptions(list(stream) + [2 * v + (0 if secret[v] else 1) for v in range(153)])
        sv_enum = list(range(154, 178))
        be = api.Batch(db2, [api.Point(sv_enum, 1 << 24, api.MODE_ENUM)], max_models=16)
        for _ in range(3):
            ce = be.run()[0]
        fe, pe = be.last_ms()
        found, _, smp, _ = be.models()
        expect = sum(int(secret[153 + r]) << r for r in range(24))
        extra["solve_enum_2^24"] = {"assignments": 1 << 24, "propagate_ms": pe, "fill_ms": fe,
                                    "assignments_per_s": (1 << 24) / ((pe + fe) * 1e-3), "conflicts": ce.n_conflict,
                                    "models": int(found), "secret_found": bool(found >= 1 and expect in [int(x) for x in smp])}
        be.close()
        db2.close()
        # tabu-search step (config 3 shape): centre {148..177} + its 177 Hamming-1 neighbours in
        # ONE batch, per-point mt19937 streams (seed = 1 + point), through the public call sequence
        from pdsat_b200 import predicter as pr
        sets = [SET_VARS] + pr.hamming1_neighbourhood(SET_VARS, list(range(1, 178)))
        n_pt = 200_000
        bn = api.Batch(db, [api.Point(sv, n_pt, api.MODE_MT19937, seed=1 + i) for i, sv in enumerate(sets)])
        bn.run()
        torch.cuda.synchronize()
        t_n = time.perf_counter()
        for _ in range(3):
            cn = bn.run()
        t_n = (time.perf_counter() - t_n) / 3
        fn, pn = bn.last_ms()
        extra["tabu_neighbourhood_178pts"] = {"points": len(sets), "samples_per_point": n_pt, "e2e_ms": t_n * 1e3,
                                              "fill_ms": fn, "propagate_ms": pn,
                                              "samples_per_s_e2e": len(sets) * n_pt / t_n,
                                              "points_per_s": len(sets) / t_n}
        bn.close()
        # ODLS order-10 pair CNF (config 5: clause DB >> shared memory, 10-literal clauses):
        # 8 cells of square A get a random value each (one-hot over the cell's 10 variables)
        from pdsat_b200 import fixtures as fx
        onv, ocl = fx.odls_pair_cnf(10)
        ooffs, olits = fx.to_csr(ocl)
        odb = api.ClauseDb(onv, ooffs, olits, device=local_rank)
        odb.set_stream(stream_t.cuda_stream)
        cells = [(1, 1), (1, 2), (2, 3), (3, 4), (4, 6), (5, 7), (6, 8), (8, 9)]
        ovars = [fx.odls_var(10, 0, i, j, z) for (i, j) in cells for z in range(10)]
        n_o = 100_000
        rng = np.random.default_rng(3)
        pick = rng.integers(0, 10, size=(n_o, len(cells)))
        vals = np.zeros((n_o, len(ovars)), dtype=np.uint8)
        for ci in range(len(cells)):
            vals[np.arange(n_o), ci * 10 + pick[:, ci]] = 1
        bo = api.Batch(odb, [api.Point(ovars, n_o, api.MODE_EXPLICIT, explicit_values=vals)])
        for _ in range(3):
            co = bo.run()[0]
        fo, po = bo.last_ms()
        oi = odb.info()
        extra["odls10_predict"] = {"samples": n_o, "set_size": len(ovars), "propagate_ms": po, "fill_ms": fo,
                                   "samples_per_s": n_o / (po * 1e-3), "conflict_frac": co.n_conflict / n_o,
                                   "mean_trail_conflict_free": co.sum_trail / max(1, n_o - co.n_conflict),
                                   "warps_per_cta": oi.warps_per_cta, "clauses": int(oi.n_clauses_live)}
        bo.close()
        odb.close()

    t = torch.tensor([dev_ms, e2e_s * 1e3, float(kernel_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, kernel_ms = (float(x) for x in t.tolist())

    if rank == 0:
        info = db.info()
        total = n * world * args.steps
        value = total / (dev_ms * 1e-3)
        e2e_value = total / (e2e_ms * 1e-3)
        # roofline: T = 30 set literals + 200 keystream literals, nothing implied
        true_lits = [int(l >> 1) + 1 if (l & 1) == 0 else -(int(l >> 1) + 1) for l in stream] + SET_VARS
        bytes_per_sample = algorithmic_bytes_per_sample(nv, cl, true_lits)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = json.load(open(peaks_path))["hbm_gbs"]
            peak_src = "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = bytes_per_sample * n / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "bcp_kernel_traffic.json")
        if os.path.exists(tp):
            per_sample = json.load(open(tp)).get("dram_bytes_per_sample")
            traffic = per_sample * n if per_sample else None
        line = {
            "metric": "decomp-set samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(n),
            "propagated_literals_per_s": value * (info.n_base_assigned + len(SET_VARS)),
            "e2e": {"value": e2e_value, "unit": "samples/s",
                    "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": 8 * api.COST_WORDS,
                    "note": "inputs are descriptors (set, seed, window): sample values are generated on the "
                            "device; H2D = point descriptors + generator seed state + jump plan"},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_sample": bytes_per_sample,
                         "dram_achieved_gbs": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None,
                         "note": "algorithmic bytes of SURVEY.md 8(d) = what a per-sample watched-literal BCP "
                                 "touches (T=230 literals); this kernel folds the 200 keystream units at upload, "
                                 "proves statically (first-wave candidate analysis) which clauses can fire and "
                                 "shares index reads across 64 samples, so frac >> 1; the physical HBM traffic is "
                                 "`traffic` (assumption words only) - see DESIGN.md"},
            "clocks": clocks,
            "kernel_ms_last": kernel_ms, "fill_ms_last": fill_ms_last, "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "check": {"n": int(n_total_check), "n_conflict": int(costs_host.n_conflict),
                      "sum_trail": int(costs_host.sum_trail), "n_after_reduce": int(n_reduced_check)},
            "reduce": reduce_mode,
            "extra": extra,
            "launch": {"grid": info.grid, "warps_per_cta": info.warps_per_cta, "smem_bytes": info.smem_bytes,
                       "live_vars": info.n_live_vars, "live_clauses": info.n_clauses_live},
        }
        if world == 1 and not args.no_cpu:
            cb = cpu_arm(args.cpu_samples, "faithful")
            cb2 = cpu_arm(args.cpu_samples * 40, "bcp")
            cb["bcp_only_value"] = cb2["value"]
            cb["bcp_only_sample"] = cb2["sample"]
            cb.pop("wall_s", None)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()  # nobody frees a peer-mapped buffer while another rank may still use it
    batch.close()
    db.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=10_000_000, help="samples per GPU per step")
    ap.add_argument("--cpu-samples", type=int, default=1500, help="CPU-arm samples per core per step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--reduce", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: fused NVLink peer-memory reduce (auto/peer; falls back to NCCL if CUDA IPC "
                         "is unavailable) or the separate NCCL all-reduce (DESIGN.md section 6)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
