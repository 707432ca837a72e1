"""Host-side mirror (Python) of the reference's predict / solve interface for the hot path.

Used by tests and bench.py; the C++ host driver (pdsat_b200/host/) implements the same
logic natively.  Names follow the reference:

  MPI_Predicter::AllocatePredictArrays   src_mpi/mpi_predicter.cpp:2346-2367  -> samples_for_set
  MPI_Predicter::GetPredict              src_mpi/mpi_predicter.cpp:1904-2115  -> predict_from_cost
  MPI_Predicter::getCurPredictTime       src_mpi/mpi_predicter.cpp:1846-1901  -> predict_from_cost (prep)
  GetDeepPredictTasks (neighbourhood)    src_mpi/mpi_predicter.cpp:2926-2935  -> hamming1_neighbourhood
  MPI_Base::MakeVarChoose schemas        src_mpi/mpi_base.cpp:275-347          -> schema_vars

Multi-GPU: one process per GPU; every rank evaluates a contiguous slice of each point's
sample range (same mt19937 stream, reached by jump-ahead) and the per-point integer
accumulators are summed with ONE all-reduce.  Integer sums make the result
independent of the number of ranks.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import api

MAX_STEP_CNF_IN_SET = 30            # src_mpi/mpi_predicter.h:17
MIN_PERCENT_SOLVED_IN_TIME = 0.5    # src_mpi/mpi_predicter.h:29
HUGE_DOUBLE = np.longdouble("1e308")  # src_mpi/mpi_base.h:43


def schema_vars(schema: str, count: int, core_len: int = 177) -> List[int]:
    """Decomposition-set schemas of MakeVarChoose (sorted ascending, as the reference does)."""
    if schema in ("", "0", "bivium_Beginning1"):
        v = list(range(1, count + 1))
    elif schema == "bivium_Beginning2":
        v = [i + 84 + 1 for i in range(count)]
    elif schema == "bivium_Beginning_halved":
        v = [i + 1 for i in range(count // 2)] + [i + 84 + 1 for i in range(count // 2)]
    elif schema == "bivium_Ending1":
        v = [84 - i for i in range(count)]
    elif schema == "bivium_Ending2":
        v = [177 - i for i in range(count)]
    elif schema == "bivium_Ending_halved":
        v = [84 - i for i in range(count // 2)] + [177 - i for i in range(count // 2)]
    elif schema == "a5":
        v = [i + 1 for i in range(9)] + [i + 19 + 1 for i in range(11)] + [i + 41 + 1 for i in range(11)]
    else:
        raise ValueError("unknown schema " + schema)
    return sorted(v)


def samples_for_set(k: int, cnf_in_set_count: int) -> int:
    """AllocatePredictArrays: min(cnf_in_set_count, 2^k), 2^k only considered for k <= 30."""
    if k > MAX_STEP_CNF_IN_SET:
        return cnf_in_set_count
    return min(cnf_in_set_count, 1 << k)


def hamming1_neighbourhood(centre: Sequence[int], universe: Sequence[int]) -> List[List[int]]:
    """All sets at Hamming distance 1 from `centre` inside `universe`
    (GetDeepPredictTasks default strategy: flip each of core_len bits)."""
    cs = set(centre)
    out = []
    for v in universe:
        s = set(cs)
        if v in s:
            s.remove(v)
        else:
            s.add(v)
        if s:
            out.append(sorted(s))
    return out


def shard_range(n: int, rank: int, world: int):
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return lo, hi


@dataclass
class PointResult:
    set_vars: List[int]
    cost: api.Cost
    sum_cost: float                 # sum_time_arr (double)
    med: np.longdouble              # med_time_arr
    predict: np.longdouble          # predict_time_arr
    solved: int                     # samples decided by BCP (conflict or model)


def predict_from_cost(cost: api.Cost, k: int, evaluation_type: str = "prep", proc_count: int = 1):
    """The estimator arithmetic on the integer accumulators.

    "prep": getCurPredictTime with solved = #samples solved on preprocessing (exactly the
    reference's prep estimate).
    "bcp_trail": cost_j = BCP trail of sample j (= Solver::propagations of a sample BCP decides
    without conflict; 0 for conflict samples).  Every partial sum of a sequential double
    accumulation of integers < 2^53 is exact, so sum == float(integer sum) bit for bit.  This is
    a BCP-only cost, NOT the reference's "propagation" (CDCL propagations of the whole solve())."""
    if evaluation_type not in ("prep", "bcp_trail"):
        raise ValueError("evaluation_type must be 'prep' or 'bcp_trail' (the CDCL-based 'propagation' cost lives in the "
                         "C++ host driver, pdsat_b200/host/pd_sat_main.cpp)")
    n = cost.n
    if evaluation_type == "prep":
        solved = cost.n_conflict + cost.n_full
        percent = float(solved) * 100 / float(n)
        if percent <= MIN_PERCENT_SOLVED_IN_TIME:
            return float(cost.sum_trail), np.longdouble(0), HUGE_DOUBLE
        prob = np.longdouble(float(solved) / float(n))
        pred = np.longdouble(float(2.0 ** k) * 3.0) / prob
        return 1e-7 * n, np.longdouble(0), pred
    s = float(cost.sum_trail)
    med = np.longdouble(s / float(n))
    pred = med / np.longdouble(float(proc_count))
    pred = pred * np.longdouble(2.0 ** k)
    return s, med, pred


def sample_variance(cost: api.Cost) -> float:
    """s^2 of the conflict-free samples' cost (WritePredictToFile, mpi_predicter.cpp:2219-2224:
    sum (x - mean)^2 / (solved - 1)) from the integer sums: (n*Sxx - Sx^2) / (n*(n-1)), the
    numerator evaluated exactly in integers."""
    n = cost.n - cost.n_conflict
    if n < 2:
        return -1.0
    num = n * cost.sumsq_trail - cost.sum_trail * cost.sum_trail
    return float(num) / float(n * (n - 1))


class Predicter:
    """Scores decomposition sets on one GPU (or one rank of a multi-GPU job)."""

    def __init__(self, db: api.ClauseDb, cnf_in_set_count: int = 1000, evaluation_type: str = "prep",
                 proc_count: int = 1, base_seed: int = 1, rank: int = 0, world: int = 1, all_reduce=None,
                 exhaustive_small_sets: bool = False):
        self.db = db
        self.cnf_in_set_count = cnf_in_set_count
        self.evaluation_type = evaluation_type
        self.proc_count = proc_count
        self.base_seed = base_seed
        self.rank, self.world = rank, world
        self.all_reduce = all_reduce  # callable(np.ndarray[int64]) -> summed array, or None
        # the reference always DRAWS its samples; enumerating all 2^k assignments of a small set
        # instead is an option, not the default
        self.exhaustive_small_sets = exhaustive_small_sets

    def points_for(self, sets: Sequence[Sequence[int]]) -> List[api.Point]:
        pts = []
        for p, sv in enumerate(sets):
            n = samples_for_set(len(sv), self.cnf_in_set_count)
            lo, hi = shard_range(n, self.rank, self.world)
            mode = api.MODE_ENUM if (self.exhaustive_small_sets and n == (1 << len(sv)) and
                                     len(sv) <= MAX_STEP_CNF_IN_SET) else api.MODE_MT19937
            pts.append(api.Point(list(sv), max(hi - lo, 0), mode, seed=self.base_seed + p, first_sample=lo))
        return pts

    def evaluate(self, sets: Sequence[Sequence[int]]) -> List[PointResult]:
        pts = self.points_for(sets)
        live = [i for i, p in enumerate(pts) if p.n_samples > 0]
        costs = [api.Cost(0, 0, 0, 0, 0, 0) for _ in pts]
        if live:
            got = self.db.eval_batch([pts[i] for i in live])
            for i, c in zip(live, got):
                costs[i] = c
        if self.all_reduce is not None:
            arr = np.asarray([[c.n, c.n_conflict, c.n_full, c.sum_trail, c.sumsq_trail, c.n_implied_any,
                               c.queue_pops, c.implied_lits] for c in costs], dtype=np.int64)
            arr = self.all_reduce(arr)
            costs = [api.Cost(*[int(x) for x in row]) for row in arr]
        out = []
        for sv, c in zip(sets, costs):
            s, med, pred = predict_from_cost(c, len(sv), self.evaluation_type, self.proc_count)
            out.append(PointResult(list(sv), c, s, med, pred, c.n_conflict + c.n_full))
        return out
