"""ctypes binding of libpdsat_b200.so (the C-ABI in include/pdsat_cuda.h).

This is plumbing for tests and bench.py; the product is the CUDA library.  There is no
CPU fallback: if the shared library is missing, or no CUDA device is present, the calls
raise.  Nothing in this package imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpdsat_b200.so")

PDSAT_OK, PDSAT_ERR_ARG, PDSAT_ERR_CUDA, PDSAT_ERR_UNSAT, PDSAT_ERR_LIMIT = 0, 1, 2, 3, 4
MODE_MT19937, MODE_COUNTER, MODE_ENUM, MODE_EXPLICIT, MODE_ENUM_BLOCKS = 0, 1, 2, 3, 4
OUT_FLAGS, OUT_TRAIL, OUT_ASSIGN = 1, 2, 4
COST_WORDS = 8

EXPORTS = [
    "pdsat_last_error", "pdsat_device_count", "pdsat_upload_cnf", "pdsat_set_fixed_assumptions",
    "pdsat_db_get_info", "pdsat_db_base_assignment", "pdsat_set_stream", "pdsat_destroy",
    "pdsat_batch_create", "pdsat_batch_reseed", "pdsat_batch_fill", "pdsat_batch_propagate",
    "pdsat_batch_sync", "pdsat_batch_last_ms", "pdsat_batch_set_timing", "pdsat_batch_launch_count", "pdsat_batch_cost_device_ptr",
    "pdsat_batch_set_cost_buffer", "pdsat_batch_costs", "pdsat_batch_flags", "pdsat_batch_trail",
    "pdsat_batch_assign", "pdsat_batch_models", "pdsat_batch_destroy", "pdsat_eval_batch",
    "pdsat_mt19937_bool_rand", "pdsat_batch_peer_export", "pdsat_batch_peer_connect",
    "pdsat_batch_peer_disconnect", "pdsat_batch_h2d_bytes", "pdsat_db_gates", "pdsat_batch_set_first_draw",
    "pdsat_enumerate", "pdsat_batch_set_blocks",
    "pdsat_cdcl_solve_many", "pdsat_batch_handoff", "pdsat_batch_sample_values",
]


class PdsatError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("pdsat error %d: %s" % (code, msg))
        self.code = code


class CPoint(C.Structure):
    _fields_ = [("set_vars", C.c_void_p), ("k", C.c_int32), ("mode", C.c_int32), ("seed", C.c_uint64),
                ("first_sample", C.c_uint64), ("n_samples", C.c_uint64), ("explicit_values", C.c_void_p),
                ("first_draw", C.c_uint64), ("block_prefix", C.c_void_p), ("block_bits", C.c_int32)]


class CCost(C.Structure):
    _fields_ = [("n", C.c_uint64), ("n_conflict", C.c_uint64), ("n_full", C.c_uint64), ("sum_trail", C.c_uint64),
                ("sumsq_trail", C.c_uint64), ("n_implied_any", C.c_uint64), ("queue_pops", C.c_uint64),
                ("implied_lits", C.c_uint64)]


class CDbInfo(C.Structure):
    _fields_ = [("n_vars", C.c_int32), ("n_live_vars", C.c_int32), ("n_clauses_in", C.c_int64),
                ("n_clauses_live", C.c_int64), ("n_lits_live", C.c_int64), ("n_base_assigned", C.c_int32),
                ("max_clause_len", C.c_int32), ("status", C.c_int32), ("warps_per_cta", C.c_int32),
                ("smem_bytes", C.c_int32), ("grid", C.c_int32), ("db_in_smem", C.c_int32), ("sm_count", C.c_int32),
                ("n_gates", C.c_int32), ("blob_bytes", C.c_int32), ("n_gate_clauses", C.c_int64)]


class CEnumStats(C.Structure):
    _fields_ = [("n_assignments", C.c_uint64), ("n_conflict", C.c_uint64), ("n_models", C.c_uint64),
                ("n_undecided", C.c_uint64), ("n_propagated", C.c_uint64), ("n_batches", C.c_uint64)]


class CCdclLimits(C.Structure):
    _fields_ = [("max_conflicts", C.c_int64), ("max_restarts", C.c_int64), ("max_seconds", C.c_double),
                ("n_threads", C.c_int32)]


class CCdclStats(C.Structure):
    _fields_ = [("n_problems", C.c_uint64), ("n_sat", C.c_uint64), ("n_unsat", C.c_uint64), ("n_unknown", C.c_uint64),
                ("decisions", C.c_uint64), ("propagations", C.c_uint64), ("conflicts", C.c_uint64),
                ("seconds", C.c_double), ("threads", C.c_int32)]


SAT, UNSAT, UNKNOWN = 0, 1, 2


def _limits(max_conflicts=-1, max_restarts=-1, max_seconds=-1.0, n_threads=0):
    return CCdclLimits(max_conflicts, max_restarts, max_seconds, n_threads)


class CGate(C.Structure):
    _fields_ = [("type", C.c_int32), ("parity", C.c_int32), ("m", C.c_int32), ("n_clauses", C.c_int32),
                ("x", C.c_int32 * 7), ("lb", C.c_int32), ("lc", C.c_int32)]


_lib = None


def lib():
    """Load the CUDA library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: build it with `python -m pdsat_b200.build` "
                           "(the CUDA path has no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.pdsat_last_error.restype = C.c_char_p
    L.pdsat_device_count.argtypes = [C.POINTER(C.c_int)]
    L.pdsat_upload_cnf.argtypes = [C.c_int, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.pdsat_set_fixed_assumptions.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.pdsat_db_get_info.argtypes = [C.c_void_p, C.POINTER(CDbInfo)]
    L.pdsat_db_base_assignment.argtypes = [C.c_void_p, C.c_void_p]
    L.pdsat_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.pdsat_destroy.argtypes = [C.c_void_p]
    L.pdsat_destroy.restype = None
    L.pdsat_batch_create.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CPoint), C.c_uint32, C.c_int32,
                                     C.POINTER(C.c_void_p)]
    L.pdsat_batch_reseed.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64]
    L.pdsat_batch_fill.argtypes = [C.c_void_p]
    L.pdsat_batch_propagate.argtypes = [C.c_void_p]
    L.pdsat_batch_sync.argtypes = [C.c_void_p]
    L.pdsat_batch_last_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.pdsat_batch_set_timing.argtypes = [C.c_void_p, C.c_int32]
    L.pdsat_batch_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.pdsat_batch_cost_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.pdsat_batch_set_cost_buffer.argtypes = [C.c_void_p, C.c_void_p]
    L.pdsat_batch_costs.argtypes = [C.c_void_p, C.POINTER(CCost)]
    L.pdsat_batch_flags.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.pdsat_batch_trail.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.pdsat_batch_assign.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p]
    L.pdsat_batch_models.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p]
    L.pdsat_batch_destroy.argtypes = [C.c_void_p]
    L.pdsat_batch_destroy.restype = None
    L.pdsat_eval_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CPoint), C.POINTER(CCost)]
    L.pdsat_mt19937_bool_rand.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
    L.pdsat_batch_peer_export.argtypes = [C.c_void_p, C.c_void_p]
    L.pdsat_batch_peer_connect.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.pdsat_batch_peer_disconnect.argtypes = [C.c_void_p]
    L.pdsat_batch_h2d_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.pdsat_db_gates.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CGate), C.POINTER(C.c_int32)]
    L.pdsat_batch_set_first_draw.argtypes = [C.c_void_p, C.c_int32, C.c_uint64]
    L.pdsat_enumerate.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_uint64,
                                  C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(CEnumStats)]
    L.pdsat_batch_set_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.pdsat_cdcl_solve_many.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.POINTER(CCdclLimits), C.c_void_p,
                                        C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64),
                                        C.POINTER(CCdclStats)]
    L.pdsat_batch_handoff.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CCdclLimits), C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(CCdclStats)]
    L.pdsat_batch_sample_values.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != PDSAT_OK:
        raise PdsatError(rc, lib().pdsat_last_error().decode())


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().pdsat_device_count(C.byref(n))
    return n.value if rc == PDSAT_OK else 0


@dataclass
class Point:
    """One decomposition set + its sample range (pdsat_point)."""
    set_vars: Sequence[int]
    n_samples: int
    mode: int = MODE_MT19937
    seed: int = 1
    first_sample: int = 0
    explicit_values: Optional[np.ndarray] = None  # [n_samples, k] of {0,1}
    first_draw: int = 0  # MT19937: draws the worker's generator has already made
    block_prefix: Optional[np.ndarray] = None  # ENUM_BLOCKS: one uint64 prefix per block
    block_bits: int = 0


@dataclass
class Cost:
    n: int
    n_conflict: int
    n_full: int
    sum_trail: int
    sumsq_trail: int
    n_implied_any: int
    queue_pops: int = 0
    implied_lits: int = 0


def _cpoints(points: Sequence[Point]):
    arr = (CPoint * len(points))()
    keep = []
    for i, p in enumerate(points):
        sv = np.ascontiguousarray(p.set_vars, dtype=np.int32)
        keep.append(sv)
        arr[i].set_vars = sv.ctypes.data
        arr[i].k = len(sv)
        arr[i].mode = p.mode
        arr[i].seed = p.seed
        arr[i].first_sample = p.first_sample
        arr[i].n_samples = p.n_samples
        arr[i].first_draw = p.first_draw
        arr[i].block_prefix = None
        arr[i].block_bits = p.block_bits
        if p.block_prefix is not None:
            bp = np.ascontiguousarray(p.block_prefix, dtype=np.uint64)
            keep.append(bp)
            arr[i].block_prefix = bp.ctypes.data
        if p.explicit_values is not None:
            ev = np.ascontiguousarray(p.explicit_values, dtype=np.uint8)
            assert ev.size == p.n_samples * len(sv)
            keep.append(ev)
            arr[i].explicit_values = ev.ctypes.data
        else:
            arr[i].explicit_values = None
    return arr, keep


def _costs(arr, n) -> List[Cost]:
    return [Cost(arr[i].n, arr[i].n_conflict, arr[i].n_full, arr[i].sum_trail, arr[i].sumsq_trail,
                 arr[i].n_implied_any, arr[i].queue_pops, arr[i].implied_lits) for i in range(n)]


class ClauseDb:
    """Clause database on one GPU (pdsat_db).  device=-1 gives a host-only handle used by
    the CPU tests of the upload logic; it cannot evaluate anything."""

    def __init__(self, n_vars: int, offsets: np.ndarray, lits: np.ndarray, device: int = 0):
        self.n_vars = int(n_vars)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        lits = np.ascontiguousarray(lits, dtype=np.int32)
        h = C.c_void_p()
        _check(lib().pdsat_upload_cnf(device, self.n_vars, len(offsets) - 1, offsets.ctypes.data, lits.ctypes.data,
                                      C.byref(h)))
        self.h = h

    @classmethod
    def from_clauses(cls, n_vars: int, clauses, device: int = 0) -> "ClauseDb":
        from .fixtures import to_csr
        offs, lits = to_csr(clauses)
        return cls(n_vars, offs, lits, device)

    def set_fixed_assumptions(self, lits: Sequence[int]) -> None:
        a = np.ascontiguousarray(lits, dtype=np.int32)
        _check(lib().pdsat_set_fixed_assumptions(self.h, a.ctypes.data if len(a) else None, len(a)))

    def info(self) -> CDbInfo:
        i = CDbInfo()
        _check(lib().pdsat_db_get_info(self.h, C.byref(i)))
        return i

    def gates(self) -> List[CGate]:
        n = C.c_int32(0)
        _check(lib().pdsat_db_gates(self.h, 0, None, C.byref(n)))
        arr = (CGate * max(n.value, 1))()
        _check(lib().pdsat_db_gates(self.h, n.value, arr, C.byref(n)))
        return [arr[i] for i in range(n.value)]

    def base_assignment(self) -> np.ndarray:
        a = np.empty(self.n_vars, dtype=np.uint8)
        _check(lib().pdsat_db_base_assignment(self.h, a.ctypes.data))
        return a

    def set_stream(self, stream_ptr: int) -> None:
        _check(lib().pdsat_set_stream(self.h, C.c_void_p(stream_ptr)))

    def enumerate(self, set_vars: Sequence[int], n_fixed_top: int = 0, fixed_top_value: int = 0,
                  stop_at_first_model: bool = False, max_survivors: int = 1 << 20):
        """BCP-pruned enumeration of the cube over set_vars (ascending): returns
        (survivor indices, is_model flags, CEnumStats)."""
        sv = np.ascontiguousarray(set_vars, dtype=np.int32)
        surv = np.zeros(max(max_survivors, 1), dtype=np.uint64)
        ism = np.zeros(max(max_survivors, 1), dtype=np.uint8)
        ns = C.c_uint64(0)
        st = CEnumStats()
        _check(lib().pdsat_enumerate(self.h, sv.ctypes.data, len(sv), n_fixed_top, fixed_top_value,
                                     1 if stop_at_first_model else 0, max_survivors, surv.ctypes.data, ism.ctypes.data,
                                     C.byref(ns), C.byref(st)))
        n = min(ns.value, max_survivors)
        return surv[:n].copy(), ism[:n].astype(bool), st

    def cdcl_solve_many(self, assumps: np.ndarray, max_models: int = 0, **limits):
        """Host CDCL on n problems of k assumption literals (rows of `assumps`, MiniSat codes):
        (status[n], props[n], model_problem[m], models[m, n_vars], CCdclStats)."""
        a = np.ascontiguousarray(assumps, dtype=np.int32)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        n, k = a.shape
        status = np.zeros(n, dtype=np.uint8)
        props = np.zeros(n, dtype=np.uint64)
        mp = np.zeros(max(max_models, 1), dtype=np.int64)
        models = np.zeros((max(max_models, 1), self.n_vars), dtype=np.uint8)
        nm = C.c_int64(0)
        st = CCdclStats()
        lim = _limits(**limits)
        _check(lib().pdsat_cdcl_solve_many(self.h, n, k, a.ctypes.data if a.size else None, C.byref(lim), status.ctypes.data,
                                           props.ctypes.data, max_models, mp.ctypes.data, models.ctypes.data, C.byref(nm),
                                           C.byref(st)))
        m = min(nm.value, max_models)
        return status, props, mp[:m], models[:m], st

    def eval_batch(self, points: Sequence[Point]) -> List[Cost]:
        """One call, host buffers in / out (pdsat_eval_batch): the e2e path."""
        arr, keep = _cpoints(points)
        out = (CCost * len(points))()
        _check(lib().pdsat_eval_batch(self.h, len(points), arr, out))
        return _costs(out, len(points))

    def close(self) -> None:
        if getattr(self, "h", None):
            lib().pdsat_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    def __init__(self, db: ClauseDb, points: Sequence[Point], out_flags: int = 0, max_models: int = 0):
        self.db = db
        self.points = list(points)
        arr, keep = _cpoints(self.points)
        h = C.c_void_p()
        _check(lib().pdsat_batch_create(db.h, len(self.points), arr, out_flags, max_models, C.byref(h)))
        self.h = h
        self.max_models = max_models

    def reseed(self, point: int, seed: int, first_sample: int) -> None:
        _check(lib().pdsat_batch_reseed(self.h, point, seed, first_sample))

    def set_first_draw(self, point: int, first_draw: int) -> None:
        _check(lib().pdsat_batch_set_first_draw(self.h, point, first_draw))

    def fill(self) -> None:
        _check(lib().pdsat_batch_fill(self.h))

    def propagate(self) -> None:
        _check(lib().pdsat_batch_propagate(self.h))

    def sync(self) -> None:
        _check(lib().pdsat_batch_sync(self.h))

    def run(self) -> List[Cost]:
        self.fill()
        self.propagate()
        return self.costs()

    def last_ms(self):
        f, p = C.c_float(0), C.c_float(0)
        _check(lib().pdsat_batch_last_ms(self.h, C.byref(f), C.byref(p)))
        return f.value, p.value

    def set_timing(self, on: bool) -> None:
        """stage timing (four event records per step on the launch stream) on / off"""
        _check(lib().pdsat_batch_set_timing(self.h, 1 if on else 0))

    def launch_count(self) -> int:
        n = C.c_int64(0)
        _check(lib().pdsat_batch_launch_count(self.h, C.byref(n)))
        return n.value

    def h2d_bytes(self) -> int:
        n = C.c_int64(0)
        _check(lib().pdsat_batch_h2d_bytes(self.h, C.byref(n)))
        return n.value

    def cost_device_ptr(self) -> int:
        p = C.c_void_p()
        _check(lib().pdsat_batch_cost_device_ptr(self.h, C.byref(p)))
        return p.value

    def set_cost_buffer(self, dptr: int) -> None:
        _check(lib().pdsat_batch_set_cost_buffer(self.h, C.c_void_p(dptr)))

    def peer_export(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's reduce buffer (exchange it with all ranks)."""
        buf = C.create_string_buffer(64)
        _check(lib().pdsat_batch_peer_export(self.h, buf))
        return buf.raw

    def peer_connect(self, handles: Sequence[bytes], my_rank: int) -> None:
        """After this, propagate() is collective and leaves the all-rank sums in the cost buffer."""
        blob = b"".join(handles)
        assert len(blob) == 64 * len(handles)
        _check(lib().pdsat_batch_peer_connect(self.h, len(handles), my_rank, blob))

    def peer_disconnect(self) -> None:
        _check(lib().pdsat_batch_peer_disconnect(self.h))

    def sample_values(self, point: int = 0, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        if n is None:
            n = self.points[point].n_samples - first
        v = np.zeros((n, len(self.points[point].set_vars)), dtype=np.uint8)
        _check(lib().pdsat_batch_sample_values(self.h, point, first, n, v.ctypes.data))
        return v

    def handoff(self, point: int = 0, max_models: int = 0, want_cost: bool = False, **limits):
        """Verdict of EVERY sample of the point: BCP's where it decides, the host CDCL's for the
        rest.  Returns (status[n], cost[n] | None, model_sample[m], models[m, n_vars], CCdclStats)."""
        n = self.points[point].n_samples
        status = np.zeros(n, dtype=np.uint8)
        cost = np.zeros(n, dtype=np.uint64) if want_cost else None
        ms = np.zeros(max(max_models, 1), dtype=np.uint64)
        models = np.zeros((max(max_models, 1), self.db.n_vars), dtype=np.uint8)
        nm = C.c_int64(0)
        st = CCdclStats()
        lim = _limits(**limits)
        _check(lib().pdsat_batch_handoff(self.h, point, C.byref(lim), status.ctypes.data,
                                         cost.ctypes.data if want_cost else None, max_models, ms.ctypes.data,
                                         models.ctypes.data, C.byref(nm), C.byref(st)))
        m = min(nm.value, max_models)
        return status, cost, ms[:m], models[:m], st

    def costs(self) -> List[Cost]:
        out = (CCost * len(self.points))()
        _check(lib().pdsat_batch_costs(self.h, out))
        return _costs(out, len(self.points))

    def flags(self, point: int = 0):
        """(conflict, full) boolean arrays of the point's samples."""
        n = self.points[point].n_samples
        g = (n + 63) // 64
        c = np.empty(g, dtype=np.uint64)
        f = np.empty(g, dtype=np.uint64)
        _check(lib().pdsat_batch_flags(self.h, point, c.ctypes.data, f.ctypes.data))
        cb = np.unpackbits(c.view(np.uint8), bitorder="little")[:n].astype(bool)
        fb = np.unpackbits(f.view(np.uint8), bitorder="little")[:n].astype(bool)
        return cb, fb

    def trail(self, point: int = 0) -> np.ndarray:
        t = np.empty(self.points[point].n_samples, dtype=np.uint32)
        _check(lib().pdsat_batch_trail(self.h, point, t.ctypes.data))
        return t

    def assign(self, point: int = 0, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        if n is None:
            n = self.points[point].n_samples - first
        a = np.empty((n, self.db.n_vars), dtype=np.uint8)
        _check(lib().pdsat_batch_assign(self.h, point, first, n, a.ctypes.data))
        return a

    def models(self):
        """(n_found, point_of, sample_of, models[stored, n_vars])"""
        cap = max(self.max_models, 1)
        pts = np.zeros(cap, dtype=np.int32)
        smp = np.zeros(cap, dtype=np.uint64)
        mod = np.zeros((cap, self.db.n_vars), dtype=np.uint8)
        n = C.c_int64(0)
        _check(lib().pdsat_batch_models(self.h, C.byref(n), pts.ctypes.data, smp.ctypes.data, mod.ctypes.data))
        stored = min(n.value, self.max_models)
        return n.value, pts[:stored], smp[:stored], mod[:stored]

    def close(self) -> None:
        if getattr(self, "h", None):
            lib().pdsat_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mt19937_bool_rand(seed: int, first: int, n: int, device: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    _check(lib().pdsat_mt19937_bool_rand(device, seed, first, n, out.ctypes.data))
    return out
