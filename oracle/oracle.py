"""ctypes front-end of the two CPU checkers — TEST INFRASTRUCTURE ONLY.

`Oracle(kind="port")`  -> oracle/liboracle.so        (plain-C restatement, bcp_oracle.c)
`Oracle(kind="ref")`   -> oracle/_ref/libpdsat_ref.so (unmodified reference MiniSat)

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs.  The product package pdsat_b200 never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpdsat_ref.so")


class OrcResult(C.Structure):
    _fields_ = [("conflict", C.c_int32), ("full", C.c_int32), ("sat", C.c_int32), ("prep", C.c_int32),
                ("ret", C.c_int32), ("trail_size", C.c_int32), ("propagations", C.c_uint64),
                ("watch_scans", C.c_uint64), ("decisions", C.c_uint64), ("conflicts", C.c_uint64)]


def build(force: bool = False) -> None:
    """Compile the checkers (the port always; _ref only where /root/reference exists)."""
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(
            os.path.join(HERE, "bcp_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_libs = {}


def _lib(kind: str):
    if kind in _libs:
        return _libs[kind]
    path = PORT_SO if kind == "port" else REF_SO
    if kind == "port" and not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_ok.argtypes = [C.c_void_p]
    L.orc_add_units.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.orc_predict_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(OrcResult), C.c_void_p]
    L.orc_bcp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OrcResult), C.c_void_p]
    L.orc_batch_mt.restype = None
    L.orc_batch_mt.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint64,
                               C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_batch_enum.restype = None
    L.orc_batch_enum.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
    if kind == "port":
        L.orc_mt19937_draws.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_bool_rand.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_sample_assumptions.argtypes = [C.c_uint32, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_enum_assumptions.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
        L.orc_predict_mean.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double),
                                       C.POINTER(C.c_longdouble), C.POINTER(C.c_longdouble)]
        L.orc_predict_prep.restype = C.c_longdouble
        L.orc_predict_prep.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
        L.orc_sample_variance.restype = C.c_double
        L.orc_sample_variance.argtypes = [C.c_void_p, C.c_uint64]
    else:
        L.orc_solve_assumps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(OrcResult),
                                        C.POINTER(C.c_int), C.c_void_p]
        L.orc_rc2.restype = C.c_int64
        L.orc_rc2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64,
                              C.POINTER(C.c_uint64), C.c_void_p]
    _libs[kind] = L
    return L


class Oracle:
    def __init__(self, n_vars: int, offsets: np.ndarray, lits: np.ndarray, kind: str = "port"):
        self.kind = kind
        self.L = _lib(kind)
        self.n_vars = int(n_vars)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        lits = np.ascontiguousarray(lits, dtype=np.int32)
        self.h = self.L.orc_create(self.n_vars, len(offsets) - 1, offsets.ctypes.data, lits.ctypes.data)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ok(self) -> bool:
        return bool(self.L.orc_ok(self.h))

    def add_units(self, lits: Sequence[int]) -> int:
        a = np.ascontiguousarray(lits, dtype=np.int32)
        return self.L.orc_add_units(self.h, a.ctypes.data, len(a))

    def predict_sample(self, assumps: Sequence[int], eval_type: int = 0, want_assigns: bool = True):
        a = np.ascontiguousarray(assumps, dtype=np.int32)
        r = OrcResult()
        asg = np.empty(self.n_vars, dtype=np.uint8) if want_assigns else None
        self.L.orc_predict_sample(self.h, a.ctypes.data, len(a), eval_type, C.byref(r),
                                  asg.ctypes.data if want_assigns else None)
        return r, asg

    def bcp(self, assumps: Sequence[int], want_assigns: bool = True):
        a = np.ascontiguousarray(assumps, dtype=np.int32)
        r = OrcResult()
        asg = np.empty(self.n_vars, dtype=np.uint8) if want_assigns else None
        self.L.orc_bcp(self.h, a.ctypes.data, len(a), C.byref(r), asg.ctypes.data if want_assigns else None)
        return r, asg

    def batch_mt(self, mode: int, seed: int, set_vars: Sequence[int], n: int, first_sample: int = 0,
                 first_draw: int = 0, extra: Sequence[int] = (), want_trail: bool = True):
        """Samples of the 1-worker mt19937 order, looped in C.  mode 0 = BCP only on the persistent
        DB, 1 = fresh solver per sample (solverProgramCalling, eval prep; `extra` literals assumed
        in every sample).  Returns (conflict[n] bool, full[n] bool, trail[n] uint32 | None,
        sums = [n_conflict, n_full, sum_trail, sumsq_trail, propagations, watch_scans])."""
        sv = np.ascontiguousarray(set_vars, dtype=np.int32)
        ex = np.ascontiguousarray(extra, dtype=np.int32)
        g = (n + 63) // 64
        cb = np.zeros(g, dtype=np.uint64)
        fb = np.zeros(g, dtype=np.uint64)
        tr = np.zeros(n, dtype=np.uint32) if want_trail else None
        sums = np.zeros(6, dtype=np.uint64)
        self.L.orc_batch_mt(self.h, mode, seed, sv.ctypes.data, len(sv), ex.ctypes.data if len(ex) else None, len(ex),
                            first_draw, first_sample, n, cb.ctypes.data, fb.ctypes.data,
                            tr.ctypes.data if want_trail else None, sums.ctypes.data)
        conflict = np.unpackbits(cb.view(np.uint8), bitorder="little")[:n].astype(bool)
        full = np.unpackbits(fb.view(np.uint8), bitorder="little")[:n].astype(bool)
        return conflict, full, tr, [int(x) for x in sums]

    def batch_enum(self, vars_sorted: Sequence[int], first: int, n: int):
        """Assignment indices first..first+n-1 (MakeAssignsFromMasks order), BCP only, looped in C:
        (conflict[n] bool, full[n] bool, sums)."""
        sv = np.ascontiguousarray(vars_sorted, dtype=np.int32)
        g = (n + 63) // 64
        cb = np.zeros(g, dtype=np.uint64)
        fb = np.zeros(g, dtype=np.uint64)
        sums = np.zeros(6, dtype=np.uint64)
        self.L.orc_batch_enum(self.h, sv.ctypes.data, len(sv), first, n, cb.ctypes.data, fb.ctypes.data, sums.ctypes.data)
        conflict = np.unpackbits(cb.view(np.uint8), bitorder="little")[:n].astype(bool)
        full = np.unpackbits(fb.view(np.uint8), bitorder="little")[:n].astype(bool)
        return conflict, full, [int(x) for x in sums]

    # ---- reference-only entry points
    def solve_assumps(self, assumps: Sequence[int], max_nof_restarts: int = 0):
        assert self.kind == "ref"
        a = np.ascontiguousarray(assumps, dtype=np.int32)
        r = OrcResult()
        st = C.c_int(0)
        model = np.empty(self.n_vars, dtype=np.uint8)
        self.L.orc_solve_assumps(self.h, a.ctypes.data, len(a), max_nof_restarts, C.byref(r), C.byref(st),
                                 model.ctypes.data)
        return r, st.value, model

    def rc2(self, d_set: Sequence[int], start: Sequence[int], diapason_size: int, cap: int):
        assert self.kind == "ref"
        ds = np.ascontiguousarray(d_set, dtype=np.int32)
        st = np.ascontiguousarray(start, dtype=np.int32)
        out = np.zeros((max(cap, 1), len(ds)), dtype=np.int32)
        valid = C.c_uint64(0)
        n = self.L.orc_rc2(self.h, ds.ctypes.data, st.ctypes.data, len(ds), diapason_size, cap, C.byref(valid),
                           out.ctypes.data)
        return valid.value, out[:n]


# ---- stateless helpers of the port
def mt19937_draws(seed: int, first: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    _lib("port").orc_mt19937_draws(seed, first, n, out.ctypes.data)
    return out


def bool_rand(seed: int, first: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    _lib("port").orc_bool_rand(seed, first, n, out.ctypes.data)
    return out


def sample_assumptions(seed: int, set_vars: Sequence[int], first_sample: int, n_samples: int) -> np.ndarray:
    sv = np.ascontiguousarray(set_vars, dtype=np.int32)
    out = np.empty((n_samples, len(sv)), dtype=np.int32)
    _lib("port").orc_sample_assumptions(seed, sv.ctypes.data, len(sv), first_sample, n_samples, out.ctypes.data)
    return out


def enum_assumptions(vars_sorted: Sequence[int], lint: int) -> np.ndarray:
    sv = np.ascontiguousarray(vars_sorted, dtype=np.int32)
    out = np.empty(len(sv), dtype=np.int32)
    _lib("port").orc_enum_assumptions(sv.ctypes.data, len(sv), lint, out.ctypes.data)
    return out


def predict_mean(cost: np.ndarray, k: int, proc_count: int):
    c = np.ascontiguousarray(cost, dtype=np.float64)
    s = C.c_double()
    med = C.c_longdouble()
    pred = C.c_longdouble()
    _lib("port").orc_predict_mean(c.ctypes.data, len(c), k, proc_count, C.byref(s), C.byref(med), C.byref(pred))
    return s.value, med.value, pred.value


def predict_prep(solved: int, n: int, k: int):
    return _lib("port").orc_predict_prep(solved, n, k)


def sample_variance(cost: np.ndarray) -> float:
    c = np.ascontiguousarray(cost, dtype=np.float64)
    return _lib("port").orc_sample_variance(c.ctypes.data, len(c))
