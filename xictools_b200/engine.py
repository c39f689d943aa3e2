"""ctypes binding of the C ABI (include/b200spice.h) -- the host-side mirror of the calls a
WRspice maintainer would make from sCHECKprms::loop / sDCTprms::loop_mt.

There is no CPU path: without the CUDA library (xictools_b200/libb200spice.so, built by
`__graft_entry__.build()`), or without a CUDA device, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .deck import Deck, NCOUNTERS

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200spice.so")

E_NOGPU = 1001


class B200Error(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"{what}: error {code}")
        self.code = code


class RunStats(C.Structure):
    _fields_ = [("launches", C.c_longlong), ("ticks", C.c_longlong),
                ("load_instances", C.c_longlong),
                ("ms_total", C.c_double), ("ms_load", C.c_double), ("ms_solve", C.c_double),
                ("ms_control", C.c_double),
                ("solve_launches", C.c_longlong), ("load_launches", C.c_longlong),
                ("algbytes_load_per_launch", C.c_double),
                ("algbytes_solve_per_launch", C.c_double)]


_lib = None


def load_library(path: str | None = None):
    """Load libb200spice.so (fails loudly if it was not built)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise B200Error(E_NOGPU, f"CUDA extension {p} is missing (run __graft_entry__.build())")
    L = C.CDLL(p)
    L.b200_version.restype = C.c_char_p
    L.b200_batch_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.b200_batch_destroy.argtypes = [C.c_void_p]
    L.b200_batch_set_param.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.b200_batch_set_params.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200_batch_get_results_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200_batch_run_tran.argtypes = [C.c_void_p]
    L.b200_batch_run_op.argtypes = [C.c_void_p]
    L.b200_batch_num_unknowns.argtypes = [C.c_void_p]
    L.b200_batch_num_points.argtypes = [C.c_void_p]
    L.b200_batch_set_saves.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.b200_batch_num_saved.argtypes = [C.c_void_p]
    L.b200_batch_unknown_names.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.b200_batch_get_counters.argtypes = [C.c_void_p, C.c_void_p]
    L.b200_batch_get_waves.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200_batch_get_waves_sel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.b200_batch_get_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.b200_batch_pattern.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.b200_batch_ordering.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200_batch_ordering_tran.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200_batch_debug_load.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200_klu_analyze.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_double, C.c_int, C.POINTER(C.c_void_p)]
    L.b200_klu_free.argtypes = [C.c_void_p]
    L.b200_klu_refactor_solve_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p]
    L.b200_klu_lu_nnz.argtypes = [C.c_void_p]
    if path is None:
        _lib = L
    return L


class Batch:
    """One deck x ntrials trials on one GPU."""

    def __init__(self, deck: Deck, ntrials: int, device: int = 0, lib=None):
        self.L = lib or load_library()
        self.deck = deck
        self.ntrials = ntrials
        cd, keep = deck.to_c()
        self._keep = (cd, keep)
        h = C.c_void_p()
        rc = self.L.b200_batch_create(C.byref(cd), ntrials, device, C.byref(h))
        if rc:
            raise B200Error(rc, "b200_batch_create")
        self.h = h
        self.n = self.L.b200_batch_num_unknowns(h)
        self.npoints = self.L.b200_batch_num_points(h)

    def close(self):
        if getattr(self, "h", None):
            self.L.b200_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_param(self, param: int, index: int, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape == (self.ntrials,)
        rc = self.L.b200_batch_set_param(self.h, param, index, v.ctypes.data)
        if rc:
            raise B200Error(rc, "b200_batch_set_param")

    def set_params(self, overrides: dict):
        for (pid, idx), vals in overrides.items():
            self.set_param(pid, idx, vals)

    def set_params_packed(self, params, indices, values):
        """All per-trial rows in one call: values[nrows][ntrials] (C-contiguous float64; a
        page-locked buffer is copied by DMA) with row r = parameter params[r] of instance
        indices[r].  `values` must stay alive and unchanged until the next run returns."""
        p = np.ascontiguousarray(params, dtype=np.int32)
        i = np.ascontiguousarray(indices, dtype=np.int32)
        v = values
        if not (isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags.c_contiguous):
            v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape == (len(p), self.ntrials)
        self._packed_keep = (p, i, v)
        rc = self.L.b200_batch_set_params(self.h, len(p), p.ctypes.data, i.ctypes.data, v.ctypes.data)
        if rc:
            raise B200Error(rc, "b200_batch_set_params")

    def results_to_device(self, sel, dwaves_ptr: int, dcounters_ptr: int = 0):
        """Counters and selected waveforms into device buffers of the caller (raw pointers)."""
        s = np.ascontiguousarray(sel, dtype=np.int32)
        rc = self.L.b200_batch_get_results_dev(self.h, len(s), s.ctypes.data, dwaves_ptr, dcounters_ptr)
        if rc:
            raise B200Error(rc, "b200_batch_get_results_dev")

    def run_tran(self):
        rc = self.L.b200_batch_run_tran(self.h)
        if rc:
            raise B200Error(rc, "b200_batch_run_tran")

    def run_op(self):
        """DC operating point of every trial (.op); waves() then holds one point per trial."""
        rc = self.L.b200_batch_run_op(self.h)
        if rc:
            raise B200Error(rc, "b200_batch_run_op")

    def names(self):
        buf = C.create_string_buffer(1 << 22)
        nb = self.L.b200_batch_unknown_names(self.h, buf, len(buf))
        generic = [s.decode() for s in buf.raw[:nb].split(b"\0")[:-1]]
        return self.deck.unknown_names(generic)

    def counters(self, out=None) -> np.ndarray:
        c = out if out is not None else np.empty((self.ntrials, NCOUNTERS), dtype=np.int32)
        rc = self.L.b200_batch_get_counters(self.h, c.ctypes.data)
        if rc:
            raise B200Error(rc, "b200_batch_get_counters")
        return c

    def times(self) -> np.ndarray:
        """The user time points (t_check sequence, tranan.cc:472-479)."""
        t = np.zeros(self.npoints)
        rc = self.L.b200_batch_get_waves(self.h, t.ctypes.data, None)
        if rc:
            raise B200Error(rc, "b200_batch_get_waves")
        return t

    def set_saves(self, unknowns):
        """`.save`: keep only these unknowns (indices into names()) at the output points; before
        the first run.  waves() then returns these columns, in this order."""
        s = np.ascontiguousarray(unknowns, dtype=np.int32)
        rc = self.L.b200_batch_set_saves(self.h, len(s), s.ctypes.data)
        if rc:
            raise B200Error(rc, "b200_batch_set_saves")

    def waves(self, sel=None, out=None):
        """(times[npoints], waves[ntrials][npoints][saved or len(sel)]); `out` = a caller-owned
        result buffer of that shape (reused across runs, may be page-locked)."""
        t = np.zeros(self.npoints)
        if sel is None:
            w = out if out is not None else np.empty((self.ntrials, self.npoints, self.L.b200_batch_num_saved(self.h)))
            rc = self.L.b200_batch_get_waves(self.h, t.ctypes.data, w.ctypes.data)
        else:
            s = np.ascontiguousarray(sel, dtype=np.int32)
            w = out if out is not None else np.empty((self.ntrials, self.npoints, len(s)))
            rc = self.L.b200_batch_get_waves(self.h, t.ctypes.data, None)
            rc = rc or self.L.b200_batch_get_waves_sel(self.h, len(s), s.ctypes.data, w.ctypes.data)
        if rc:
            raise B200Error(rc, "b200_batch_get_waves")
        return t, w

    def stats(self) -> RunStats:
        st = RunStats()
        self.L.b200_batch_get_stats(self.h, C.byref(st))
        return st

    def pattern(self):
        n, nnz = C.c_int(), C.c_int()
        ap, ai = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
        self.L.b200_batch_pattern(self.h, C.byref(n), C.byref(nnz), C.byref(ap), C.byref(ai))
        return (np.array(ap[:n.value + 1], dtype=np.int32),
                np.array(ai[:nnz.value], dtype=np.int32))

    def ordering(self):
        r, c = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
        self.L.b200_batch_ordering(self.h, C.byref(r), C.byref(c))
        return np.array(r[:self.n], dtype=np.int32), np.array(c[:self.n], dtype=np.int32)

    def ordering_tran(self):
        """The pivot order of the transient modes (= ordering() for a single-sequence batch)."""
        r, c = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
        self.L.b200_batch_ordering_tran(self.h, C.byref(r), C.byref(c))
        return np.array(r[:self.n], dtype=np.int32), np.array(c[:self.n], dtype=np.int32)

    def debug_load(self, mode, time, delta, order, rhsold=None):
        """One sCKT::load + loadGmin at a forced state for all trials.
        rhsold: [n+1][ntrials] or None.  Returns (Ax[nnz][ntrials], rhs[n+1][ntrials])."""
        Ap, Ai = self.pattern()
        nnz = len(Ai)
        Ax = np.zeros((nnz, self.ntrials))
        rhs = np.zeros((self.n + 1, self.ntrials))
        ro = None
        if rhsold is not None:
            ro = np.ascontiguousarray(rhsold, dtype=np.float64)
            assert ro.shape == (self.n + 1, self.ntrials)
        rc = self.L.b200_batch_debug_load(self.h, mode, time, delta, order,
                                          ro.ctypes.data if ro is not None else None,
                                          Ax.ctypes.data, rhs.ctypes.data)
        if rc:
            raise B200Error(rc, "b200_batch_debug_load")
        return Ax, rhs


class BatchKLU:
    """KLU-hook seam: analyze once, then batched refactor+solve (host buffers)."""

    def __init__(self, n, Ap, Ai, Ax0, pivrel=1e-3, pivabs=1e-13, device=0, lib=None):
        self.L = lib or load_library()
        self.n = n
        self.Ap = np.ascontiguousarray(Ap, dtype=np.int32)
        self.Ai = np.ascontiguousarray(Ai, dtype=np.int32)
        a0 = np.ascontiguousarray(Ax0, dtype=np.float64)
        h = C.c_void_p()
        rc = self.L.b200_klu_analyze(n, self.Ap.ctypes.data, self.Ai.ctypes.data, a0.ctypes.data,
                                     pivrel, pivabs, device, C.byref(h))
        if rc:
            raise B200Error(rc, "b200_klu_analyze")
        self.h = h

    def refactor_solve(self, Ax, B):
        """Ax[ntrials][nnz], B[ntrials][n] -> (X[ntrials][n], status[ntrials])"""
        Ax = np.ascontiguousarray(Ax, dtype=np.float64)
        X = np.array(B, dtype=np.float64, copy=True, order="C")
        nt = Ax.shape[0]
        st = np.zeros(nt, dtype=np.int32)
        rc = self.L.b200_klu_refactor_solve_host(self.h, nt, Ax.ctypes.data, X.ctypes.data,
                                                 st.ctypes.data)
        if rc:
            raise B200Error(rc, "b200_klu_refactor_solve_host")
        return X, st

    def close(self):
        if getattr(self, "h", None):
            self.L.b200_klu_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
