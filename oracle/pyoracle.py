"""ctypes wrappers for the two CPU checkers.  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by
xictools_b200/.

  RefSim   oracle/_ref/libwrsref.so   the UNMODIFIED reference compiled headless
  Oracle   oracle/libwrsoracle.so     the plain-C restatement (wrs_oracle.c)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libwrsref.so")
ORACLE_SO = os.path.join(HERE, "libwrsoracle.so")


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class RefSim:
    """One process-global instance: the reference keeps global singletons."""
    _lib = None

    def __init__(self):
        if RefSim._lib is None:
            RefSim._lib = C.CDLL(REF_SO)
            L = RefSim._lib
            L.wrsref_names.restype = C.POINTER(C.c_char)
            L.wrsref_load_order.restype = C.POINTER(C.c_char)
        self.L = RefSim._lib

    def run(self, spice_text: str) -> dict:
        L = self.L
        err = L.wrsref_run_deck(spice_text.encode())
        cnt = (C.c_int * 9)()
        L.wrsref_get_counters(cnt)
        tm = (C.c_double * 7)()
        L.wrsref_get_times(tm)
        npts, nn = L.wrsref_num_points(), L.wrsref_num_names()
        t = np.zeros(npts)
        v = np.zeros((npts, max(nn, 1)))
        if npts:
            L.wrsref_get_points(t.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        nb = C.c_int()
        p = L.wrsref_names(C.byref(nb))
        names = [s.decode() for s in bytes(p[:nb.value]).split(b"\0")[:-1]]
        p = L.wrsref_load_order(C.byref(nb))
        order = [s.decode() for s in bytes(p[:nb.value]).split(b"\0")[:-1]]
        keys = ("error", "size", "accepted", "rejected", "numiter", "timepts", "traniter",
                "itercut", "trapcut")
        out = dict(zip(keys, list(cnt)))
        out.update(err=err, times=t, vals=v, names=names, load_order=order,
                   t_load=tm[0], t_decomp=tm[1], t_solve=tm[2], t_reorder=tm[3],
                   t_cvchk=tm[4], t_tran=tm[5], t_total=tm[6])
        return out


def run_harvest(spice_text: str, libpath: str, ntrials: int, params=None, op_only=False, saves=None) -> dict:
    """Reference parser + set-up circuit -> shim/harvest.cc -> C ABI -> batched run (GPU), in
    the same call that runs the reference's own doTask on that circuit.
    params: [(param_id, instance_name, values[ntrials])].  ntrials = 0: host-only harvest check
    (returns the engine's analysis sizes of the harvested deck)."""
    sim = RefSim()
    L = sim.L
    params = params or []
    n = len(params)
    ids = (C.c_int*max(n, 1))(*[p[0] for p in params])
    names = (C.c_char_p*max(n, 1))(*[p[1].encode() for p in params])
    arrs = [np.ascontiguousarray(p[2], dtype=np.float64) for p in params]
    vals = (C.c_void_p*max(n, 1))(*[a.ctypes.data for a in arrs])
    L.wrsref_run_deck_harvest.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int]
    sv = np.ascontiguousarray(saves if saves is not None else [], dtype=np.int32)
    L.wrsref_harvest_set_saves(len(sv), sv.ctypes.data_as(C.c_void_p))       # `.save` list (columns of `waves`)
    err = L.wrsref_run_deck_harvest(spice_text.encode(), libpath.encode(), ntrials, n, ids, names,
                                    vals, int(op_only))
    L.wrsref_harvest_set_saves(0, None)
    out = {"ref_err": err, "rc": L.wrsref_harvest_rc()}
    cnt = (C.c_int*9)()
    L.wrsref_get_counters(cnt)
    out["ref"] = dict(zip(("error", "size", "accepted", "rejected", "numiter", "timepts", "traniter",
                           "itercut", "trapcut"), list(cnt)))
    npts, nn = L.wrsref_num_points(), L.wrsref_num_names()
    t = np.zeros(npts)
    v = np.zeros((npts, max(nn, 1)))
    if npts:
        L.wrsref_get_points(t.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
    out["ref_times"], out["ref_vals"] = t, v
    if ntrials <= 0:
        st = (C.c_longlong*32)()
        L.wrsref_harvest_stats(st)
        out["stats"] = list(st)
        return out
    if out["rc"] == 0:
        N, P = L.wrsref_harvest_n(), L.wrsref_harvest_npoints()
        counters = np.zeros((ntrials, 9), dtype=np.int32)
        times = np.zeros(P)
        waves = np.zeros((ntrials, P, N))
        L.wrsref_harvest_get(counters.ctypes.data_as(C.c_void_p), times.ctypes.data_as(C.c_void_p),
                             waves.ctypes.data_as(C.c_void_p))
        out.update(counters=counters, times=times, waves=waves)
        ms, ticks = C.c_double(), C.c_longlong()
        L.wrsref_harvest_last_run(C.byref(ms), C.byref(ticks))
        out.update(ms_total=ms.value, ticks=ticks.value)
    return out


def run_harvest_oracle(spice_text: str, params=None) -> dict:
    """Reference parser + set-up circuit -> shim/harvest.cc -> the plain-C ORACLE, next to the
    reference's own doTask on that circuit (for decks whose device tables only the reference's
    setup can make: TJM coefficient sets).  params: [(param_id, instance_name, value)]."""
    sim = RefSim()
    L = sim.L
    params = params or []
    n = len(params)
    ids = (C.c_int*max(n, 1))(*[p[0] for p in params])
    names = (C.c_char_p*max(n, 1))(*[p[1].encode() for p in params])
    vals = (C.c_double*max(n, 1))(*[float(p[2]) for p in params])
    L.wrsref_run_deck_harvest_oracle.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    err = L.wrsref_run_deck_harvest_oracle(spice_text.encode(), ORACLE_SO.encode(), n, ids, names, vals)
    cnt = (C.c_int*9)()
    L.wrsref_get_counters(cnt)
    out = {"ref_err": err, "rc": L.wrsref_oracle_rc(),
           "ref": dict(zip(("error", "size", "accepted", "rejected", "numiter", "timepts", "traniter",
                            "itercut", "trapcut"), list(cnt)))}
    npts, nn = L.wrsref_num_points(), L.wrsref_num_names()
    t = np.zeros(npts)
    v = np.zeros((npts, max(nn, 1)))
    if npts:
        L.wrsref_get_points(t.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
    out["ref_times"], out["ref_vals"] = t, v
    if out["rc"] == 0:
        N, P = L.wrsref_oracle_n(), L.wrsref_oracle_npoints()
        oc = np.zeros(9, dtype=np.int32)
        ot = np.zeros(P)
        ov = np.zeros((P, N))
        L.wrsref_oracle_get(oc.ctypes.data_as(C.c_void_p), ot.ctypes.data_as(C.c_void_p),
                            ov.ctypes.data_as(C.c_void_p))
        out.update(counters=oc, times=ot, vals=ov)
    return out


class Oracle:
    _lib = None

    def __init__(self, deck, overrides: dict | None = None, trial: int = 0):
        if Oracle._lib is None:
            Oracle._lib = C.CDLL(ORACLE_SO)
            L = Oracle._lib
            L.wo_create.restype = C.c_void_p
            L.wo_create.argtypes = [C.c_void_p]
            for f in ("wo_destroy", "wo_run_tran", "wo_run_op", "wo_num_unknowns", "wo_nnz", "wo_num_points",
                      "wo_num_accepted"):
                getattr(L, f).argtypes = [C.c_void_p]
            L.wo_set_param.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
            L.wo_set_ordering.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.wo_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.wo_unknown_names.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            L.wo_get_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.wo_get_counters.argtypes = [C.c_void_p, C.c_void_p]
            L.wo_get_accepted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.wo_debug_load.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
            L.wo_lu_solve_static.argtypes = [C.c_int] + [C.c_void_p] * 6
        self.L = Oracle._lib
        self.deck = deck
        cd, keep = deck.to_c()
        self._keep = (cd, keep)
        self.h = self.L.wo_create(C.byref(cd))
        if overrides:
            for (pid, idx), vals in overrides.items():
                v = vals[trial] if hasattr(vals, "__len__") else vals
                self.L.wo_set_param(self.h, pid, idx, float(v))
        self.n = self.L.wo_num_unknowns(self.h)
        self.nnz = self.L.wo_nnz(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.wo_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def names(self):
        buf = C.create_string_buffer(1 << 20)
        nb = self.L.wo_unknown_names(self.h, buf, len(buf))
        generic = [s.decode() for s in buf.raw[:nb].split(b"\0")[:-1]]
        return self.deck.unknown_names(generic)

    def pattern(self):
        ap, ai = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
        self.L.wo_pattern(self.h, C.byref(ap), C.byref(ai))
        Ap = np.array(ap[:self.n + 1], dtype=np.int32)
        Ai = np.array(ai[:self.nnz], dtype=np.int32)
        return Ap, Ai

    def set_ordering_tran(self, rowperm, colperm):
        """The engine's pivot order of the transient modes (two-program batches)."""
        r = np.ascontiguousarray(rowperm, dtype=np.int32)
        c = np.ascontiguousarray(colperm, dtype=np.int32)
        self.L.wo_set_ordering_tran.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.wo_set_ordering_tran(self.h, r.ctypes.data, c.ctypes.data)

    def follow(self, batch):
        """Use the pivot order(s) of an engine batch."""
        self.set_ordering(*batch.ordering())
        self.set_ordering_tran(*batch.ordering_tran())
        return self

    def set_ordering(self, rowperm, colperm):
        r = np.ascontiguousarray(rowperm, dtype=np.int32)
        c = np.ascontiguousarray(colperm, dtype=np.int32)
        self.L.wo_set_ordering(self.h, r.ctypes.data, c.ctypes.data)

    def run(self, analysis: str = "tran") -> dict:
        err = self.L.wo_run_op(self.h) if analysis == "op" else self.L.wo_run_tran(self.h)
        cnt = np.zeros(10, dtype=np.int32)
        self.L.wo_get_counters(self.h, cnt.ctypes.data)
        npts = self.L.wo_num_points(self.h)
        t = np.zeros(npts)
        v = np.zeros((npts, self.n))
        self.L.wo_get_points(self.h, t.ctypes.data, v.ctypes.data)
        na = self.L.wo_num_accepted(self.h)
        at = np.zeros(na)
        av = np.zeros((na, self.n))
        ai = np.zeros(na, dtype=np.int32)
        self.L.wo_get_accepted(self.h, at.ctypes.data, av.ctypes.data, ai.ctypes.data)
        return dict(err=err, accepted=int(cnt[1]), rejected=int(cnt[2]), numiter=int(cnt[3]),
                    timepts=int(cnt[4]), traniter=int(cnt[5]), itercut=int(cnt[6]),
                    trapcut=int(cnt[7]), times=t, vals=v, names=self.names(),
                    acc_times=at, acc_vals=av, acc_iters=ai)

    def debug_load(self, mode, time, delta, order, rhsold=None):
        Ax = np.zeros(self.nnz)
        rhs = np.zeros(self.n + 1)
        ro = None
        if rhsold is not None:
            ro = np.ascontiguousarray(rhsold, dtype=np.float64)
        err = self.L.wo_debug_load(self.h, mode, time, delta, order,
                                   ro.ctypes.data if ro is not None else None,
                                   Ax.ctypes.data, rhs.ctypes.data)
        return err, Ax, rhs


def lu_solve_static(n, Ap, Ai, Ax, rowperm, colperm, b):
    Oracle._lib or C.CDLL(ORACLE_SO)
    L = Oracle._lib or C.CDLL(ORACLE_SO)
    L.wo_lu_solve_static.argtypes = [C.c_int] + [C.c_void_p] * 6
    Ap = np.ascontiguousarray(Ap, dtype=np.int32)
    Ai = np.ascontiguousarray(Ai, dtype=np.int32)
    Ax = np.ascontiguousarray(Ax, dtype=np.float64)
    r = np.ascontiguousarray(rowperm, dtype=np.int32)
    c = np.ascontiguousarray(colperm, dtype=np.int32)
    x = np.array(b, dtype=np.float64, copy=True)
    err = L.wo_lu_solve_static(n, Ap.ctypes.data, Ai.ctypes.data, Ax.ctypes.data,
                               r.ctypes.data, c.ctypes.data, x.ctypes.data)
    return err, x


# ---------------------------------------------------------------------------------------------
# The reference with its KLU plug-in hook compiled in (oracle/_ref/libwrsref_klu.so).  Run in a
# child process: the plug-in path comes from the environment ($XT_KLU_PATH, kluif.cc:121) and
# the reference keeps process-global state.
# ---------------------------------------------------------------------------------------------
REF_KLU_SO = os.path.join(HERE, "_ref", "libwrsref_klu.so")


def have_ref_klu() -> bool:
    return os.path.exists(REF_KLU_SO)


_KLU_WORKER = r"""
import ctypes as C, json, sys
import numpy as np
so, deck_path = sys.argv[1], sys.argv[2]
L = C.CDLL(so)
L.wrsref_names.restype = C.POINTER(C.c_char)
ok = L.wrsref_find_klu()
out = {"klu_ok": int(ok)}
if deck_path != "-":
    text = open(deck_path).read()
    err = L.wrsref_run_deck(text.encode())
    cnt = (C.c_int*9)(); L.wrsref_get_counters(cnt)
    npts, nn = L.wrsref_num_points(), L.wrsref_num_names()
    t = np.zeros(npts); v = np.zeros((npts, max(nn, 1)))
    if npts:
        L.wrsref_get_points(t.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
    nb = C.c_int(); p = L.wrsref_names(C.byref(nb))
    names = [s.decode() for s in bytes(p[:nb.value]).split(b"\0")[:-1]]
    keys = ("error", "size", "accepted", "rejected", "numiter", "timepts", "traniter", "itercut", "trapcut")
    out.update(dict(zip(keys, list(cnt))))
    out.update(err=err, times=t.tolist(), vals=v.tolist(), names=names)
sys.stdout.write("\n@@RESULT@@" + json.dumps(out) + "\n")
"""


def run_ref_klu(spice_text: str | None, klu_path: str | None, timeout: float = 600.0) -> dict:
    """Run one deck through the KLU-enabled reference build with $XT_KLU_PATH = klu_path
    (None: unset -> the reference falls back to Sparse).  spice_text None: only report whether
    the reference's loader accepted the plug-in."""
    import json
    import subprocess
    import sys
    import tempfile
    env = dict(os.environ)
    env.pop("XT_KLU_PATH", None)
    if klu_path:
        env["XT_KLU_PATH"] = klu_path
    with tempfile.TemporaryDirectory() as td:
        dp = "-"
        if spice_text is not None:
            dp = os.path.join(td, "deck.cir")
            open(dp, "w").write(spice_text)
        r = subprocess.run([sys.executable, "-c", _KLU_WORKER, REF_KLU_SO, dp], env=env,
                           capture_output=True, text=True, timeout=timeout)
    if "@@RESULT@@" not in r.stdout:
        raise RuntimeError(f"KLU worker failed (rc {r.returncode}): {r.stderr[-2000:]}")
    out = json.loads(r.stdout.split("@@RESULT@@")[1])
    if "vals" in out:
        out["times"] = np.array(out["times"])
        out["vals"] = np.array(out["vals"], dtype=float).reshape(len(out["times"]), max(len(out["names"]), 1)) \
            if len(out["times"]) else np.zeros((0, max(len(out["names"]), 1)))
    out["stdout"] = r.stdout.split("@@RESULT@@")[0]
    out["stderr"] = r.stderr
    return out
