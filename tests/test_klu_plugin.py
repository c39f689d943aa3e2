"""The reference's real sparse-solver seam: WRspice's KLU plug-in loader (KLUif::find_klu,
wrspice/src/sparse/kluif.cc:117-266) must accept libb200spice.so through $XT_KLU_PATH, and the
reference's own KLUmatrix (the spMatlabMatrix subclass, kluif.cc:268-681) must then run a whole
analysis through klu_analyze / klu_factor / klu_refactor / klu_solve served by the B200.

oracle/_ref/libwrsref_klu.so is the UNMODIFIED reference compiled with -DUSE_KLU (see
oracle/refbuild/Makefile); the CPU tests here check loading and the no-fallback rule, the GPU
tests check the numbers against the same reference running its own Sparse 1.3 solver."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle.pyoracle import RefSim, have_ref, have_ref_klu, run_ref_klu
from tests.decks import variants
from xictools_b200 import engine
from xictools_b200.deck import ring_oscillator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RELTOL, VNTOL = 1e-3, 1e-6

# wrspice/include/kluif.h:196-226 (the 18 pointers KLUif resolves) + the version probe
KLU_SYMBOLS = ["klu_defaults", "klu_analyze", "klu_factor", "klu_ld_factor", "klu_z_factor",
               "klu_refactor", "klu_ld_refactor", "klu_z_refactor", "klu_solve", "klu_ld_solve",
               "klu_z_solve", "klu_tsolve", "klu_ld_tsolve", "klu_z_tsolve", "klu_free_symbolic",
               "klu_free_numeric", "klu_ld_free_numeric", "klu_z_free_numeric",
               "SuiteSparse_version"]


class KluCommon(C.Structure):
    _fields_ = [("tol", C.c_double), ("memgrow", C.c_double), ("initmem_amd", C.c_double),
                ("initmem", C.c_double), ("maxwork", C.c_double), ("btf", C.c_int),
                ("ordering", C.c_int), ("scale", C.c_int), ("user_order", C.c_void_p),
                ("user_data", C.c_void_p), ("halt_if_singular", C.c_int), ("status", C.c_int),
                ("nrealloc", C.c_int), ("structural_rank", C.c_int), ("numerical_rank", C.c_int),
                ("singular_col", C.c_int), ("noffdiag", C.c_int), ("flops", C.c_double),
                ("rcond", C.c_double), ("condest", C.c_double), ("rgrowth", C.c_double),
                ("work", C.c_double), ("memusage", C.c_size_t), ("mempeak", C.c_size_t)]


class KluSymbolic(C.Structure):
    _fields_ = [("symmetry", C.c_double), ("est_flops", C.c_double), ("lnz", C.c_double),
                ("unz", C.c_double), ("Lnz", C.c_void_p), ("n", C.c_int), ("nz", C.c_int),
                ("P", C.POINTER(C.c_int)), ("Q", C.POINTER(C.c_int)), ("R", C.c_void_p),
                ("nzoff", C.c_int), ("nblocks", C.c_int), ("maxblock", C.c_int),
                ("ordering", C.c_int), ("do_btf", C.c_int), ("structural_rank", C.c_int)]


def test_every_symbol_the_loader_resolves_is_exported_and_declared():
    lib = engine.load_library()
    assert not [s for s in KLU_SYMBOLS if not hasattr(lib, s)]
    hdr = open(os.path.join(ROOT, "include", "b200_klu_abi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    assert not [s for s in KLU_SYMBOLS if not re.search(r"\b%s\s*\(" % s, hdr)]


def test_version_is_the_one_the_reference_was_compiled_for():
    lib = engine.load_library()
    v = (C.c_int*3)()
    lib.SuiteSparse_version(v)
    assert list(v) == [4, 4, 6]              # KLU/Makefile.in:11, kluif.cc:157-166


def test_defaults_analyze_free_without_a_gpu():
    """klu_defaults / klu_analyze / klu_free_symbolic are host-only: pattern validation and the
    status / singular_col / P / Q fields the adapter reads."""
    lib = engine.load_library()
    lib.klu_analyze.restype = C.POINTER(KluSymbolic)
    c = KluCommon()
    assert lib.klu_defaults(C.byref(c)) == 1
    assert (c.tol, c.btf, c.ordering, c.scale, c.halt_if_singular, c.status) == (1e-3, 1, 0, 2, 1, 0)
    Ap = (C.c_int*4)(0, 2, 4, 5)
    Ai = (C.c_int*5)(0, 1, 0, 1, 2)
    s = lib.klu_analyze(3, Ap, Ai, C.byref(c))
    assert s and c.status == 0 and s.contents.n == 3 and s.contents.nz == 5
    assert [s.contents.P[i] for i in range(3)] == [0, 1, 2]
    sp = C.POINTER(KluSymbolic)(s.contents)
    lib.klu_free_symbolic(C.byref(sp), C.byref(c))
    assert not sp
    bad = (C.c_int*5)(1, 0, 0, 1, 2)         # rows not ascending in column 0
    assert not lib.klu_analyze(3, Ap, bad, C.byref(c)) and c.status == -3


@pytest.mark.skipif(not have_ref_klu(), reason="oracle/_ref/libwrsref_klu.so not built")
def test_the_reference_loader_accepts_the_library():
    assert run_ref_klu(None, engine.LIB_PATH)["klu_ok"] == 1
    assert run_ref_klu(None, None)["klu_ok"] == 0
    assert run_ref_klu(None, os.path.join(ROOT, "oracle", "libwrsoracle.so"))["klu_ok"] == 0


@pytest.mark.skipif(not have_ref_klu(), reason="oracle/_ref/libwrsref_klu.so not built")
def test_plugged_in_without_a_gpu_fails_loudly():
    """No CPU fallback behind the plug-in: the reference gets an error from factor(), and its
    own Sparse path (no plug-in) still runs the deck."""
    lib = engine.load_library()
    if lib.b200_cuda_device_count() > 0:
        pytest.skip("GPU present")
    txt = ring_oscillator(5).to_spice()
    r = run_ref_klu(txt, engine.LIB_PATH)
    assert r["klu_ok"] == 1 and r["err"] != 0 and r["accepted"] == 0
    r = run_ref_klu(txt, None)
    assert r["err"] == 0 and r["accepted"] > 0


@pytest.mark.gpu
@pytest.mark.skipif(not (have_ref_klu() and have_ref()), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["ro5", "inv4", "rlc_series", "lc_ladder", "transformer", "rc_sin",
                                  "ro5_level3", "ro5_gear3", "jtl6", "jtl6_dc", "jtl6_dc_phase", "jtl6_dc_accel"])
def test_reference_with_the_b200_plugin_matches_its_own_sparse_run(name):
    """parser -> sCKT::doTask -> NIiter -> KLUmatrix::factor/refactor/solve -> libb200spice.so
    (GPU) against the same reference with Sparse 1.3: identical step / iteration counts,
    waveforms within reltol/vntol at every output point."""
    d = variants()[name]
    txt = d.to_spice()
    r = run_ref_klu(txt, engine.LIB_PATH)
    assert r["klu_ok"] == 1, r["stderr"][-500:]
    s = RefSim().run(txt)
    assert r["err"] == s["err"] == 0, (r["err"], s["err"], r["stderr"][-500:])
    for k in ("accepted", "rejected", "numiter", "timepts"):
        assert r[k] == s[k], (k, r[k], s[k])
    assert r["names"] == s["names"]
    got, ref = r["vals"], s["vals"]
    assert got.shape == ref.shape
    tol = RELTOL*np.maximum(np.abs(got), np.abs(ref)) + VNTOL
    assert (np.abs(got - ref) <= tol).all(), float((np.abs(got - ref)/tol).max())
