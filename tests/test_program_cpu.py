"""Host logic of the solve (no GPU): the static pivot order, the symbolic LU, the assembly runs
and the level-scheduled task program that k_assemble / k_lu execute are compiled on the host.
b200_debug_program_check runs that compiled program on one value set in kernel order; the
result must equal the oracle's static LU (same pivot order, same split rule for long sums)
BIT FOR BIT, and solve the system."""
import ctypes as C

import numpy as np
import pytest

from oracle.pyoracle import Oracle, lu_solve_static
from tests.decks import variants
from xictools_b200 import engine
from xictools_b200.deck import ring_oscillator

MODEDCOP, MODEINITJCT = 0x10, 0x200
MODETRAN, MODEINITTRAN, MODEINITFLOAT = 0x1, 0x1000, 0x100


def program_check(n, Ap, Ai, Ax, b):
    L = engine.load_library()
    L.b200_debug_program_check.argtypes = [C.c_int] + [C.c_void_p]*8 + [C.c_int]
    Ap = np.ascontiguousarray(Ap, np.int32); Ai = np.ascontiguousarray(Ai, np.int32)
    Ax = np.ascontiguousarray(Ax, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(n); rp = np.zeros(n, np.int32); cp = np.zeros(n, np.int32)
    st = np.zeros(8, np.int64)
    rc = L.b200_debug_program_check(n, Ap.ctypes.data, Ai.ctypes.data, Ax.ctypes.data, b.ctypes.data,
                                    x.ctypes.data, rp.ctypes.data, cp.ctypes.data, st.ctypes.data, 8)
    keys = ("nnzLU", "img", "scratch", "tasks", "ops", "levels", "rounds", "steps")
    return rc, x, rp, cp, dict(zip(keys, st.tolist()))


def dense(n, Ap, Ai, Ax):
    A = np.zeros((n, n))
    for j in range(n):
        for p in range(Ap[j], Ap[j+1]):
            A[Ai[p], j] += Ax[p]
    return A


def check_system(n, Ap, Ai, Ax, b, rtol=1e-9):
    rc, x, rp, cp, st = program_check(n, Ap, Ai, Ax, b)
    assert rc == 0
    err, xo = lu_solve_static(n, Ap, Ai, Ax, rp, cp, b)
    assert err == 0
    assert np.array_equal(x.view(np.int64), xo.view(np.int64)), np.abs(x - xo).max()
    A = dense(n, Ap, Ai, Ax)
    xr = np.linalg.solve(A, b)
    assert np.allclose(x, xr, rtol=rtol, atol=rtol*np.abs(xr).max())
    return st


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("name", ["inv4", "ro5", "lc_ladder", "rlc_series", "inv_lsupply"])
def test_program_matches_oracle_static_lu_on_deck_matrices(name, wide, monkeypatch):
    """`wide`: the 32-bit-slot encoding of the same program (images beyond one warp's
    shared-memory window, k_lu_wide) -- forced here on small systems, bit-identical results."""
    if wide:
        monkeypatch.setenv("B200_LU_WIDE", "1")
    d = variants()[name]
    o = Oracle(d)
    Ap, Ai = o.pattern()
    rng = np.random.default_rng(3)
    x0 = np.concatenate([[0.0], rng.uniform(0.0, 3.0, o.n)])
    for mode, delta, order in ((MODEDCOP | MODEINITJCT, 0.0, 1), (MODETRAN | MODEINITFLOAT, 1e-11, 2)):
        err, Ax, rhs = o.debug_load(mode, 1e-10, delta or 1e-11, order, x0)
        assert err == 0
        check_system(o.n, Ap, Ai, Ax, rhs[1:])


def test_split_rule_constants_agree():
    L = engine.load_library()
    cd, keep = ring_oscillator(3).to_c()
    out = (C.c_longlong*32)()
    assert L.b200_deck_analysis_stats(C.byref(cd), out, 32) == 0
    lo = C.CDLL(Oracle._lib._name) if Oracle._lib else None
    if lo is None:
        Oracle(ring_oscillator(3))
        lo = Oracle._lib
    s, c = C.c_int(), C.c_int()
    lo.wo_lu_split_rule(C.byref(s), C.byref(c))
    # B200_AS_LU_SPLIT / B200_AS_LU_CHUNK follow B200_AS_LONGEST_RUN (include/b200spice.h)
    assert (out[14], out[15]) == (s.value, c.value)


def test_long_sums_are_split_and_still_exact():
    """An arrowhead matrix: the last row / column couples to every unknown, so its pivot, its
    forward-solve row and its backward-solve row are sums of n-1 products -- far beyond the
    split threshold, twice recursively at n = 700."""
    for n in (40, 700):
        rng = np.random.default_rng(n)
        cols = []
        for j in range(n - 1):
            cols.append([(j, 4.0 + rng.uniform()), (n - 1, rng.uniform(-1, 1))])
        cols.append([(i, rng.uniform(-1, 1)) for i in range(n - 1)] + [(n - 1, 10.0 + n)])
        Ap = [0]; Ai = []; Ax = []
        for c in cols:
            c.sort()
            Ai += [i for i, _ in c]; Ax += [v for _, v in c]; Ap.append(len(Ai))
        b = rng.uniform(-1, 1, n)
        st = check_system(n, Ap, Ai, Ax, b)
        if n == 700:
            assert st["scratch"] > 40          # partial sums exist
            assert st["steps"] < 400           # and no lane carries a 699-product chain


def test_random_sparse_systems():
    rng = np.random.default_rng(11)
    for n in (5, 33, 120):
        A = np.zeros((n, n))
        for i in range(n):
            A[i, i] = 5.0 + rng.uniform()
            for j in rng.choice(n, size=min(n - 1, 3), replace=False):
                if j != i:
                    A[i, j] = rng.uniform(-1, 1)
        Ap = [0]; Ai = []; Ax = []
        for j in range(n):
            nz = np.nonzero(A[:, j])[0]
            Ai += nz.tolist(); Ax += A[nz, j].tolist(); Ap.append(len(Ai))
        check_system(n, Ap, Ai, Ax, rng.uniform(-1, 1, n))


def test_ro501_program_shape():
    """The bench deck: what one warp of k_lu executes per trial."""
    L = engine.load_library()
    cd, keep = ring_oscillator(501).to_c()
    out = (C.c_longlong*32)()
    assert L.b200_deck_analysis_stats(C.byref(cd), out, 32) == 0
    n, nnzlu, img, levels, rounds, steps = out[0], out[2], out[5], out[9], out[10], out[11]
    assert n == 503
    assert img*8 <= 32*1024                  # one trial image: seven fit an SM's shared memory
    assert levels <= 40 and steps <= 500     # nested-dissection-like order: shallow and short


def test_large_line_compiles_to_a_wide_program():
    """A 3000-junction line (6001 unknowns): the image (36 k doubles) exceeds the 227 KB window,
    the analysis -- pivot order, symbolic LU, level schedule -- stays well under a second and
    emits the wide encoding with the same shallow schedule as the small lines."""
    import time
    from xictools_b200.workloads import jtl
    L = engine.load_library()
    cd, keep = jtl(3000).to_c()
    out = (C.c_longlong*32)()
    t0 = time.time()
    assert L.b200_deck_analysis_stats(C.byref(cd), out, 32) == 0
    assert time.time() - t0 < 5.0
    n, img, levels = out[0], out[5], out[9]
    assert n == 6001 and img*8 > 229000 and levels < 64
    assert out[16] == 1                     # B200_AS_WIDE
