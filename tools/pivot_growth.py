"""Diagnostic (CPU): element growth of the LU under the engine's static pivot order, on the
oracle's matrices of a deck in the DC start (MODEINITJCT) and the first transient step."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.pyoracle import Oracle
from xictools_b200 import engine
from tests.decks import variants

MODETRAN, MODETRANOP, INITJCT, INITTRAN, INITFLOAT, UIC = 0x1, 0x20, 0x200, 0x1000, 0x100, 0x10000


def engine_order(d):
    L = engine.load_library()
    cd, keep = d.to_c()
    st = (C.c_longlong*32)()
    L.b200_deck_analysis_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    rc = L.b200_deck_analysis_stats(C.byref(cd), st, 16)
    if rc:
        return None
    n, nnz = st[0], st[1]
    Ap = (C.c_int*(n+1))(); Ai = (C.c_int*nnz)(); rp = (C.c_int*n)(); cp = (C.c_int*n)()
    L.b200_deck_analysis_arrays.argtypes = [C.c_void_p] + [C.c_void_p]*4
    L.b200_deck_analysis_arrays(C.byref(cd), Ap, Ai, rp, cp)
    rp2 = (C.c_int*n)(); cp2 = (C.c_int*n)(); npr = C.c_int()
    L.b200_deck_analysis_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200_deck_analysis_order(C.byref(cd), 0, rp, cp, C.byref(npr))
    L.b200_deck_analysis_order(C.byref(cd), 1, rp2, cp2, C.byref(npr))
    return n, np.array(Ap), np.array(Ai), np.array(rp), np.array(cp), np.array(rp2), np.array(cp2), npr.value


def growth(n, Ap, Ai, Ax, rp, cp):
    A = np.zeros((n, n))
    for j in range(n):
        for p in range(Ap[j], Ap[j+1]):
            A[Ai[p], j] = Ax[p]
    B = A[np.ix_(rp, cp)].copy()
    amax = np.abs(B).max()
    g = 1.0
    worst = np.inf
    for k in range(n):
        piv = B[k, k]
        colmax = np.abs(B[k:, k]).max()
        if piv == 0:
            return np.inf, 0.0
        if np.count_nonzero(B[k, k+1:]):      # a pivot alone in its row propagates nothing: any size is safe
            worst = min(worst, abs(piv)/colmax if colmax else 1.0)
        B[k+1:, k] /= piv
        B[k+1:, k+1:] -= np.outer(B[k+1:, k], B[k, k+1:])
        g = max(g, np.abs(B[k+1:, k+1:]).max()/amax if k + 1 < n else g)
    return g, worst


def main():
  names = sys.argv[1:] or sorted(variants())
  for name in names:
      d = variants()[name]
      eo = engine_order(d)
      if eo is None:
          print(name, "analysis refused"); continue
      n, Ap, Ai, rp, cp, rp2, cp2, npr = eo
      o = Oracle(d)
      ts, tp, t0, tm, uic = d.tran_args
      out = []
      if not uic:
          err, Ax, rhs = o.debug_load(MODETRANOP | INITJCT, 0.0, 0.0, 1)
          Ax = Ax.copy()
          # loadGmin during the DC start adds diagGmin; take the plain matrix + gmin on node diagonals
          out.append(("dc", growth(n, Ap, Ai, Ax + 0.0, rp, cp)))
      delta = min(ts, (tp - t0)/50*1.001)*0.1
      err, Ax, rhs = o.debug_load(MODETRAN | INITTRAN | (UIC if uic else 0), delta, delta, 1)
      out.append(("tran", growth(n, Ap, Ai, Ax, rp2, cp2)))
      print(f"{name:22s} n={n:4d} programs {npr} " + "  ".join(f"{k}: growth {g:9.3g} min|piv|/colmax {w:9.3g}" for k, (g, w) in out))


if __name__ == '__main__':
  main()
