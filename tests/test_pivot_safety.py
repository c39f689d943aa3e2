"""Numerical quality of the engine's static pivot sequences (host logic, no GPU): on the
oracle's matrices of every test deck -- the DC start and the first transient step -- the LU
under the engine's order(s) must show no element growth worth the name and every pivot must
pass the reference's relative threshold (pivrel 1e-3, spfactor.cc RelThreshold / KLU tol).
The failure this guards against: pivoting an inductor's +-1 branch coupling in a transient,
where its (br,br) entry is -L*ag0 -- multipliers of that size multiply up along an inductor
chain (4e12 over the 12 stages of `jtl12`, which then died with E_TIMESTEP)."""
import numpy as np
import pytest

from tests.decks import variants
from tools.pivot_growth import engine_order, growth, MODETRAN, MODETRANOP, INITJCT, INITTRAN, UIC
from oracle.pyoracle import Oracle


SINGULAR = {"float_cap"}        # singular on purpose: no order factors it


@pytest.mark.parametrize("name", sorted(set(variants()) - SINGULAR))
def test_static_orders_are_growth_free(name):
    d = variants()[name]
    n, Ap, Ai, rp, cp, rp2, cp2, npr = engine_order(d)
    o = Oracle(d)
    ts, tp, t0, tm, uic = d.tran_args
    if not uic:
        err, Ax, _ = o.debug_load(MODETRANOP | INITJCT, 0.0, 0.0, 1)
        g, w = growth(n, Ap, Ai, Ax, rp, cp)
        # (a pivot below the relative threshold is harmless where nothing grows: in the phase-mode
        # DC matrix a junction's diagonal Ic*cos(phi) ~ 2.5e-4 sits in a column of +-1 couplings,
        # and what its multipliers add to the branch rows stays below their 2*pi*L/Phi0 ~ 1e4)
        assert g <= 1e3 and (w >= 1e-3 or g <= 2.0), ("dc", g, w)
    delta = min(ts, (tp - t0)/50*1.001)*0.1
    err, Ax, _ = o.debug_load(MODETRAN | INITTRAN | (UIC if uic else 0), delta, delta, 1)
    g, w = growth(n, Ap, Ai, Ax, rp2, cp2)
    assert g <= 10.0 and w >= 1e-3, ("tran", g, w, npr)


def test_inductor_decks_keep_two_sequences():
    """DC needs the +-1 couplings as pivots (the branch diagonal is 0 there); a transient must
    not use them: such decks get two programs, MOS / RC decks and `uic` decks one."""
    two = {k for k in variants() if engine_order(variants()[k])[7] == 2}
    assert {"lc_ladder", "rlc_series", "rl_isrc", "inv_lsupply"} <= two
    assert not ({"ro5", "inv4", "rc_sin", "jtl12", "rlc_uic"} & two)
