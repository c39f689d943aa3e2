"""Josephson tunnel-junction model (JJ level 3, mmjco / TJM).  Its coefficient
sets come out of the reference's own setup (built-in `tjm1` / `tjm2` tables, tjmcoeffs.cc), so
TJM decks exist as SPICE text only and travel: reference parser -> set-up sCKT ->
shim/harvest.cc -> (a) the plain-C oracle, pinned here against the same circuit's doTask, and
(b) the CUDA engine through the C ABI on the GPU box."""
import numpy as np
import pytest

from oracle.pyoracle import RefSim, have_ref, run_harvest, run_harvest_oracle
from xictools_b200 import engine
from xictools_b200.deck import (C_ACCEPTED, C_NPOINTS, C_NUMITER, C_REJECTED, C_STATUS, C_TIMEPTS,
                                DEV_JJ, P_TJM_ICS)
from xictools_b200.workloads import jj_single, jtl

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
RELTOL, VNTOL = 1e-3, 1e-6
RSJ = "jj(rtype=1 cct=1 vg=0.0028 delv=0.0001 icrit=0.001 rn=1.7 rsub=17.0 cap=7e-13)"
TJM1 = "jj(level=3 coeffset=tjm1 icrit=0.001 cap=7e-13)"


def tjm_text(deck, card=TJM1, ics=None):
    """The RSJ deck's text with its model card replaced by a level-3 card (and per-junction ics)."""
    txt = deck.to_spice()
    assert RSJ in txt
    txt = txt.replace(RSJ, card)
    if ics:
        lines = []
        for ln in txt.splitlines():
            tok = ln.split()
            if tok and tok[0] in ics:
                ln = " ".join(t for t in tok if not t.startswith("ics=")) + f" ics={ics[tok[0]]!r}"
            lines.append(ln)
        txt = "\n".join(lines) + "\n"
    return txt


CASES = {
    "single": (lambda: jj_single(), TJM1),
    "jtl6": (lambda: jtl(6), TJM1),
    "jtl6_tjm2_vshunt": (lambda: jtl(6), "jj(level=3 coeffset=tjm2 nterms=10 icrit=0.001 cap=7e-13 vshunt=0.5m)"),
    "jtl6_phase": (lambda: jtl(6, phase_nodes=True), TJM1),
    "jtl6_rtype0": (lambda: jtl(6), "jj(level=3 coeffset=tjm1 rtype=0 icrit=0.001 cap=7e-13)"),
    "jtl6_gear": (lambda: jtl(6).options(method=1), TJM1),
    "jtl6_dc": (lambda: jtl(6, dc_bias=True), TJM1),
    "jtl6_dc_phase_tjm2": (lambda: jtl(6, dc_bias=True, phase_nodes=True), "jj(level=3 coeffset=tjm2 nterms=10 icrit=0.001 cap=7e-13)"),
    "jtl12_vshunt": (lambda: jtl(12), "jj(level=3 coeffset=tjm1 icrit=0.001 cap=7e-13 vshunt=0.6m)"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_the_reference(name):
    mk, card = CASES[name]
    r = run_harvest_oracle(tjm_text(mk(), card))
    assert r["ref_err"] == 0 and r["rc"] == 0
    c, ref = r["counters"], r["ref"]
    assert (int(c[C_STATUS]), int(c[C_ACCEPTED]), int(c[C_REJECTED]), int(c[C_NUMITER]), int(c[C_TIMEPTS])) == \
        (0, ref["accepted"], ref["rejected"], ref["numiter"], ref["timepts"])
    assert r["vals"].shape == r["ref_vals"].shape
    assert np.abs(r["vals"] - r["ref_vals"]).max() < 1e-9


def _mc(ntr, seed=3):
    d = jtl(6)
    rng = np.random.default_rng(seed)
    jjs = [e for e in d.kind_tables()[DEV_JJ]]
    vals = {e.name: np.concatenate([[e.params["ics"]], e.params["ics"]*(1 + 0.1*rng.standard_normal(ntr - 1))])
            for e in jjs}
    return d, vals


def test_oracle_matches_the_reference_per_trial_ics():
    d, vals = _mc(4)
    counts = set()
    for t in range(4):
        ref = RefSim().run(tjm_text(d, ics={k: float(v[t]) for k, v in vals.items()}))
        r = run_harvest_oracle(tjm_text(d), [(P_TJM_ICS, k, float(v[t])) for k, v in vals.items()])
        assert r["rc"] == 0 and ref["err"] == 0
        c = r["counters"]
        assert (int(c[C_ACCEPTED]), int(c[C_REJECTED]), int(c[C_NUMITER])) == \
            (ref["accepted"], ref["rejected"], ref["numiter"])
        assert np.abs(r["vals"] - ref["vals"]).max() < 1e-9
        counts.add(ref["numiter"])
    assert len(counts) > 1


def _check(r, ntr, ref_of_trial):
    assert r["rc"] == 0
    cnt, w = r["counters"], r["waves"]
    for t in range(ntr):
        ref = ref_of_trial(t)
        ek = tuple(int(cnt[t, k]) for k in (C_STATUS, C_ACCEPTED, C_REJECTED, C_NUMITER, C_TIMEPTS))
        ok = (ref["error"], ref["accepted"], ref["rejected"], ref["numiter"], ref["timepts"])
        assert ek == ok, (t, ek, ok)
        want = ref["vals"]
        assert int(cnt[t, C_NPOINTS]) == want.shape[0]
        got = w[t, :want.shape[0]]
        tol = RELTOL*np.maximum(np.abs(got), np.abs(want)) + VNTOL
        assert (np.abs(got - want) <= tol).all(), (t, float((np.abs(got - want)/tol).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_engine_matches_the_same_circuits_dotask(name):
    mk, card = CASES[name]
    r = run_harvest(tjm_text(mk(), card), engine.LIB_PATH, 3)
    assert r["ref_err"] == 0
    ref = dict(r["ref"], vals=r["ref_vals"])
    _check(r, 3, lambda t: ref)


@pytest.mark.gpu
def test_engine_monte_carlo_per_trial_ics():
    ntr = 9
    d, vals = _mc(ntr, seed=8)
    r = run_harvest(tjm_text(d), engine.LIB_PATH, ntr, [(P_TJM_ICS, k, v) for k, v in vals.items()])
    sim = RefSim()

    def ref_of_trial(t):
        s = sim.run(tjm_text(d, ics={k: float(v[t]) for k, v in vals.items()}))
        return dict(s, error=s["err"])
    _check(r, ntr, ref_of_trial)


def test_tjm_without_uic_is_analysed():
    """A TJM deck without `uic` starts from a phase-mode DC operating point (tjmload.cc:78-120):
    the harvested deck is accepted by the engine's analysis (CPU part; the GPU run is
    tests/test_golden.py[jtl6_dc_tjm] and test_harvest.py)."""
    d = jtl(3, dc_bias=True)
    r = run_harvest(tjm_text(d), engine.LIB_PATH, 0)
    assert r["rc"] == 0
