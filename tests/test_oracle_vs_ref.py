"""Pins the C oracle restatement (oracle/wrs_oracle.c) against the UNMODIFIED reference
compiled headless (oracle/_ref): identical accepted/rejected step counts and Newton
iteration counts, waveforms at every user time point within roundoff."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle, RefSim, have_ref
from tests.decks import op_decks, variants
from xictools_b200.deck import ring_oscillator, mc_variation

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")

KEYS = ("err", "accepted", "rejected", "numiter", "timepts", "itercut", "trapcut")
# order-1 integration with ~60 rejected steps is chaotic in step size: the counts still match
# but roundoff (different LU) is amplified exponentially along the run
LOOSE = {"ro5_ord1": 5e-2, "inv_lsupply": 1e-5, "ro5_level2": 1e-7, "ro3_level2_simple": 1e-7, "inv4_level2": 1e-7}     # inv_lsupply: 18 rejected steps, 1e-7 relative


@pytest.mark.parametrize("name", sorted(variants()))
def test_counts_and_waves_match_reference(name):
    d = variants()[name]
    ref = RefSim().run(d.to_spice())
    assert ref["load_order"] == d.load_order_names()
    o = Oracle(d).run()
    assert {k: o[k] for k in KEYS} == {k: ref[k] for k in KEYS}
    assert len(o["times"]) == len(ref["times"])
    np.testing.assert_allclose(o["times"], ref["times"], rtol=0, atol=1e-18)
    if not len(ref["times"]):
        return                              # a run that fails before its first output point
    idx = [o["names"].index(nm) for nm in ref["names"]]
    dv = np.abs(o["vals"][:, idx] - ref["vals"]).max()
    assert dv < LOOSE.get(name, 1e-9), dv


def test_monte_carlo_trials_match_reference():
    d = ring_oscillator(5)
    ov = mc_variation(d, 4, seed=3)
    for t in range(4):
        trial = {k: v[t] for k, v in ov.items()}
        ref = RefSim().run(d.to_spice(trial))
        o = Oracle(d, ov, t).run()
        assert {k: o[k] for k in KEYS} == {k: ref[k] for k in KEYS}
        idx = [o["names"].index(nm) for nm in ref["names"]]
        assert np.abs(o["vals"][:, idx] - ref["vals"]).max() < 1e-9


def test_jj_monte_carlo_trials_match_reference():
    """Per-trial critical currents (let @b1[ics] = ... of an .exec block): instance-level
    re-derivation of criti / cap / g0 / gn / gs / cr1 / cr2 (jjset.cc:613-705)."""
    from tests.decks import jtl
    from xictools_b200.deck import DEV_JJ, P_JJ_ICS
    d = jtl(6)
    rng = np.random.default_rng(5)
    ntr = 4
    ov = {}
    for i, e in enumerate(d.kind_tables()[DEV_JJ]):
        v = np.full(ntr, e.params["ics"])
        v[1:] *= 1.0 + 0.1*rng.standard_normal(ntr - 1)
        ov[(P_JJ_ICS, i)] = v
    counts = set()
    for t in range(ntr):
        ref = RefSim().run(d.to_spice({k: v[t] for k, v in ov.items()}))
        o = Oracle(d, ov, t).run()
        assert {k: o[k] for k in KEYS} == {k: ref[k] for k in KEYS}
        idx = [o["names"].index(nm) for nm in ref["names"]]
        assert np.abs(o["vals"][:, idx] - ref["vals"]).max() < 1e-9
        counts.add(ref["numiter"])
    assert len(counts) > 1


def test_uic_failure_mode_matches_reference():
    # the reference aborts this UIC start with E_MATHOVF (errors.h:119); so must the oracle
    d = ring_oscillator(5, uic=True)
    ref = RefSim().run(d.to_spice())
    o = Oracle(d).run()
    assert (o["err"], o["accepted"], o["numiter"]) == (ref["err"], ref["accepted"], ref["numiter"])


@pytest.mark.parametrize("name", sorted(op_decks()))
def test_dc_operating_point_matches_reference(name):
    """.op (dcoan.cc): same Newton iteration count (incl. gmin / source stepping) and the same
    operating point."""
    d = op_decks()[name]
    ref = RefSim().run(d.to_spice(analysis="op"))
    o = Oracle(d).run("op")
    assert (o["err"], o["numiter"]) == (ref["err"], ref["numiter"])
    assert len(o["times"]) == len(ref["times"]) == 1
    idx = [o["names"].index(nm) for nm in ref["names"]]
    # the ring oscillator's DC solution is its metastable point: a high-gain loop amplifies the
    # roundoff of the different LU pivot orders to ~1e-8 (vntol is 1e-6)
    tol = 1e-6 if name == "op_ro5" else 1e-9
    assert np.abs(o["vals"][:, idx] - ref["vals"]).max() < tol


def test_dc_sweep_points_match_reference():
    """Transfer curve of the inverter chain: each sweep point = one deck with its own vin."""
    from xictools_b200.deck import P_SRC_DC
    d = op_decks()["op_inv_dc"]
    isrc = [e.name for e in d.kind_tables()[3]].index("vin") if False else None
    from xictools_b200.deck import DEV_SRC
    isrc = [e.name for e in d.kind_tables()[DEV_SRC]].index("vin")
    for vin in (0.0, 0.9, 1.5, 1.65, 2.4, 3.3):
        ov = {(P_SRC_DC, isrc): vin}
        ref = RefSim().run(d.to_spice(ov, analysis="op"))
        o = Oracle(d, {(P_SRC_DC, isrc): [vin]}, 0).run("op")
        assert (o["err"], o["numiter"]) == (ref["err"], ref["numiter"]), vin
        idx = [o["names"].index(nm) for nm in ref["names"]]
        assert np.abs(o["vals"][:, idx] - ref["vals"]).max() < 1e-9


def _res_variation(d, ntr, seed, sigma=0.1):
    from xictools_b200.deck import DEV_RES, P_RES_R
    rng = np.random.default_rng(seed)
    ov = {}
    for i, e in enumerate(d.kind_tables()[DEV_RES]):
        v = np.full(ntr, e.params["r"])
        v[1:] *= 1.0 + sigma*rng.standard_normal(ntr - 1)
        ov[(P_RES_R, i)] = v
    return ov


@pytest.mark.parametrize("name", ["rlc_series", "ro3_rd", "jtl6_dc", "jtl6"])
def test_resistor_monte_carlo_trials_match_reference(name):
    """Per-trial resistor values (B200_P_RES_R): preloaded conductances (restemp.cc:209-229) and,
    on the phase nodes of a Josephson deck, resistors loaded in place (resload.cc:250-292)."""
    d = variants()[name]
    ntr = 3
    ov = _res_variation(d, ntr, 21)
    for t in range(ntr):
        ref = RefSim().run(d.to_spice({k: v[t] for k, v in ov.items()}))
        o = Oracle(d, ov, t).run()
        assert {k: o[k] for k in KEYS} == {k: ref[k] for k in KEYS}
        idx = [o["names"].index(nm) for nm in ref["names"]]
        assert np.abs(o["vals"][:, idx] - ref["vals"]).max() < LOOSE.get(name, 1e-9)


def _mos_variation(d, ntr, seed):
    """Every per-trial MOS parameter kind at once: W, KP, GAMMA, LAMBDA, PHI (plus the usual
    VTO, L), as private model cards / instance values per device."""
    from xictools_b200.deck import (DEV_MOS, P_MOS_GAMMA, P_MOS_KP, P_MOS_L, P_MOS_LAMBDA, P_MOS_PHI, P_MOS_VTO,
                                    P_MOS_W)
    rng = np.random.default_rng(seed)
    ov = {}
    for i, e in enumerate(d.kind_tables()[DEV_MOS]):
        mp = d.mos_models[e.model]["params"]
        nominal = {P_MOS_VTO: mp["vto"], P_MOS_KP: mp["kp"], P_MOS_GAMMA: mp.get("gamma", 0.4),
                   P_MOS_LAMBDA: mp.get("lambda", 0.02), P_MOS_PHI: mp.get("phi", 0.65),
                   P_MOS_L: e.params["l"], P_MOS_W: e.params["w"]}
        for pid, v0 in nominal.items():
            v = np.full(ntr, float(v0))
            v[1:] *= 1.0 + 0.04*rng.standard_normal(ntr - 1)
            ov[(pid, i)] = v
    return ov


def test_every_mos_parameter_kind_varied_matches_reference():
    d = ring_oscillator(5)
    ntr = 3
    ov = _mos_variation(d, ntr, 41)
    counts = set()
    for t in range(ntr):
        ref = RefSim().run(d.to_spice({k: v[t] for k, v in ov.items()}))
        o = Oracle(d, ov, t).run()
        assert {k: o[k] for k in KEYS} == {k: ref[k] for k in KEYS}
        idx = [o["names"].index(nm) for nm in ref["names"]]
        assert np.abs(o["vals"][:, idx] - ref["vals"]).max() < 1e-9
        counts.add(ref["numiter"])
    assert len(counts) > 1
