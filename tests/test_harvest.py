"""Row (b), the drop-in boundary end to end: the REFERENCE's parser reads the deck, the
reference sets the circuit up, shim/harvest.cc walks that sCKT (model / instance lists, node
table, task options, .tran job) into a b200_deck and runs it through the C ABI -- no Python deck
builder between the reference and the engine.  The same call runs the reference's own doTask on
that very circuit; the two must agree: identical step / iteration counts per trial, waveforms
within reltol/vntol."""
import ctypes as C

import numpy as np
import pytest

from oracle.pyoracle import RefSim, have_ref, run_harvest
from tests.decks import variants
from xictools_b200 import engine
from xictools_b200.deck import (C_ACCEPTED, C_NPOINTS, C_NUMITER, C_REJECTED, C_STATUS, C_TIMEPTS, DEV_MOS,
                                P_MOS_L, P_MOS_VTO, mc_variation, ring_oscillator)

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
RELTOL, VNTOL = 1e-3, 1e-6

DECKS = ["ro5", "inv4", "lc_ladder", "transformer", "rc_sin", "rlc_series", "ro5_level2",
         "ro5_level3", "ro5_level6", "rlc_uic", "ro3_rd", "ro9_one_ic", "ro5_gear3", "ro17", "inv7",
         "ro5_trapcheck", "rc_exp", "rc_sffm", "rc_am", "rc_gpulse", "rc_gpulse_rpt", "rc_spulse", "rc_spulse_theta", "jtl6_dc_sfq", "jj_single", "jj_single_rtype3", "jtl6", "jtl6_phase", "jtl6_vshunt", "jtl12",
         "jtl6_accel", "jtl6_dc_accel", "squid_coupled", "squid_coupled_uic", "jtl6_dc", "jtl6_dc_phase", "jtl6_dc_vshunt", "jtl12_dc"]


@pytest.mark.parametrize("name", DECKS)
def test_harvested_deck_is_the_deck(name):
    """Host-only: the engine's analysis of the harvested deck (unknowns, pattern, LU fill,
    contribution count, state count, program sizes) equals that of the hand-built one."""
    d = variants()[name]
    r = run_harvest(d.to_spice(), engine.LIB_PATH, 0)
    assert r["ref_err"] == 0 and r["rc"] == 0
    L = engine.load_library()
    cd, keep = d.to_c()
    st = (C.c_longlong*32)()
    L.b200_deck_analysis_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    assert L.b200_deck_analysis_stats(C.byref(cd), st, 16) == 0
    assert r["stats"][:16] == list(st)[:16]
    assert r["stats"][0] == r["ref"]["size"]                 # matrix size of the reference's own setup


def test_harvest_refuses_what_is_not_built():
    txt = "diode\nv1 a 0 1\nd1 a 0 dm\n.model dm d\n.tran 1n 10n\n"
    L = RefSim().L
    r = run_harvest(txt, engine.LIB_PATH, 0)
    # the reference build here has no diode model either: the parse fails, nothing is harvested
    assert r["ref_err"] != 0 or r["rc"] == 13


def test_run_through_the_shim_without_a_gpu_fails_loudly():
    if engine.load_library().b200_cuda_device_count() > 0:
        pytest.skip("GPU present")
    r = run_harvest(ring_oscillator(5).to_spice(), engine.LIB_PATH, 2)
    assert r["ref_err"] == 0 and r["ref"]["accepted"] > 0
    assert r["rc"] == engine.E_NOGPU


def _check(r, ntr, ref_of_trial):
    assert r["rc"] == 0
    cnt, w = r["counters"], r["waves"]
    for t in range(ntr):
        ref = ref_of_trial(t)
        ek = tuple(int(cnt[t, k]) for k in (C_STATUS, C_ACCEPTED, C_REJECTED, C_NUMITER, C_TIMEPTS))
        ok = (ref["error"], ref["accepted"], ref["rejected"], ref["numiter"], ref["timepts"])
        assert ek == ok, (t, ek, ok)
        want = ref["vals"]
        assert int(cnt[t, C_NPOINTS]) == want.shape[0]       # output points actually written
        got = w[t, :want.shape[0]]
        assert got.shape == want.shape
        tol = RELTOL*np.maximum(np.abs(got), np.abs(want)) + VNTOL
        assert (np.abs(got - want) <= tol).all(), (t, float((np.abs(got - want)/tol).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", DECKS)
def test_parser_to_shim_to_engine_matches_the_same_circuits_dotask(name):
    d = variants()[name]
    r = run_harvest(d.to_spice(), engine.LIB_PATH, 3)
    assert r["ref_err"] == 0
    ref = dict(r["ref"], vals=r["ref_vals"])
    _check(r, 3, lambda t: ref)


@pytest.mark.gpu
def test_shim_on_the_chaotic_step_deck():
    """inv_lsupply (inductor in the supply, ~18 LTE rejections): step counts identical; the
    engine's pivot order differs from Sparse 1.3's, so roundoff may move a Newton iteration
    (same documented allowance as tests/test_golden.py)."""
    d = variants()["inv_lsupply"]
    r = run_harvest(d.to_spice(), engine.LIB_PATH, 2)
    assert r["ref_err"] == 0 and r["rc"] == 0
    for t in range(2):
        c = r["counters"][t]
        assert (int(c[C_ACCEPTED]), int(c[C_REJECTED]), int(c[C_TIMEPTS])) == \
            (r["ref"]["accepted"], r["ref"]["rejected"], r["ref"]["timepts"])
        assert abs(int(c[C_NUMITER]) - r["ref"]["numiter"]) <= 2


@pytest.mark.gpu
def test_monte_carlo_through_the_shim():
    """Per-trial vto / L rows addressed by the reference's instance names."""
    d = ring_oscillator(7)
    ntr = 6
    ov = mc_variation(d, ntr, seed=11)
    mos = d.kind_tables()[DEV_MOS]
    params = [(pid, mos[i].name, v) for (pid, i), v in ov.items()]
    r = run_harvest(d.to_spice(), engine.LIB_PATH, ntr, params)
    sim = RefSim()

    def ref_of_trial(t):
        s = sim.run(d.to_spice({k: v[t] for k, v in ov.items()}))
        return dict(s, vals=s["vals"])
    _check(r, ntr, ref_of_trial)


@pytest.mark.gpu
def test_resistor_monte_carlo_through_the_shim():
    """Per-trial resistor values (B200_P_RES_R) addressed by the reference's instance names, on a
    Josephson line biased through resistors (preloaded ones and ones loaded in place)."""
    from xictools_b200.deck import DEV_RES, P_RES_R
    d = variants()["jtl6_dc"]
    ntr = 5
    rng = np.random.default_rng(31)
    res = d.kind_tables()[DEV_RES]
    ov = {}
    for i, e in enumerate(res):
        v = np.full(ntr, e.params["r"])
        v[1:] *= 1.0 + 0.1*rng.standard_normal(ntr - 1)
        ov[(P_RES_R, i)] = v
    params = [(pid, res[i].name, v) for (pid, i), v in ov.items()]
    r = run_harvest(d.to_spice(), engine.LIB_PATH, ntr, params)
    sim = RefSim()

    def ref_of_trial(t):
        s = sim.run(d.to_spice({k: v[t] for k, v in ov.items()}))
        return dict(s, vals=s["vals"])
    _check(r, ntr, ref_of_trial)


PATTERN_DECKS = {
    # PULSE / GPULSE with bit patterns (inptran.cc:285-520): a finite string that repeats from its
    # second bit, a PRBS that runs forever, a string repeated twice and then silent
    "pulse_bits": "v1 a 0 pulse(0 1 0.1n 0.05n 0.05n 0.15n 0.5n b1011 rb=2 r=-1)",
    "pulse_prbs": "v1 a 0 pulse(0.2 1 0.05n 0.05n 0.05n 0.1n 0.4n bprbs7 r=-1)",
    "gpulse_bits": "v1 a 0 gpulse(0 1 0.2n 0.06n 0.4n b110 r=2)",
}


def _pattern_text(src_line):
    return f"pattern source\n{src_line}\nr1 a b 1k\nc1 b 0 0.1p\n.tran 0.02n 6n\n"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(PATTERN_DECKS))
def test_bit_patterns_through_the_shim_into_the_oracle(name):
    from oracle.pyoracle import run_harvest_oracle
    r = run_harvest_oracle(_pattern_text(PATTERN_DECKS[name]))
    assert r["rc"] == 0 and r["ref_err"] == 0
    c = r["counters"]
    assert tuple(int(c[k]) for k in (C_STATUS, C_ACCEPTED, C_REJECTED, C_NUMITER)) == \
        (0, r["ref"]["accepted"], r["ref"]["rejected"], r["ref"]["numiter"])
    assert np.abs(r["vals"] - r["ref_vals"]).max() < 1e-9
    assert np.ptp(r["ref_vals"][:, 0]) > 0.5          # the source really pulses


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PATTERN_DECKS))
def test_bit_patterns_through_the_shim_into_the_engine(name):
    r = run_harvest(_pattern_text(PATTERN_DECKS[name]), engine.LIB_PATH, 2)
    assert r["ref_err"] == 0
    ref = dict(r["ref"], vals=r["ref_vals"])
    _check(r, 2, lambda t: ref)
