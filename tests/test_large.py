"""Large flat netlist (BASELINE.json configs[3]): 50001-stage ring oscillator, 100 002 MOSFETs,
50 003 unknowns.  The fixture tests/golden/large_ro50001.json holds the unmodified reference's
run of the first four Monte-Carlo trials (tests/golden/make_large_golden.py); the engine must
reproduce every trial's counts exactly and the saved signals within reltol / vntol."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RELTOL, VNTOL, ABSTOL = 1e-3, 1e-6, 1e-12


def _gold():
    with open(os.path.join(HERE, "golden", "large_ro50001.json")) as f:
        return json.load(f)


def test_fixture_shape():
    g = _gold()
    assert g["nstages"] == 50001 and len(g["trials"]) >= 4
    for r in g["trials"]:
        assert r["counters"]["err"] == 0 and len(r["vals"]) == len(r["times"]) > 50


@pytest.mark.gpu
def test_large_flat_matches_the_reference():
    import bench
    from xictools_b200 import engine
    from xictools_b200.deck import (ring_oscillator, C_ACCEPTED, C_NUMITER, C_REJECTED, C_STATUS)
    g = _gold()
    nref = len(g["trials"])
    d = ring_oscillator(g["nstages"])
    d.tran(0.02e-9, 2e-9)
    ov = bench.make_variation(d, g["ntrials"])
    b = engine.Batch(d, nref)
    b.set_params({k: v[:nref] for k, v in ov.items()})
    b.run_tran()
    cnt = b.counters()
    names = b.names()
    _, w = b.waves([names.index(x) for x in g["signals"]])
    b.close()
    for t, r in enumerate(g["trials"]):
        c = r["counters"]
        got = tuple(int(cnt[t, k]) for k in (C_STATUS, C_ACCEPTED, C_REJECTED, C_NUMITER))
        assert got == (c["err"], c["accepted"], c["rejected"], c["numiter"]), (t, got, c)
        ref = np.array(r["vals"])
        npt = ref.shape[0]
        tol = RELTOL*np.maximum(np.abs(w[t, :npt]), np.abs(ref))
        tol[:, :-1] += VNTOL
        tol[:, -1] += ABSTOL
        assert not (np.abs(w[t, :npt] - ref) > tol).any(), (t, float((np.abs(w[t, :npt] - ref)/tol).max()))
