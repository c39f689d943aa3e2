"""Generates tests/golden/*.json from the UNMODIFIED reference compiled headless (oracle/_ref).

Run in the build container (needs /root/reference so that oracle/_ref can be built):

    python tests/golden/make_golden.py

For every deck of tests/decks.py (and 4 Monte Carlo trials of RO-5) the reference's own
transient is run through its public entry points (sFtCirc/sCKT::doTask, see
oracle/refbuild/harness.cc) and the fixture stores: the step/iteration counters
(sSTATS: accepted, rejected, numiter, timepts, itercut, trapcut), the error code, the SPICE
text of the deck, the user time points and every saved node voltage / branch current at
those points (full double precision, repr round-trip).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.pyoracle import RefSim, have_ref          # noqa: E402
from tests.decks import op_decks, variants             # noqa: E402
from xictools_b200.deck import ring_oscillator, mc_variation   # noqa: E402

KEYS = ("err", "accepted", "rejected", "numiter", "timepts", "itercut", "trapcut")


def record(deck, overrides=None, analysis="tran"):
    text = deck.to_spice(overrides, analysis=analysis)
    r = RefSim().run(text)
    return {
        "spice": text,
        "counters": {k: int(r[k]) for k in KEYS},
        "names": list(r["names"]),
        "times": [float(x) for x in r["times"]],
        "vals": [[float(x) for x in row] for row in r["vals"]],
        "load_order": list(r["load_order"]),
    }


def main():
    assert have_ref(), "oracle/_ref is not built (needs /root/reference)"
    out = {}
    for name, d in sorted(variants().items()):
        out[name] = record(d)
    d = ring_oscillator(5)
    ov = mc_variation(d, 4, seed=3)
    for t in range(4):
        out[f"ro5_mc_seed3_trial{t}"] = record(d, {k: v[t] for k, v in ov.items()})
    for name, d in sorted(op_decks().items()):
        out[name] = record(d, analysis="op")
    for name, rec in out.items():
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(rec, f)
        print(name, rec["counters"], len(rec["times"]), "points x", len(rec["names"]), "signals")


if __name__ == "__main__":
    main()
