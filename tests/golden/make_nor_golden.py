"""Generates tests/golden/nor_gate.json: the RSFQ NOR gate of the reference's own Monte-Carlo
example (wrspice/examples/JJexamples/nor_mc.cir: 26 junctions with phase nodes in five kinds of
cells, 177 unknowns, biased from a DC rail -- no `uic`), flattened for the headless harness by
tools/flatten_spice.py, with the example's own varied parameters (eight critical currents, four
bias resistors, one inductor of cell X0) drawn per trial.  What the example needs from the
reference's script front end is dropped: the .control/.exec blocks, `.stop ... call cb` and the
logic comparison source (a patterned PULSE); the PRBS bit patterns of the two data inputs
(`gpulse(... bprbs6 r=-1)`) are kept.  Every trial is run by the UNMODIFIED reference (oracle/_ref); the fixture
holds the flat deck, the per-trial values, the reference's counters and three signals at every
tenth output point.  Run in the build container:

    python tests/golden/make_nor_golden.py
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.pyoracle import RefSim                      # noqa: E402
from tools.flatten_spice import flatten                 # noqa: E402
from xictools_b200.deck import P_IND_L, P_JJ_ICS, P_RES_R   # noqa: E402

SRC = "/root/reference/wrspice/examples/JJexamples/nor_mc.cir"
NTRIALS, TSTOP, DECIM = 6, "0.25n", 10
SIGNALS = ["3", "5", "15"]            # jtl output, inverter output, flip-flop output (the example plots them)
# parameter of the example -> (engine parameter id, instance name in the flat deck)
VARIED = {**{f"b{k}ics": (P_JJ_ICS, f"b{k}.x0") for k in range(8)},
          **{f"r{k}r": (P_RES_R, f"r{k}.x0") for k in range(4)}, "l5l": (P_IND_L, "l5.x0")}


def source_text():
    src = open(SRC).read()
    src = "\n".join(l for l in src.splitlines() if not l.lower().startswith("vout"))
    return src.replace(".tran 0.1p 1n", f".tran 0.1p {TSTOP}")


def main():
    src = source_text()
    nominal = flatten(src)
    from tools.flatten_spice import parameters
    nom = parameters(src)
    rng = np.random.default_rng(2026)
    values = {k: np.full(NTRIALS, nom[k]) for k in VARIED}
    for k in VARIED:
        values[k][1:] *= 1.0 + 0.05*rng.standard_normal(NTRIALS - 1)
    trials = []
    cols = None
    for t in range(NTRIALS):
        r = RefSim().run(flatten(src, {k: values[k][t] for k in VARIED}))
        idx = [r["names"].index(s) for s in SIGNALS]
        assert cols in (None, idx)
        cols = idx
        trials.append(dict(counters={k: int(r[k]) for k in ("err", "accepted", "rejected", "numiter", "timepts")},
                           npoints=len(r["times"]),
                           vals=[[float(r["vals"][p, i]) for i in idx] for p in range(0, len(r["times"]), DECIM)]))
        print(t, trials[-1]["counters"])
    out = dict(spice=nominal, signals=SIGNALS, signal_cols=cols, decimation=DECIM, ntrials=NTRIALS,
               params=[dict(name=k, id=VARIED[k][0], instance=VARIED[k][1], values=[float(x) for x in values[k]])
                       for k in VARIED],
               trials=trials)
    with open(os.path.join(HERE, "nor_gate.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
