"""`check -m` on the reference's own Monte-Carlo example, batched.

examples/JJexamples/nor_mc.cir of the reference runs 49 trials of an RSFQ NOR gate one after the
other, drawing eight critical currents, four bias resistors and one inductance per trial, and
writes `nor_mc.m00` with a PASS / FAIL line per trial.  This script does the same run as ONE
batch on the GPU through the reference-side shim (parser and circuit set-up by the reference,
harvest, per-trial rows by instance name, save list), then writes the verdict file in the
reference's layout (xictools_b200/mcreport.py) and trial 0 as a rawfile WRspice can `load`.

The example's own verdict compares the gate's SFQ pulse count with a logic signal built by its
script front end; here a trial passes if the output flip-flop (junction b3.x3) has emitted as
many flux quanta at the end of the run as in the nominal trial 0.

    python examples/nor_mc_batch.py [ntrials] [outdir]          (needs a B200 and oracle/_ref)
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.pyoracle import RefSim, run_harvest                 # noqa: E402
from xictools_b200 import engine                                # noqa: E402
from xictools_b200.deck import C_NPOINTS                        # noqa: E402
from xictools_b200.mcreport import verdicts, write_report, yield_fraction   # noqa: E402
from xictools_b200.rawfile import write_rawfile                  # noqa: E402
from xictools_b200.refrand import gauss_stream                   # noqa: E402


def sfq_count(phase):
    """JJdev::accept's pulse count of a total phase (jjacct.cc:86-114): one more quantum once
    the phase is 5*pi/4 into the next turn."""
    return np.floor((np.abs(phase) + 0.75*math.pi)/(2*math.pi)).astype(int)


def main():
    ntrials = int(sys.argv[1]) if len(sys.argv) > 1 else 49
    outdir = sys.argv[2] if len(sys.argv) > 2 else "."
    with open(os.path.join(ROOT, "tests", "golden", "nor_gate.json")) as f:
        gold = json.load(f)
    text = gold["spice"].replace(".tran 0.1p 0.25n", ".tran 0.1p 0.5n")
    names = RefSim().run(text.replace(".tran 0.1p 0.5n", ".tran 0.1p 1p"))["names"]   # unknown order
    keep = ["12.x3", "11.x3", "15", "3", "5"]        # phase nodes of b3.x3 / b2.x3, three voltages
    saves = [names.index(k) for k in keep]
    prm = gold["params"]
    g = gauss_stream(1, (ntrials - 1)*len(prm)).reshape(ntrials - 1, len(prm))
    rows, alters = [], [[] for _ in range(ntrials)]
    for k, p in enumerate(prm):
        v = np.full(ntrials, p["values"][0])
        v[1:] *= 1.0 + 0.1*g[:, k]
        rows.append((p["id"], p["instance"], v))
        dev, key = p["instance"], {80: "ics", 48: "r", 56: "l"}[p["id"]]
        for t in range(ntrials):
            alters[t].append((dev, key, float(v[t])))
    r = run_harvest(text, engine.LIB_PATH, ntrials, rows, saves=saves)
    if r["rc"]:
        raise SystemExit(f"shim error {r['rc']}")
    cnt, times, waves = r["counters"], r["times"], r["waves"]
    npts = int(cnt[0, C_NPOINTS])                   # output points the nominal trial wrote
    nominal = sfq_count(waves[0, npts - 1, 0])
    fail = verdicts(cnt, times, waves, lambda tt, w: sfq_count(w[npts - 1, 0]) == nominal)
    rep = write_report(os.path.join(outdir, "nor_mc.m00"), "nor_mc.cir", fail, alters)
    raw = write_rawfile(os.path.join(outdir, "nor_mc_trial0.raw"), "NOR gate, trial 0", keep, times[:npts],
                        waves[:1, :npts], binary=True)
    print(f"{ntrials} trials in {r['ms_total']:.0f} ms on the device ({r['ticks']} ticks); nominal output count "
          f"{int(nominal)}; yield {100*yield_fraction(fail):.1f} %; wrote {rep} and {raw}")


if __name__ == "__main__":
    main()
