"""Generates tests/golden/large_ro50001.json from the UNMODIFIED reference (oracle/_ref): the
first NREF Monte-Carlo trials of bench.py's `large_flat` workload (50001-stage ring oscillator,
100 002 level-1 MOSFETs, .tran 0.02n 2n; per-device vto / L from the reference's RNG stream,
trial 0 nominal) -- counters, three saved signals at every output point, and the reference's
wall time per trial on this container's cores.  ~2 min per trial; run in the build container:

    python tests/golden/make_large_golden.py
"""
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

NSTAGES, NTRIALS, NREF = 50001, 16, 4
WANT = ["n0", f"n{NSTAGES//2}", "vdd#branch"]


def _one(txt):
    from oracle.pyoracle import RefSim
    t0 = time.perf_counter()
    r = RefSim().run(txt)
    el = time.perf_counter() - t0
    idx = [r["names"].index(nm) for nm in WANT]
    return dict(counters={k: int(r[k]) for k in ("err", "accepted", "rejected", "numiter", "timepts")},
                seconds=el, times=[float(x) for x in r["times"]],
                vals=[[float(row[i]) for i in idx] for row in r["vals"]])


def main():
    import bench
    from xictools_b200.deck import ring_oscillator
    deck = ring_oscillator(NSTAGES)
    deck.tran(0.02e-9, 2e-9)
    ov = bench.make_variation(deck, NTRIALS)
    texts = [deck.to_spice({k: v[t] for k, v in ov.items()}) for t in range(NREF)]
    with mp.get_context("fork").Pool(NREF) as pool:
        res = pool.map(_one, texts)
    out = dict(nstages=NSTAGES, ntrials=NTRIALS, signals=WANT, trials=res,
               note="oracle/_ref (unmodified reference, Sparse 1.3), one trial per core, parse included")
    with open(os.path.join(HERE, "large_ro50001.json"), "w") as f:
        json.dump(out, f)
    for t, r in enumerate(res):
        print(t, r["counters"], f"{r['seconds']:.1f} s")


if __name__ == "__main__":
    main()
