#!/usr/bin/env python
"""Generates xictools_b200/data/tjm_setup.json: what the reference's setup makes of a level-3
(TJM) model card -- the post-setup model values, the instance quantities that do not depend on
the junction's area, and the renormalised coefficient set p / A / B (sTJMinstance::tjm_init1,
tjmset.cc:869-918; built-in tables of tjmcoeffs.cc).  A hand-built deck (bench.py's rsfq_tjm
workload) carries these numbers; decks that come through shim/harvest.cc get them live.
Run in the build container (needs oracle/_ref):  python tools/make_tjm_fixture.py"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import RefSim      # noqa: E402

CARDS = {
    "tjm1": "jj(level=3 coeffset=tjm1 icrit=0.001 cap=7e-13)",
    "tjm1_vshunt": "jj(level=3 coeffset=tjm1 icrit=0.001 cap=7e-13 vshunt=0.6m)",
    "tjm2": "jj(level=3 coeffset=tjm2 nterms=10 icrit=0.001 cap=7e-13)",
}
out = {}
L = RefSim().L
for name, card in CARDS.items():
    txt = f"tjm setup\ni1 0 1 pulse(0 1m 5p 10p)\nb1 1 0 jm\nr1 1 0 1\n.model jm {card}\n.tran 1p 10p uic\n"
    buf = (C.c_double*256)()
    n = L.wrsref_tjm_setup_dump(txt.encode(), buf, 256)
    assert n > 19, (name, n)
    v = list(buf[:n])
    keys_m = ("rtype", "ictype", "nterms", "criti", "cap", "cmu", "gmu", "r0", "rsint", "vshunt", "tsaccl",
              "tsfact", "vgnom")
    keys_i = ("tcf", "vdpbak", "omegaJ", "kgap", "kgap_rejpt", "alphaN")
    out[name] = {"card": card, "model": dict(zip(keys_m, v[:13])), "instance": dict(zip(keys_i, v[13:19])),
                 "coef": v[19:]}
    for k in ("rtype", "ictype", "nterms"):
        out[name]["model"][k] = int(out[name]["model"][k])
json.dump(out, open(os.path.join(ROOT, "xictools_b200", "data", "tjm_setup.json"), "w"), indent=1)
print({k: (v["model"]["nterms"], len(v["coef"])) for k, v in out.items()})
