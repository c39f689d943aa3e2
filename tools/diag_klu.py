"""Diagnostic: reference + B200 KLU plug-in vs reference Sparse, per deck."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.pyoracle import run_ref_klu, RefSim
from tests.decks import variants
from xictools_b200 import engine
names = sys.argv[1:] or ["ro5", "rlc_series", "lc_ladder", "transformer", "rc_sin", "ro5_level3", "ro5_gear3"]
for name in names:
    txt = variants()[name].to_spice()
    r = run_ref_klu(txt, engine.LIB_PATH); s = RefSim().run(txt)
    n = min(len(r["times"]), len(s["times"]))
    dv = np.abs(r["vals"][:n] - s["vals"][:n]).max() if n else -1
    print(name, "err", r["err"], s["err"], [(r[k], s[k]) for k in ("accepted", "rejected", "numiter")], "maxdiff", dv)
    lines = [l for l in r["stderr"].splitlines() if "status" in l or "CUDA" in l]
    print("   ", len([l for l in r["stderr"].splitlines() if "klu_factor" in l]), "factor calls;", lines[:5])
