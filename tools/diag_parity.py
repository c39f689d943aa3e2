"""Diagnostic: which MC trials differ from the oracle in counters, and is it a batch effect?"""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle.pyoracle import Oracle
from xictools_b200 import engine
from xictools_b200.deck import (C_ACCEPTED, C_NPOINTS, C_NUMITER, C_REJECTED, C_STATUS, C_TIMEPTS,
                                mc_variation, ring_oscillator)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 17
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 64
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 5
d = ring_oscillator(ns)
ov = mc_variation(d, nt, seed=seed)
b = engine.Batch(d, nt)
b.set_params(ov)
b.run_tran()
cnt = b.counters()
_, w = b.waves()
rp, cp = b.ordering()
rp2, cp2 = b.ordering_tran()
names = b.names()
bad = []
for t in range(nt):
    o = Oracle(d, ov, t); o.set_ordering(rp, cp); o.set_ordering_tran(rp2, cp2); r = o.run()
    ek = tuple(int(cnt[t, k]) for k in (C_ACCEPTED, C_REJECTED, C_NUMITER, C_TIMEPTS, C_NPOINTS))
    ok = (r["accepted"], r["rejected"], r["numiter"], r["timepts"], len(r["times"]))
    idx = [names.index(nm) for nm in r["names"]]
    npt = min(len(r["times"]), int(cnt[t, C_NPOINTS]))
    got = w[t, :npt][:, idx]
    md = np.abs(got - r["vals"][:npt]).max()
    if ek != ok:
        bad.append(t)
        print("trial", t, "engine", ek, "oracle", ok, "max wave diff %.3g" % md)
    elif md > 1e-9:
        print("trial", t, "same counts, wave diff %.3g" % md)
print("mismatching trials:", bad, "of", nt)
# single-trial batches for the mismatching ones
for t in bad[:4]:
    b1 = engine.Batch(d, 1)
    b1.set_params({k: v[t:t+1].copy() for k, v in ov.items()})
    b1.run_tran()
    c1 = b1.counters()
    print("  trial", t, "alone:", tuple(int(c1[0, k]) for k in (C_ACCEPTED, C_REJECTED, C_NUMITER, C_TIMEPTS, C_NPOINTS)))
    b1.close()
    # per accepted point iteration counts of the oracle
    o = Oracle(d, ov, t); o.set_ordering(rp, cp); o.set_ordering_tran(rp2, cp2); r = o.run()
    print("   oracle acc_iters", r.get("acc_iters", [])[:40])
