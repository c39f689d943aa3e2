"""Large-deck probe (GPU): ring oscillator of N stages (2N MOSFETs), T trials, wide LU path.
usage: python tools/t_large.py N T [tstop_ns]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xictools_b200 import engine
from xictools_b200.deck import ring_oscillator, mc_variation, C_STATUS, C_ACCEPTED, C_REJECTED, C_NUMITER

N, T = int(sys.argv[1]), int(sys.argv[2])
tstop = float(sys.argv[3])*1e-9 if len(sys.argv) > 3 else 2e-9
d = ring_oscillator(N)
d.tran(0.02e-9, tstop)
t0 = time.time()
b = engine.Batch(d, T)
t1 = time.time()
if T > 1:
    b.set_params(mc_variation(d, T, seed=1))
t2 = time.time()
b.run_tran()
t3 = time.time()
c = b.counters()
st = b.stats() if hasattr(b, "stats") else None
print(json.dumps(dict(N=N, T=T, create_s=t1 - t0, params_s=t2 - t1, run_s=t3 - t2,
                      status=np.unique(c[:, C_STATUS]).tolist(),
                      trial0=[int(c[0, k]) for k in (C_ACCEPTED, C_REJECTED, C_NUMITER)],
                      trials_per_s=T/(t3 - t2),
                      stats=None if st is None else dict(ms_total=st.ms_total, ticks=int(st.ticks),
                          load_ms_per_tick=st.ms_load/max(st.load_launches, 1),
                          solve_ms_per_tick=st.ms_solve/max(st.solve_launches, 1)))))
