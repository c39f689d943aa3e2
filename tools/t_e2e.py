import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xictools_b200.montecarlo import MonteCarlo
from bench import make_deck, make_variation
T=10000
deck=make_deck(501); ov=make_variation(deck,T,seed=100)
mc=MonteCarlo(deck, T, save=["n0","n250","vdd#branch"], local_variation=ov)
b=mc.runner.batch
for it in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for (pid, idx), v in mc.variation.items():
        b.set_param(pid, idx, v)
    t1=time.perf_counter()
    b.run_tran(); t2=time.perf_counter()
    st=b.stats()
    cnt=b.counters(); t3=time.perf_counter()
    tt,w=b.waves(mc.sel); t4=time.perf_counter()
    print(f"it {it}: set_param {t1-t0:.3f} run_tran {t2-t1:.3f} (events {st.ms_total/1e3:.3f}) counters {t3-t2:.3f} waves {t4-t3:.3f}", flush=True)
