import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from xictools_b200.montecarlo import MonteCarlo
from bench import make_deck, make_variation
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
T=int(sys.argv[1]) if len(sys.argv)>1 else 4096
deck=make_deck(501); ov=make_variation(deck,T,seed=100+rank)
mc=MonteCarlo(deck, world*T, save=["n0","n250","vdd#branch"], rank=rank, world=world, device=local, local_variation=ov)
for it in range(3):
    dist.barrier(); torch.cuda.synchronize(); t0=time.perf_counter()
    for (pid, idx), v in mc.variation.items():
        mc.runner.batch.set_param(pid, idx, v)
    t1=time.perf_counter()
    mc.runner.batch.run_tran(); t2=time.perf_counter()
    cnt=mc.runner.batch.counters(); tt,w=mc.runner.batch.waves(mc.sel); t3=time.perf_counter()
    mc.local=(cnt,tt,w)
    mc.gather(device=torch.device("cuda",local)); torch.cuda.synchronize(); t4=time.perf_counter()
    dist.barrier(); t5=time.perf_counter()
    print(f"rank {rank} it {it}: set_param {t1-t0:.3f} run {t2-t1:.3f} d2h {t3-t2:.3f} gather {t4-t3:.3f} barrier {t5-t4:.3f}", flush=True)
dist.destroy_process_group()
