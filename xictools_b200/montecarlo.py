"""Monte Carlo dispatcher: the batched replacement of WRspice's trial loop.

Reference: `sCHECKprms::loop` / `trial` (wrspice/src/fte/check.cc:1472-1733) runs the trials of a
`check -m` / `.monte` analysis one after the other, re-sourcing the deck with the trial's
parameter values each time; the chained-DC variant `sDCTprms::loop_mt`
(analysis/dctprms.cc:924-1011) gives every worker thread its own sCKT.  Here every trial is a
column of one batch that shares the sparsity pattern: a rank owns the trials
`rank, rank + world, rank + 2*world, ...` (interleaved, so slow trials spread evenly -- SURVEY
section 8e), runs them to completion on its GPU with no data-path communication, and the only
exchange is the final gather of counters and selected waveforms to rank 0 (NCCL over
NVLink/NVSwitch on GPUs; any torch.distributed backend works, the CPU tests use gloo).

The per-rank runner is the CUDA batch (`engine.Batch`).  `runner=` exists so the host logic
(sharding, padding, gather order) can be exercised without a GPU by the tests.
"""
from __future__ import annotations

import numpy as np

from .deck import Deck, NCOUNTERS


def shard_trials(ntrials: int, world: int, rank: int) -> np.ndarray:
    """Global trial numbers owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return np.arange(rank, ntrials, world, dtype=np.int64)


def shard_capacity(ntrials: int, world: int) -> int:
    """Trials of the largest shard (gather buffers are padded to this)."""
    return (ntrials + world - 1)//world


class EngineRunner:
    """One CUDA batch on this rank's device (fails loudly without the CUDA library / a GPU)."""

    def __init__(self, deck: Deck, ntrials: int, device: int = 0):
        from . import engine
        self.batch = engine.Batch(deck, ntrials, device=device)

    def names(self):
        return self.batch.names()

    def run(self, overrides, sel):
        for (pid, idx), v in overrides.items():
            self.batch.set_param(pid, idx, v)
        self.batch.run_tran()
        cnt = self.batch.counters()
        times, w = self.batch.waves(sel)
        return cnt, times, w

    def close(self):
        self.batch.close()


class MonteCarlo:
    """Batched Monte Carlo transient of one deck.

    variation: {(param_id, instance_index): array[ntrials]} in GLOBAL trial order (trial 0 is
    conventionally the nominal circuit).  save: names of the unknowns to return
    (default: all).  After run(), rank 0 holds
        counters [ntrials][NCOUNTERS] int32, times [npoints], waves [ntrials][npoints][nsave]
    in global trial order; other ranks hold their own shard only.
    """

    def __init__(self, deck: Deck, ntrials: int, variation: dict | None = None, *, save=None,
                 rank: int = 0, world: int = 1, device: int = 0, runner=None,
                 local_variation: dict | None = None):
        self.deck, self.ntrials = deck, int(ntrials)
        self.rank, self.world = int(rank), int(world)
        self.mine = shard_trials(self.ntrials, self.world, self.rank)
        self.variation = {}
        for key, v in (variation or {}).items():
            v = np.ascontiguousarray(v, dtype=np.float64)
            if v.shape != (self.ntrials,):
                raise ValueError(f"variation {key}: expected {self.ntrials} values, got {v.shape}")
            self.variation[key] = np.ascontiguousarray(v[self.mine])
        # alternatively the caller hands over this rank's slice only (arrays of len(self.mine))
        for key, v in (local_variation or {}).items():
            v = np.ascontiguousarray(v, dtype=np.float64)
            if v.shape != (len(self.mine),):
                raise ValueError(f"local variation {key}: expected {len(self.mine)} values")
            self.variation[key] = v
        self.nlocal = len(self.mine)
        if self.nlocal == 0:
            raise ValueError("more ranks than trials")
        self.runner = runner(deck, self.nlocal) if runner else EngineRunner(deck, self.nlocal, device)
        names = self.runner.names()
        self.save = list(save) if save is not None else list(names)
        self.sel = [names.index(s) for s in self.save]
        self.counters = self.times = self.waves = None

    def run_local(self):
        cnt, times, w = self.runner.run(self.variation, self.sel)
        self.local = (np.asarray(cnt), np.asarray(times), np.asarray(w))
        return self.local

    def gather(self, device=None):
        """Gather the shards to rank 0 (collective; every rank must call it)."""
        cnt, times, w = self.local
        if self.world == 1:
            self.counters, self.times, self.waves = cnt, times, w
            return self
        import torch
        import torch.distributed as dist
        cap = shard_capacity(self.ntrials, self.world)
        npts, nsel = w.shape[1], w.shape[2]
        dev = device if device is not None else "cpu"
        pc = torch.zeros((cap, cnt.shape[1]), dtype=torch.int32, device=dev)
        pw = torch.zeros((cap, npts, nsel), dtype=torch.float64, device=dev)
        pc[:self.nlocal] = torch.from_numpy(np.ascontiguousarray(cnt, dtype=np.int32)).to(dev)
        pw[:self.nlocal] = torch.from_numpy(np.ascontiguousarray(w)).to(dev)
        gc = [torch.empty_like(pc) for _ in range(self.world)] if self.rank == 0 else None
        gw = [torch.empty_like(pw) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(pc, gc, dst=0)
        dist.gather(pw, gw, dst=0)
        if self.rank == 0:
            C = np.zeros((self.ntrials, cnt.shape[1]), dtype=np.int32)
            W = np.zeros((self.ntrials, npts, nsel))
            for r in range(self.world):
                own = shard_trials(self.ntrials, self.world, r)
                C[own] = gc[r][:len(own)].cpu().numpy()
                W[own] = gw[r][:len(own)].cpu().numpy()
            self.counters, self.times, self.waves = C, times, W
        else:
            self.counters, self.times, self.waves = cnt, times, w
        return self

    def run(self, device=None):
        self.run_local()
        return self.gather(device)

    def close(self):
        self.runner.close()


assert NCOUNTERS > 0
