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

    accepts_packed = True

    def __init__(self, deck: Deck, ntrials: int, device: int = 0, analysis: str = "tran"):
        from . import engine
        self.batch = engine.Batch(deck, ntrials, device=device)
        self.packed = None              # (param ids, instance indices, values[nrows][ntrials])
        self._cnt = self._w = None
        self.analysis = analysis

    def _go(self):
        if self.analysis == "op":
            self.batch.run_op()
        else:
            self.batch.run_tran()

    def names(self):
        return self.batch.names()

    def set_saves(self, sel):
        """`.save`: the device keeps only these unknowns at the output points."""
        self.batch.set_saves(sel)

    def upload(self, overrides):
        if self.packed is not None:
            self.batch.set_params_packed(*self.packed)      # one transfer + a device scatter
        else:
            for (pid, idx), v in overrides.items():
                self.batch.set_param(pid, idx, v)

    def run(self, overrides, sel):
        self.upload(overrides)
        self._go()
        if self._cnt is None:
            self._cnt = _host_buffer((self.batch.ntrials, NCOUNTERS), np.int32)
        cnt = self.batch.counters(out=self._cnt)
        shape = (self.batch.ntrials, self.batch.npoints, len(sel))
        if self._w is None or self._w.shape != shape:
            self._w = _host_buffer(shape, np.float64)
        times, w = self.batch.waves(sel, out=self._w)
        return cnt, times, w

    def run_device(self, overrides, sel, dwaves_ptr, dcounters_ptr):
        """Run, leaving counters and selected waveforms in device buffers of the caller."""
        self.upload(overrides)
        self._go()
        self.batch.results_to_device(sel, dwaves_ptr, dcounters_ptr)
        return self.batch.times()

    def close(self):
        self.batch.close()


def _host_buffer(shape, dtype):
    """A reusable result / parameter buffer: page-locked when torch + CUDA are there (DMA both
    ways), plain numpy otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = {np.float64: torch.float64, np.int32: torch.int32}[dtype]
            t = torch.empty(tuple(shape), dtype=tdt, pin_memory=True)
            a = t.numpy()
            _KEEP.append(t)
            return a
    except Exception:
        pass
    return np.empty(shape, dtype=dtype)


_KEEP = []


class MonteCarlo:
    """Batched Monte Carlo transient of one deck.

    variation: {(param_id, instance_index): array[ntrials]} in GLOBAL trial order (trial 0 is
    conventionally the nominal circuit).  save: names of the unknowns to return
    (default: all).  After run(), rank 0 holds
        counters [ntrials][NCOUNTERS] int32, times [npoints], waves [ntrials][npoints][nsave]
    in global trial order; other ranks hold their own shard only.

    This rank's rows live in ONE [nrows][nlocal] array (`packed`; page-locked on a GPU box), of
    which `variation[key]` are views: the CUDA runner uploads it in a single transfer.
    """

    def __init__(self, deck: Deck, ntrials: int, variation: dict | None = None, *, save=None,
                 rank: int = 0, world: int = 1, device: int = 0, runner=None,
                 local_variation: dict | None = None, analysis: str = "tran"):
        self.deck, self.ntrials = deck, int(ntrials)
        self.analysis = analysis
        self.rank, self.world = int(rank), int(world)
        self.mine = shard_trials(self.ntrials, self.world, self.rank)
        self.nlocal = len(self.mine)
        if self.nlocal == 0:
            raise ValueError("more ranks than trials")
        rows = []
        for key, v in (variation or {}).items():
            v = np.ascontiguousarray(v, dtype=np.float64)
            if v.shape != (self.ntrials,):
                raise ValueError(f"variation {key}: expected {self.ntrials} values, got {v.shape}")
            rows.append((key, v[self.mine]))
        # alternatively the caller hands over this rank's slice only (arrays of len(self.mine))
        for key, v in (local_variation or {}).items():
            v = np.ascontiguousarray(v, dtype=np.float64)
            if v.shape != (self.nlocal,):
                raise ValueError(f"local variation {key}: expected {self.nlocal} values")
            rows.append((key, v))
        self.runner = runner(deck, self.nlocal) if runner else EngineRunner(deck, self.nlocal, device, analysis)
        packed_ok = getattr(self.runner, "accepts_packed", False)
        self.packed = (_host_buffer if packed_ok else np.empty)((len(rows), self.nlocal), np.float64)
        self.variation = {}
        for r, (key, v) in enumerate(rows):
            self.packed[r] = v
            self.variation[key] = self.packed[r]
        if packed_ok and rows:
            self.runner.packed = (np.array([k[0] for k, _ in rows], dtype=np.int32),
                                  np.array([k[1] for k, _ in rows], dtype=np.int32), self.packed)
        names = self.runner.names()
        self.save = list(save) if save is not None else list(names)
        self.sel = [names.index(s) for s in self.save]
        if save is not None and hasattr(self.runner, "set_saves"):
            self.runner.set_saves(self.sel)         # the output buffer holds the saved unknowns only
        self.counters = self.times = self.waves = None
        self._dev = None

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
        self._gather_tensors(pc, pw, times)
        if self.rank != 0:
            self.counters, self.times, self.waves = cnt, times, w
        return self

    def _gather_tensors(self, pc, pw, times):
        """pc [cap][NCOUNTERS], pw [cap][npts][nsel] on any device: gather to rank 0 and undo the
        interleaving there (trial i*world + r = row i of rank r) before the one copy to host."""
        import torch
        import torch.distributed as dist
        gc = gw = None
        if self.rank == 0:
            gc = list(torch.empty((self.world,) + tuple(pc.shape), dtype=pc.dtype, device=pc.device).unbind(0))
            gw = list(torch.empty((self.world,) + tuple(pw.shape), dtype=pw.dtype, device=pw.device).unbind(0))
        dist.gather(pc, gc, dst=0)
        dist.gather(pw, gw, dst=0)
        if self.rank == 0:
            C = torch.stack(gc, 1).reshape((-1,) + tuple(pc.shape[1:]))[:self.ntrials]
            W = torch.stack(gw, 1).reshape((-1,) + tuple(pw.shape[1:]))[:self.ntrials]
            if C.is_cuda:
                if self._dev is None or self._dev[1].shape != W.shape:
                    self._dev = (torch.empty(C.shape, dtype=C.dtype, pin_memory=True),
                                 torch.empty(W.shape, dtype=W.dtype, pin_memory=True))
                self._dev[0].copy_(C, non_blocking=True)
                self._dev[1].copy_(W, non_blocking=True)
                torch.cuda.synchronize()
                C, W = self._dev
            self.counters, self.times, self.waves = C.numpy(), times, W.numpy()

    def run(self, device=None):
        """Run this rank's shard and gather.  With a CUDA runner and world > 1 the results go
        from the engine's device buffers straight into the NCCL gather (no host round trip)."""
        if self.world > 1 and device is not None and hasattr(self.runner, "run_device") \
                and str(device).startswith("cuda"):
            import torch
            cap = shard_capacity(self.ntrials, self.world)
            b = self.runner.batch
            pc = torch.zeros((cap, NCOUNTERS), dtype=torch.int32, device=device)
            pw = torch.zeros((cap, b.npoints, len(self.sel)), dtype=torch.float64, device=device)
            times = self.runner.run_device(self.variation, self.sel, pw.data_ptr(), pc.data_ptr())
            self._gather_tensors(pc, pw, times)
            if self.rank != 0:
                self.counters, self.times, self.waves = pc[:self.nlocal], times, pw[:self.nlocal]
            return self
        self.run_local()
        return self.gather(device)

    def close(self):
        self.runner.close()


def dc_sweep(deck: Deck, source: str, values, *, save=None, rank: int = 0, world: int = 1,
             device: int = 0, runner=None) -> MonteCarlo:
    """Batched DC operating-point sweep (BASELINE.json configs[4]): every sweep value of the
    source's DC level is one trial of a `.op` batch, sharded over the ranks like Monte-Carlo
    trials -- the batched form of the chained-DC / loop analysis the reference threads one sCKT
    per worker (sDCTprms::loop_mt, analysis/dctprms.cc:924-1011; DCOanalysis::anFunc per point,
    dcoan.cc:100-121).  Returns the MonteCarlo object; run() gathers to rank 0, where
    waves[point][0][k] is unknown save[k] at sweep value values[point]."""
    from .deck import DEV_SRC, P_SRC_DC
    values = np.ascontiguousarray(values, dtype=np.float64)
    srcs = [e.name for e in deck.kind_tables()[DEV_SRC]]
    if source.lower() not in srcs:
        raise ValueError(f"no source named {source}")
    return MonteCarlo(deck, len(values), {(P_SRC_DC, srcs.index(source.lower())): values}, save=save,
                      rank=rank, world=world, device=device, runner=runner, analysis="op")


assert NCOUNTERS > 0
