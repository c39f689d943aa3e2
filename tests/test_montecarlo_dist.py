"""Host logic of the N>1 path: interleaved trial sharding and the final gather to rank 0
(SURVEY section 8e), exercised at world_size 2 over gloo on CPU.  The per-rank runner is
injected: here it is the CPU oracle (test-only), on a GPU box it is the CUDA batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from xictools_b200.deck import (C_ACCEPTED, C_NPOINTS, C_NUMITER, C_REJECTED, C_STATUS,   # noqa: E402
                                C_TIMEPTS, NCOUNTERS, mc_variation, ring_oscillator)
from xictools_b200.montecarlo import MonteCarlo, shard_capacity, shard_trials           # noqa: E402


class OracleRunner:
    """Test stand-in for engine.Batch: runs this rank's trials through the CPU oracle."""

    def __init__(self, deck, ntrials):
        from oracle.pyoracle import Oracle
        self.Oracle, self.deck, self.ntrials = Oracle, deck, ntrials
        self._names = Oracle(deck).names()

    def names(self):
        return self._names

    def run(self, overrides, sel):
        cnt = np.zeros((self.ntrials, NCOUNTERS), dtype=np.int32)
        waves = times = None
        for t in range(self.ntrials):
            r = self.Oracle(self.deck, overrides, t).run()
            if waves is None:
                times = r["times"]
                waves = np.zeros((self.ntrials, len(times), len(sel)))
            cnt[t, C_STATUS], cnt[t, C_ACCEPTED], cnt[t, C_REJECTED] = r["err"], r["accepted"], r["rejected"]
            cnt[t, C_NUMITER], cnt[t, C_TIMEPTS], cnt[t, C_NPOINTS] = r["numiter"], r["timepts"], len(r["times"])
            idx = [r["names"].index(self._names[s]) for s in sel]
            waves[t] = r["vals"][:, idx]
        return cnt, times, waves

    def close(self):
        pass


def test_shards_partition_the_trials():
    for n, w in ((10, 1), (10, 2), (11, 4), (7, 8), (10000, 8)):
        if n < w:
            continue
        allt = np.concatenate([shard_trials(n, w, r) for r in range(w)])
        assert sorted(allt.tolist()) == list(range(n))
        assert max(len(shard_trials(n, w, r)) for r in range(w)) == shard_capacity(n, w)
        assert shard_trials(n, w, 0)[0] == 0            # the nominal trial stays on rank 0
    with pytest.raises(ValueError):
        shard_trials(4, 2, 2)


NTRIALS = 5          # odd: ranks own 3 and 2 trials -> the gather pads the short shard


def _problem():
    d = ring_oscillator(3)
    return d, mc_variation(d, NTRIALS, seed=11)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, ov = _problem()
    mc = MonteCarlo(d, NTRIALS, ov, save=["n0", "vdd#branch"], rank=rank, world=world,
                    runner=OracleRunner)
    mc.run()
    if rank == 0:
        np.savez(out, counters=mc.counters, waves=mc.waves, times=mc.times)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_rank(tmp_path):
    d, ov = _problem()
    one = MonteCarlo(d, NTRIALS, ov, save=["n0", "vdd#branch"], runner=OracleRunner).run()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    assert got["counters"].shape == (NTRIALS, NCOUNTERS)
    np.testing.assert_array_equal(got["counters"], one.counters)
    np.testing.assert_array_equal(got["waves"], one.waves)
    np.testing.assert_array_equal(got["times"], one.times)
    # the trials really differ (the gather order matters)
    assert len({int(x) for x in one.counters[:, C_NUMITER]}) > 1


def test_variation_length_is_checked():
    d, ov = _problem()
    bad = {k: v[:-1] for k, v in ov.items()}
    with pytest.raises(ValueError):
        MonteCarlo(d, NTRIALS, bad, runner=OracleRunner)


def test_dc_sweep_builds_one_trial_per_point():
    """dc_sweep: every sweep value is a trial of a `.op` batch addressed at the source's DC row."""
    from xictools_b200.deck import P_SRC_DC
    from xictools_b200.montecarlo import dc_sweep
    from xictools_b200.workloads import opamp

    seen = {}

    class Probe:
        def __init__(self, deck, ntrials):
            seen["n"] = ntrials

        def names(self):
            return ["out", "y"]

        def run(self, overrides, sel):
            seen["ov"] = overrides
            cnt = np.zeros((seen["n"], NCOUNTERS), dtype=np.int32)
            return cnt, np.zeros(1), np.zeros((seen["n"], 1, len(sel)))

        def close(self):
            pass

    vals = np.linspace(1.5, 1.8, 7)
    mc = dc_sweep(opamp(), "vinp", vals, save=["out"], runner=Probe)
    mc.run()
    assert mc.analysis == "op" and seen["n"] == 7
    (pid, idx), row = next(iter(seen["ov"].items()))
    assert pid == P_SRC_DC and np.array_equal(row, vals)
    assert mc.waves.shape == (7, 1, 1)
    with pytest.raises(ValueError):
        dc_sweep(opamp(), "nosuch", vals, runner=Probe)
