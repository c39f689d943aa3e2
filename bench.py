#!/usr/bin/env python
"""bench.py -- Monte-Carlo transient trials/sec of the batched Newton engine (BASELINE.json).

A "step" = one complete batched transient (DCOP + every accepted/rejected time step + output
interpolation) of T trials per GPU of the RO-501 CMOS ring oscillator (1002 MOSFETs + 501
capacitors + Vdd) with per-trial, per-device Vth0 / L variation: BASELINE.json configs[1]
at the 1k-device target size.  Trials shard across ranks with no data-path collective
(weak scaling); the only collective is the final gather of counters/waveforms.

  python bench.py --gpus 1 --steps 3 --warmup 3            our arm
  python bench.py --impl reference ...                     the reference's own CPU path
                                                           (oracle/_ref, all host cores)
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSTAGES = 501
TRIALS_PER_GPU = 10000
CLOAD, TSTEP, TSTOP = 0.1e-12, 0.1e-9, 10e-9          # 0.1 pF, .tran 0.1n 10n
SEED = 1
WORKLOAD = (f"ro{NSTAGES}_mc (SURVEY 8d M2): {NSTAGES}-stage CMOS ring oscillator (MOS level 1, "
            f"{2*NSTAGES} MOSFETs + {NSTAGES} x 0.1 pF), .tran 0.1n 10n, per-trial per-device "
            "vto (sigma 3%) and L (sigma 2%) from the reference's RNG stream "
            "(sRnd::gauss, seed 1 + rank; trial 0 nominal)")
SAVE_NODES = ["n0", f"n{NSTAGES//2}", "vdd#branch"]
RELTOL, VNTOL, ABSTOL = 1e-3, 1e-6, 1e-12


def bench_config(nstages, trials_per_gpu, world):
    """The workload description, identical in both arms (the reference arm times a bounded
    sample of the same job)."""
    return {"workload": WORKLOAD if nstages == NSTAGES else f"ro{nstages}_mc",
            "trials_per_gpu": trials_per_gpu, "unknowns": nstages + 2,
            "l2": "working set ~13 GB per GPU >> 126 MB L2 (no flush needed)"}


def make_deck(nstages=NSTAGES):
    from xictools_b200.deck import ring_oscillator
    return ring_oscillator(nstages, cload=CLOAD, tstep=TSTEP, tstop=TSTOP)


def make_variation(deck, ntrials, seed=SEED):
    """deck.mc_variation: per-trial, per-device vto / L from the reference's own deviates."""
    from xictools_b200.deck import mc_variation
    return mc_variation(deck, ntrials, seed=seed)


# ------------------------------------------------------------------------------------------
# reference CPU arm (oracle/_ref): one process per core, one trial per call, deck re-parsed per
# trial exactly as the reference's own Monte-Carlo loop does (fte/check.cc:1649-1733)
# ------------------------------------------------------------------------------------------
def _ref_worker(args):
    texts, want = args
    import numpy as np
    from oracle.pyoracle import RefSim
    sim = RefSim()
    out = []
    for txt in texts:
        r = sim.run(txt)
        w = None
        if want:
            idx = [r["names"].index(nm) for nm in want]
            w = np.ascontiguousarray(r["vals"][:, idx])
        out.append((r["err"], r["accepted"], r["rejected"], r["numiter"], r["t_load"],
                    r["t_decomp"], r["t_solve"], r["timepts"], w))
    return out


def run_reference_sample(ntrials, cores, seed=SEED, ov=None, want=None, nstages=NSTAGES):
    """Time `ntrials` trials of the workload on `cores` host processes (trial t takes the
    values ov[..][t]; by default the bench's own variation).  Returns dict."""
    from oracle.pyoracle import have_ref
    if not have_ref():
        raise RuntimeError("oracle/_ref/libwrsref.so missing")
    deck = make_deck(nstages)
    if ov is None:
        ov = make_variation(deck, ntrials, seed)
    texts = [deck.to_spice({k: v[t] for k, v in ov.items()}) for t in range(ntrials)]
    chunks = [(texts[i::cores], want) for i in range(cores)]
    chunks = [c for c in chunks if c[0]]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks)) as pool:
        pool.map(_ref_worker, [([texts[0]], None)]*len(chunks))      # warm (page in the .so)
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, chunks)
        el = time.perf_counter() - t0
    flat = [None]*ntrials
    for i, r in enumerate(res):
        for j, x in enumerate(r):
            flat[i + j*cores] = x
    ok = sum(1 for x in flat if x[0] == 0)
    return dict(seconds=el, trials=ntrials, ok=ok, cores=len(chunks),
                trials_per_s=ntrials/el,
                t_load=sum(x[4] for x in flat), t_decomp=sum(x[5] for x in flat),
                t_solve=sum(x[6] for x in flat), iters=sum(x[3] for x in flat), per_trial=flat)


def compare_with_reference(per_trial, cnt, waves, current_last=True, saved=None):
    """The parity record of the bench line: GPU trials 0..k-1 (counters [T][NCOUNTERS], waves
    [T][npoints][nsave] as the public API returned them) against the reference's own run of the
    same trials.  Counts: (error, accepted, rejected, numiter, timepts) identical per trial.
    Waves: every saved output point within reltol*max(|a|,|b|) + vntol (abstol for the branch
    current), reported as the worst error/tolerance ratio."""
    import numpy as np
    from xictools_b200.deck import (C_ACCEPTED, C_NUMITER, C_REJECTED, C_STATUS, C_TIMEPTS)
    k = len(per_trial)
    same = 0
    worst = 0.0
    wave_ok = 0
    off = []
    for t, r in enumerate(per_trial):
        err, acc, rej, nit, _, _, _, tpts, w = r
        ek = tuple(int(cnt[t, c]) for c in (C_STATUS, C_ACCEPTED, C_REJECTED, C_NUMITER, C_TIMEPTS))
        ok = (int(err), int(acc), int(rej), int(nit), int(tpts))
        if ek == ok:
            same += 1
        elif len(off) < 8:
            off.append({"trial": t, "engine": ek, "reference": ok})
        if w is None or err:
            continue
        npt = min(w.shape[0], waves.shape[1])
        got, ref = waves[t, :npt], w[:npt]
        tol = RELTOL*np.maximum(np.abs(got), np.abs(ref))
        if current_last:
            tol[:, :-1] += VNTOL
            tol[:, -1] += ABSTOL                  # the last saved unknown is a branch current
        else:
            tol += VNTOL
        ratio = float((np.abs(got - ref)/tol).max())
        worst = max(worst, ratio)
        wave_ok += int(ratio <= 1.0)
    return {"vs": "oracle/_ref (unmodified reference, Sparse 1.3 pivoting)", "trials_compared": k,
            "counts_identical": same, "waves_within_tol": wave_ok,
            "max_wave_err_over_tol": worst, "first_offenders": off,
            "tolerance": f"reltol {RELTOL} * max(|a|,|b|) + vntol {VNTOL} (abstol {ABSTOL} for branch currents), every output point of {saved or SAVE_NODES}"}


def bench_rsfq(local, ntrials=2048, njj=1000, tjm=None):
    """Secondary workload (BASELINE.json configs[2]): Monte-Carlo transient of a
    Josephson transmission line tiled to ~1000 junctions (RSJ level 1, the `jtl` cell of
    examples/JJexamples/nor_mc.cir), per-trial critical currents ics*(1+N(0,0.1)) from the
    reference's RNG stream, `.tran 0.5p 120p uic`.  One warm-up + one timed run, then the first
    trials are re-run through the unmodified reference for the parity record and the CPU rate."""
    import numpy as np
    from xictools_b200.workloads import jtl
    from xictools_b200 import engine
    from xictools_b200.deck import DEV_JJ, DEV_TJM, P_JJ_ICS, P_TJM_ICS, C_STATUS, C_NUMITER
    from xictools_b200.refrand import gauss_stream
    deck = jtl(njj, tjm=tjm)
    jjs = deck.kind_tables()[DEV_TJM if tjm else DEV_JJ]
    P_JJ_ICS = P_TJM_ICS if tjm else P_JJ_ICS
    g = gauss_stream(SEED, (ntrials - 1)*len(jjs)).reshape(ntrials - 1, len(jjs))
    vals = np.empty((len(jjs), ntrials))
    for i, e in enumerate(jjs):
        vals[i, 0] = e.params["ics"]
        vals[i, 1:] = e.params["ics"]*(1.0 + 0.1*g[:, i])
    b = engine.Batch(deck, ntrials, device=local)
    b.set_params_packed([P_JJ_ICS]*len(jjs), list(range(len(jjs))), vals)
    b.run_tran()
    b.run_tran()
    st = b.stats()
    cnt = b.counters()
    names = b.names()
    save = [names.index("a0"), names.index(f"a{njj//2}"), names.index(f"a{njj-1}")]
    _, w = b.waves(save)
    ok = int((cnt[:, C_STATUS] == 0).sum())
    kind = f"tunnel-junction model level 3, coefficient set {tjm}" if tjm else "RSJ level 1"
    out = {"workload": f"rsfq_mc: {njj}-junction Josephson transmission line ({kind}, {njj} JJ + "
                       f"{njj} L + {njj+1} R + {njj+1} I), .tran 0.5p 120p uic, per-trial ics sigma 10 %",
           "trials": ntrials, "unknowns": b.n, "ok_trials": ok,
           "value": ntrials/(st.ms_total*1e-3), "unit": "trials/s", "ms_per_step": st.ms_total,
           "ticks": int(st.ticks), "newton_iters_per_trial": float(cnt[:, C_NUMITER].mean()),
           "load_ms_per_tick": st.ms_load/max(st.load_launches, 1),
           "solve_ms_per_tick": st.ms_solve/max(st.solve_launches, 1),
           "roofline_load": {"achieved_GBs": st.algbytes_load_per_launch/(st.ms_load/max(st.load_launches, 1)*1e-3)/1e9
                             * min(1.0, float(cnt[:, C_NUMITER].mean())/max(1.0, float(st.ticks)))}}
    try:
        cores = host_cores()
        k = min(ntrials, max(cores, 16))
        ov = {(P_JJ_ICS, i): vals[i, :k] for i in range(len(jjs))}
        texts = [deck.to_spice({key: v[t] for key, v in ov.items()}) for t in range(k)]
        chunks = [(texts[i::cores], ["a0", f"a{njj//2}", f"a{njj-1}"]) for i in range(cores)]
        chunks = [c for c in chunks if c[0]]
        ctx = mp.get_context("fork")
        with ctx.Pool(len(chunks)) as pool:
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, chunks)
            el = time.perf_counter() - t0
        flat = [None]*k
        for i, r in enumerate(res):
            for j, x in enumerate(r):
                flat[i + j*len(chunks)] = x
        out["cpu_baseline"] = {"value": k/el, "unit": "trials/s", "cores": len(chunks), "kind": "reference",
                               "sample": f"trials 0..{k-1} in {el:.1f} s (oracle/_ref, Sparse 1.3)"}
        out["parity"] = compare_with_reference(flat, cnt, w, current_last=False,
                                               saved=["a0", f"a{njj//2}", f"a{njj-1}"])
    except Exception as e:          # pragma: no cover
        out["parity"] = {"error": repr(e)}
    b.close()
    return out


def bench_large_flat(local, ntrials=16, time_reference=False):
    """Secondary workload (BASELINE.json configs[3]): a large flat CMOS netlist -- a ring
    oscillator of 50001 stages, 100 002 level-1 MOSFETs, 50 003 unknowns -- run as a single
    transient and as a small Monte-Carlo batch.  The LU image (400 k doubles per trial) is far
    beyond one warp's shared-memory window: this runs the wide task program (k_lu_wide: one CTA
    per trial, image in L2 / HBM).  Parity: the first trials against tests/golden/
    large_ro50001.json, which tests/golden/make_large_golden.py wrote from the unmodified
    reference (2 - 10 minutes of CPU per trial, so it is not re-run inside the default bench;
    `--large-ref` times trial 0 on this box's host)."""
    import json as _json
    import numpy as np
    from xictools_b200 import engine
    from xictools_b200.deck import ring_oscillator, C_STATUS, C_NUMITER
    with open(os.path.join(ROOT, "tests", "golden", "large_ro50001.json")) as f:
        gold = _json.load(f)
    nstages = gold["nstages"]
    assert ntrials == gold["ntrials"]
    want = gold["signals"]
    deck = ring_oscillator(nstages)
    deck.tran(0.02e-9, 2e-9)
    ov = make_variation(deck, ntrials)
    out = {"workload": f"large_flat: {nstages}-stage CMOS ring oscillator, {2*nstages} MOSFETs (level 1) + {nstages} C, "
                       f".tran 0.02n 2n; single transient and a {ntrials}-trial Monte Carlo (vto / L as the headline workload)"}
    pool = pending = None
    if time_reference:
        pool = mp.get_context("fork").Pool(1)
        t_ref0 = time.perf_counter()
        pending = pool.apply_async(_ref_worker, (([deck.to_spice({k: v[0] for k, v in ov.items()})], want),))
    try:
        b1 = engine.Batch(deck, 1, device=local)
        b1.run_tran()
        s1 = b1.stats()
        c1 = b1.counters()
        out["single"] = {"seconds": s1.ms_total*1e-3, "ok": int(c1[0, C_STATUS] == 0), "newton_iters": int(c1[0, C_NUMITER]),
                         "ms_per_newton_iter": s1.ms_total/max(1, int(c1[0, C_NUMITER])), "unknowns": b1.n,
                         "load_ms_per_tick": s1.ms_load/max(s1.load_launches, 1),
                         "solve_ms_per_tick": s1.ms_solve/max(s1.solve_launches, 1)}
        b1.close()
        b = engine.Batch(deck, ntrials, device=local)
        b.set_params(ov)
        b.run_tran()
        st = b.stats()
        cnt = b.counters()
        names = b.names()
        _, w = b.waves([names.index(x) for x in want])
        out.update({"trials": ntrials, "ok_trials": int((cnt[:, C_STATUS] == 0).sum()),
                    "value": ntrials/(st.ms_total*1e-3), "unit": "trials/s", "ms_per_step": st.ms_total,
                    "ticks": int(st.ticks),
                    "load_ms_per_tick": st.ms_load/max(st.load_launches, 1),
                    "solve_ms_per_tick": st.ms_solve/max(st.solve_launches, 1),
                    "device_load_instances_per_s": (3*nstages + 1)*float(cnt[:, C_NUMITER].sum())/(st.ms_total*1e-3)})
        b.close()
        flat = []
        for r in gold["trials"]:
            c = r["counters"]
            flat.append((c["err"], c["accepted"], c["rejected"], c["numiter"], 0, 0, 0, c["timepts"], np.array(r["vals"])))
        out["parity"] = compare_with_reference(flat, cnt, w, current_last=True, saved=want)
        out["parity"]["vs"] = "tests/golden/large_ro50001.json (written by the unmodified reference, oracle/_ref)"
        secs = [r["seconds"] for r in gold["trials"]]
        out["cpu_baseline"] = {"value": len(secs)/max(secs), "unit": "trials/s", "cores": len(secs), "kind": "reference",
                               "sample": f"RECORDED when the fixture was written (build container, 4 trials side by side, one per core): "
                                         f"{min(secs):.0f}-{max(secs):.0f} s per trial with the per-device .model cards of the Monte Carlo; "
                                         f"the nominal circuit with one shared model card alone: {gold.get('single_nominal_seconds', 0):.0f} s"}
        if pending is not None:
            pending.get(timeout=3000)
            el = time.perf_counter() - t_ref0
            out["cpu_baseline"] = {"value": 1.0/el, "unit": "trials/s", "cores": 1, "kind": "reference",
                                   "sample": f"trial 0 on this host in {el:.1f} s (oracle/_ref, Sparse 1.3)"}
    except Exception as e:          # pragma: no cover
        out["error"] = repr(e)
    finally:
        if pool is not None:
            pool.terminate()
    return out


def bench_nor_mc(local, ntrials=16384, nref=16):
    """Secondary workload: the reference's OWN Monte-Carlo example (examples/JJexamples/
    nor_mc.cir, an RSFQ NOR gate: 26 junctions with phase nodes in five cell types, 177 unknowns,
    started from the phase-mode DC operating point, `.tran 0.1p 1n`), flattened for the headless
    harness (tests/golden/nor_gate.json holds the flat deck).  Its 13 varied parameters -- eight
    critical currents, four bias resistors, an inductor -- get per-trial values (sigma 5 %, the
    reference's RNG stream) addressed by instance name through shim/harvest.cc: the deck text is
    parsed and set up by the reference, harvested, and run by the batched engine.  The first
    `nref` trials are re-run by the unmodified reference for the parity record and the CPU rate."""
    import json as _json
    import numpy as np
    from oracle.pyoracle import run_harvest
    from xictools_b200 import engine
    from xictools_b200.refrand import gauss_stream
    from tools.flatten_spice import parameters      # noqa: F401  (the fixture carries the nominal values)
    with open(os.path.join(ROOT, "tests", "golden", "nor_gate.json")) as f:
        gold = _json.load(f)
    text = gold["spice"].replace(".tran 0.1p 0.25n", ".tran 0.1p 1n")
    assert ".tran 0.1p 1n" in text
    prm = gold["params"]
    gs = gauss_stream(SEED, (ntrials - 1)*len(prm)).reshape(ntrials - 1, len(prm))
    vals = []
    for k, p in enumerate(prm):
        v = np.full(ntrials, p["values"][0])
        v[1:] *= 1.0 + 0.05*gs[:, k]
        vals.append(v)
    # `.save` of the three signals the example plots: 10 001 output points x 177 unknowns x 16 384
    # trials would be 232 GB, the saved three are 3.9 GB
    cols = gold["signal_cols"]
    r = run_harvest(text, engine.LIB_PATH, ntrials, [(p["id"], p["instance"], v) for p, v in zip(prm, vals)], saves=cols)
    if r["rc"]:
        return {"error": f"shim rc {r['rc']}"}
    cnt = r["counters"]
    out = {"workload": "nor_mc: the RSFQ NOR gate of examples/JJexamples/nor_mc.cir (26 JJ, 177 unknowns, phase-mode DC start), "
                       ".tran 0.1p 1n, the example's 13 varied parameters (ics, R, L) sigma 5 % per trial, .save of 3 signals; parser -> shim -> engine",
           "trials": ntrials, "ok_trials": int((cnt[:, 0] == 0).sum()),
           "value": ntrials/(r["ms_total"]*1e-3), "unit": "trials/s", "ms_per_step": r["ms_total"], "ticks": int(r["ticks"]),
           "newton_iters_per_trial": float(cnt[:, 3].mean())}
    try:
        cores = host_cores()
        k = min(nref, ntrials)
        texts = []
        for t in range(k):
            tx = text
            for p, v in zip(prm, vals):
                # the flat deck carries each varied value once, on its instance line
                inst = p["instance"]
                lines = tx.splitlines()
                for i, ln in enumerate(lines):
                    tok = ln.split()
                    if tok and tok[0] == inst:
                        if inst[0] == "b":
                            tok = [x for x in tok if not x.startswith("ics=")] + [f"ics={float(v[t])!r}"]
                        else:
                            tok[3] = repr(float(v[t]))
                        lines[i] = " ".join(tok)
                tx = "\n".join(lines) + "\n"
            texts.append(tx)
        chunks = [(texts[i::cores], None) for i in range(cores)]
        chunks = [c for c in chunks if c[0]]
        with mp.get_context("fork").Pool(len(chunks)) as pool:
            t0 = time.perf_counter()
            res = pool.map(_ref_worker_cols, [(c[0], cols) for c in chunks])
            el = time.perf_counter() - t0
        flat = [None]*k
        for i, rr in enumerate(res):
            for j, x in enumerate(rr):
                flat[i + j*len(chunks)] = x
        out["cpu_baseline"] = {"value": k/el, "unit": "trials/s", "cores": len(chunks), "kind": "reference",
                               "sample": f"trials 0..{k-1} in {el:.1f} s (oracle/_ref, Sparse 1.3)"}
        same = steps_same = wave_ok = 0
        worst = 0.0
        diffs = {}
        for t, x in enumerate(flat):
            err, acc, rej, nit, w = x
            steps_same += int((int(cnt[t, 0]), int(cnt[t, 1]), int(cnt[t, 2])) == (err, acc, rej))
            same += int(int(cnt[t, 3]) == nit)
            dk = str(int(cnt[t, 3]) - nit)
            diffs[dk] = diffs.get(dk, 0) + 1
            got = r["waves"][t][:w.shape[0]]
            tol = RELTOL*np.maximum(np.abs(got), np.abs(w)) + VNTOL
            ratio = float((np.abs(got - w)/tol).max())
            worst = max(worst, ratio)
            wave_ok += int(ratio <= 1.0)
        out["parity"] = {"vs": "oracle/_ref (unmodified reference, Sparse 1.3 pivoting)", "trials_compared": k,
                         "steps_identical": steps_same, "newton_iterations_identical": same,
                         "newton_iterations_engine_minus_reference": diffs,
                         "newton_iterations_note": "the reference re-orders once during the gmin stepping of this deck's operating point "
                                                   "(its own trace: 'Matrix is singular, iterno=1 / Forced reordering', niiter.cc:303-331) "
                                                   "and spends one more iteration there than the engine's static sequence: engine = reference - 1; the reference itself, run through the KLU plug-in, takes the engine's count (tests/test_nor.py)",
                         "waves_within_tol": wave_ok, "max_wave_err_over_tol": worst}
    except Exception as e:          # pragma: no cover
        out["parity"] = {"error": repr(e)}
    return out


def _ref_worker_cols(args):
    texts, cols = args
    import numpy as np
    from oracle.pyoracle import RefSim
    sim = RefSim()
    out = []
    for txt in texts:
        r = sim.run(txt)
        out.append((r["err"], r["accepted"], r["rejected"], r["numiter"], np.ascontiguousarray(r["vals"][:, cols])))
    return out


def bench_dc_sweep(local, npoints=100000):
    """Secondary workload (BASELINE.json configs[4]): 1e5-point DC operating-point sweep of the
    two-stage opamp (MOS level 1), every point a trial of one `.op` batch."""
    import numpy as np
    from xictools_b200.montecarlo import dc_sweep
    from xictools_b200.workloads import opamp
    from xictools_b200.deck import C_STATUS, C_NUMITER
    deck = opamp()
    vin = np.linspace(1.55, 1.75, npoints)
    mc = dc_sweep(deck, "vinp", vin, save=["out", "y", "tail"], device=local)
    mc.run()
    t0 = time.perf_counter()
    mc.run()
    el = time.perf_counter() - t0
    st = mc.runner.batch.stats()
    cnt = mc.counters
    out = {"workload": f"dc_sweep: two-stage CMOS opamp (9 MOSFETs level 1), {npoints} operating points "
                       "(vinp 1.55 .. 1.75 V), gmin stepping per point",
           "points": npoints, "ok_points": int((cnt[:, C_STATUS] == 0).sum()),
           "value": npoints/(st.ms_total*1e-3), "unit": "points/s", "e2e_value": npoints/el,
           "ms_device": st.ms_total, "ticks": int(st.ticks),
           "newton_iters_per_point": float(cnt[:, C_NUMITER].mean())}
    try:
        from oracle.pyoracle import RefSim
        from xictools_b200.deck import DEV_SRC, P_SRC_DC
        isrc = [e.name for e in deck.kind_tables()[DEV_SRC]].index("vinp")
        sim = RefSim()
        idx = np.linspace(0, npoints - 1, 64).astype(int)
        same = 0
        worst = 0.0
        t0 = time.perf_counter()
        for p in idx:
            r = sim.run(deck.to_spice({(P_SRC_DC, isrc): float(vin[p])}, analysis="op"))
            same += int((r["err"], r["numiter"]) == (int(cnt[p, C_STATUS]), int(cnt[p, C_NUMITER])))
            ref = np.array([r["vals"][0][r["names"].index(nm)] for nm in ("out", "y", "tail")])
            got = mc.waves[p, 0]
            worst = max(worst, float((np.abs(got - ref)/(RELTOL*np.maximum(np.abs(got), np.abs(ref)) + VNTOL)).max()))
        el_ref = time.perf_counter() - t0
        out["parity"] = {"vs": "oracle/_ref", "points_compared": len(idx), "iterations_identical": same,
                         "max_err_over_tol": worst}
        out["cpu_baseline"] = {"value": len(idx)/el_ref, "unit": "points/s", "cores": 1, "kind": "reference",
                               "sample": f"{len(idx)} points, deck re-parsed per point"}
    except Exception as e:          # pragma: no cover
        out["parity"] = {"error": repr(e)}
    mc.close()
    return out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    # bounded sample: ~0.2 s per trial per core
    per_step = max(cores*64, 256)
    for _ in range(args.warmup):
        run_reference_sample(max(cores, 8), cores)
    t_tot, n_tot, last = 0.0, 0, None
    for k in range(args.steps):
        last = run_reference_sample(per_step, cores, seed=SEED + k)
        t_tot += last["seconds"]
        n_tot += last["trials"]
    value = n_tot/t_tot
    line = {
        "impl": "reference",
        "metric": "Monte Carlo transient trials/sec", "value": value, "unit": "trials/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3*t_tot/args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(NSTAGES, TRIALS_PER_GPU, args.gpus),
        "run": {"trials_per_step": per_step,
                "solver": "Sparse 1.3 (what the snapshot executes; KLU source absent)"},
        "cpu_baseline": {"value": value, "unit": "trials/s", "cores": last["cores"],
                         "kind": "reference",
                         "sample": f"{per_step} trials/step of the same deck, one process per core, "
                                   "deck re-parsed per trial (oracle/_ref = unmodified reference)"},
        "e2e": {"value": value, "unit": "trials/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.samples.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm)//2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bench_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from xictools_b200.deck import C_NUMITER, C_STATUS
    from xictools_b200.montecarlo import MonteCarlo

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    T = args.trials
    deck = make_deck(args.stages)
    ov = make_variation(deck, T, seed=SEED + rank)       # this rank's shard of the world*T trials
    save = ["n0", f"n{args.stages//2}", "vdd#branch"]        # = SAVE_NODES at the default size
    mc = MonteCarlo(deck, world*T, save=save, rank=rank, world=world, device=local,
                    local_variation=ov)
    batch = mc.runner.batch
    # mc.packed: this rank's [nrows][T] parameter rows in ONE page-locked array (H2D = one DMA)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_kernel_only():
        batch.run_tran()

    def step_e2e():
        # the call a user makes: host parameter arrays in (H2D), counters + selected waveforms
        # out (D2H), shards gathered to rank 0 over NCCL/NVLink when N > 1
        mc.run(device=torch.device("cuda", local))
        if world == 1:
            return mc.local[0], mc.local[2]
        if rank == 0:                       # rank 0's own trials are rows 0, world, 2*world, ...
            return mc.counters[0::world], mc.waves[0::world]
        return None, None

    mc.runner.upload(mc.variation)
    for _ in range(args.warmup):
        step_kernel_only()

    # ---- kernel-resident timing: CUDA events on the engine's own stream (ms_total of
    # b200_batch_get_stats brackets every launch of the run), wall clock as a cross-check ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    launches = ticks = 0
    ms_load = ms_solve = ms_dev = 0.0
    nload = nsolve = 0
    for _ in range(args.steps):
        step_kernel_only()
        st = batch.stats()
        launches += st.launches
        ticks += st.ticks
        ms_dev += st.ms_total
        ms_load += st.ms_load
        ms_solve += st.ms_solve
        nload += st.load_launches
        nsolve += st.solve_launches
    barrier()
    el_wall = time.perf_counter() - t0
    el = ms_dev*1e-3
    clocks = sampler.stop() if rank == 0 else None
    st = batch.stats()
    cnt = batch.counters()
    ok_trials = int((cnt[:, C_STATUS] == 0).sum())

    if args.kernel_only:
        if rank == 0:
            print(f"bench: kernel-only {world*T*args.steps/el:.1f} trials/s, {1e3*el/args.steps:.1f} ms/step, "
                  f"load {ms_load/max(nload, 1):.3f} solve {ms_solve/max(nsolve, 1):.3f} ms/tick, "
                  f"{ticks//args.steps} ticks", file=sys.stderr)
        return 0

    # ---- end-to-end timing (host buffers, public API) ----
    for _ in range(max(1, min(args.warmup, 2))):     # NCCL connection set-up, first-touch of host buffers
        step_e2e()
    sampler2 = ClockSampler(local)
    if rank == 0:
        sampler2.start()
    barrier()
    t1 = time.perf_counter()
    e2e_steps = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        cnt_e, w_e = step_e2e()
        e2e_steps.append(time.perf_counter() - ts)
    barrier()
    el_e2e = time.perf_counter() - t1
    clocks_e2e = sampler2.stop() if rank == 0 else None
    if rank == 0:
        print("bench: e2e seconds per step", ["%.3f" % x for x in e2e_steps],
              "engine events of the last one %.3f" % (batch.stats().ms_total*1e-3), file=sys.stderr)

    # ---- strong scaling in the same run (N > 1): the 10 000-trial job split over the N GPUs ----
    strong = None
    if world > 1:
        from xictools_b200 import engine as _eng
        Ts = (TRIALS_PER_GPU + world - 1)//world
        Ts = min(Ts, T)
        bs = _eng.Batch(deck, Ts, device=local)
        bs.set_params({key: v[:Ts] for key, v in mc.variation.items()})
        bs.run_tran()
        barrier()
        ms_s = 0.0
        for _ in range(args.steps):
            bs.run_tran()
            ms_s += bs.stats().ms_total
        barrier()
        ts_ = torch.tensor([ms_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
        bs.close()
        v_s = world*Ts*args.steps/(float(ts_[0])*1e-3)
        strong = {"total_trials": world*Ts, "trials_per_gpu": Ts, "value": v_s, "unit": "trials/s",
                  "ms_per_step": float(ts_[0])/args.steps,
                  "note": "same deck, the 10 000-trial job split over the GPUs (device-timed, max "
                          "over ranks); efficiency = this rate / the weak-scaling rate of the same "
                          "run; what limits it is wave quantisation of the per-trial kernels at "
                          f"{Ts} trials (grids of {Ts}//32 warps on 148 SMs)"}

    tmax = torch.tensor([el, el_e2e, el_wall], dtype=torch.float64, device="cuda")
    ninst = 2*args.stages + args.stages + 1
    agg = torch.tensor([float(ok_trials), float(launches), float(cnt[:, C_NUMITER].sum())],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    el, el_e2e, el_wall = float(tmax[0]), float(tmax[1]), float(tmax[2])
    if strong is not None:
        strong["efficiency_vs_weak"] = strong["value"]/(world*T*args.steps/el)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s"
        # dominant kernel: the device-load launch (k_load_mos + cap + src of one tick)
        load_ms = ms_load/max(nload, 1)
        solve_ms = ms_solve/max(nsolve, 1)
        # a trial takes part in a tick's load + solve only while it iterates (parked or finished
        # trials move no bytes): live fraction = Newton iterations per trial / ticks
        live = min(1.0, float(cnt[:, C_NUMITER].mean())/max(1.0, ticks/args.steps))
        ach = live*st.algbytes_load_per_launch/(load_ms*1e-3)/1e9 if load_ms > 0 else 0.0
        ach_solve = live*st.algbytes_solve_per_launch/(solve_ms*1e-3)/1e9 if solve_ms > 0 else 0.0
        value = world*T*args.steps/el
        e2e_val = world*T*args.steps/el_e2e
        h2d = int(mc.packed.nbytes)
        d2h = int(cnt_e.nbytes + w_e.nbytes)
        # CPU baseline on this box's cores (bounded sample) = the SAME trials the GPU just ran
        # (rank 0's trials 0..k-1 with their own parameter values), so the same run also gives
        # the parity record: counts identical per trial, waveforms within reltol/vntol
        cpu = parity = None
        try:
            cores = host_cores()
            k = min(T, max(cores*64, 1024))
            ovk = {key: v[:k] for key, v in mc.variation.items()}
            s = run_reference_sample(k, cores, ov=ovk, want=save, nstages=args.stages)
            tsum = s["t_load"] + s["t_decomp"] + s["t_solve"]
            cpu = {"value": s["trials_per_s"], "unit": "trials/s", "cores": s["cores"],
                   "kind": "reference",
                   "sample": f"trials 0..{k-1} of rank 0's batch (same parameter values) in "
                             f"{s['seconds']:.1f} s, one process per core, deck re-parsed per trial "
                             "(oracle/_ref = unmodified reference, Sparse 1.3; KLU source absent)",
                   "split_s_per_trial": {"load": s["t_load"]/k, "decomp": s["t_decomp"]/k,
                                         "solve": s["t_solve"]/k},
                   "load_fraction": s["t_load"]/tsum if tsum > 0 else None}
            parity = compare_with_reference(s["per_trial"], cnt_e, w_e)
            one = run_reference_sample(8, 1, ov={key: v[:8] for key, v in ovk.items()},
                                       nstages=args.stages)
            cpu["one_core"] = {"value": one["trials_per_s"], "unit": "trials/s", "cores": 1,
                               "sample": f"trials 0..7 in {one['seconds']:.1f} s"}
        except Exception as e:          # pragma: no cover
            cpu = cpu or {"value": None, "unit": "trials/s", "cores": 0, "kind": "reference",
                          "sample": f"unavailable: {e!r}"}
        iters = int(cnt[:, C_NUMITER].sum())
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture
        # (profiles/traffic.json: bytes per launch at the captured trial count; every per-trial
        # array scales linearly with the number of trials)
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_load_mos"]
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"])*T/tj["trials"]
            traffic_src = (f"{tj['source']}: {tj['trials']} trials captured, scaled by "
                           f"{T}/{tj['trials']} trials")
        except Exception:
            pass
        line = {
            "metric": "Monte Carlo transient trials/sec", "value": value, "unit": "trials/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3*el/args.steps, "wall_ms_per_step": 1e3*el_wall/args.steps,
            "timing": "CUDA events on the launching stream, max over ranks",
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.stages, T, world),
            "run": {"ok_trials": int(agg[0]), "ticks_per_step": ticks//args.steps,
                    "newton_iters_per_trial": iters/max(T, 1),
                    "device_load_instances_per_s": ninst*float(agg[2])*args.steps/el,
                    "device_load_instances_note": "(MOS + C + V instances) x sum over trials of "
                                                  "Newton iterations (= load calls) / device time"},
            "e2e": {"value": e2e_val, "unit": "trials/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "clocks": clocks_e2e},
            "gpu_launches": int(agg[1]),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_load_mos+k_load_cap+k_load_src (one tick)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach/peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "ms_per_launch": load_ms, "live_trial_fraction": live,
                         "solve_kernel": {"achieved": ach_solve, "ms_per_launch": solve_ms,
                                          "frac": ach_solve/peak}},
            "cpu_baseline": cpu,
            "parity": parity,
        }
        if strong is not None:
            line["strong"] = strong
        if world == 1 and not args.no_secondary:
            try:
                line["secondary"] = {"rsfq_mc": bench_rsfq(local)}
            except Exception as e:          # pragma: no cover
                line["secondary"] = {"rsfq_mc": {"error": repr(e)}}
            try:
                line["secondary"]["rsfq_tjm_mc"] = bench_rsfq(local, ntrials=1024, tjm="tjm1")
            except Exception as e:          # pragma: no cover
                line["secondary"]["rsfq_tjm_mc"] = {"error": repr(e)}
            try:
                line["secondary"]["dc_sweep"] = bench_dc_sweep(local)
            except Exception as e:          # pragma: no cover
                line["secondary"]["dc_sweep"] = {"error": repr(e)}
            try:
                line["secondary"]["nor_mc"] = bench_nor_mc(local)
            except Exception as e:          # pragma: no cover
                line["secondary"]["nor_mc"] = {"error": repr(e)}
            try:
                line["secondary"]["large_flat"] = bench_large_flat(local, time_reference=args.large_ref)
            except Exception as e:          # pragma: no cover
                line["secondary"]["large_flat"] = {"error": repr(e)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_JSON_OUT = None


def emit(line: dict):
    """The one JSON line, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries the JSON line and nothing else: keep the original fd for emit() and point
    # fd 1 at stderr, so banners that libraries write to stdout (NCCL prints its version there
    # when NCCL_DEBUG is set in the environment) cannot precede it
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trials", type=int, default=TRIALS_PER_GPU)
    ap.add_argument("--stages", type=int, default=NSTAGES)
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary (RSFQ) workload")
    ap.add_argument("--large-ref", action="store_true",
                    help="large_flat: also time the reference's trial 0 on this host (minutes)")
    ap.add_argument("--kernel-only", action="store_true",
                    help="profiling aid (ncu): device-resident steps only, no e2e / CPU legs, no JSON contract")
    args = ap.parse_args()
    if args.impl == "reference":
        return bench_reference(args)
    return bench_ours(args)


if __name__ == "__main__":
    sys.exit(main())
