"""The Monte-Carlo verdict product: what `check -m` leaves behind.

The reference runs a deck's `.control` block after every trial; the block sets `checkFAIL`, and
sCHECKprms::trial writes one line per trial into `<deck>.mNN` (wrspice/src/fte/check.cc):

  df_open  :2120-2195   "Monte Carlo Analysis from <program>-<version>" / "Date: ..." / "File: ..."
  trial    :1676-1722   "[DATA] %3d %3d trial %3d\\t\\tPASS|FAIL", then sFtCirc::printAlter: one
                        "%-16s %-16s %s" line per value the `.exec` block set for that trial
                        (fte/device.cc:2400-2420)
  numbering:1677        trials walk a (2*step1+1) x (2*step2+1) grid, i = -step1..step1 fastest;
                        run number = (j + step2)*(2*step1 + 1) + i + step1 + 1
  verdicts :1702-1710   a trial that ends with E_ITERLIM counts as FAIL; any other error stops
                        the analysis (ch_nogo)

In the batched flow the trials come back together (counters + waveforms); `verdicts` applies the
host's pass/fail predicate to each, `write_report` writes the file the reference's own reader
(processReturn, check.cc:2334-2391: "[DATA] d1 d2 s1 s2 PASS|FAIL") and `mplot` consume.
No simulation arithmetic here.
"""
from __future__ import annotations

import re
import time as _time

import numpy as np

from .deck import C_STATUS

E_ITERLIM = 103          # B200_E_ITERLIM (include/b200spice.h) == the reference's E_ITERLIM (errors.h)


def grid_steps(ntrials: int) -> tuple:
    """(step1, step2) of the smallest square-ish grid that holds `ntrials` runs; the reference's
    default is 3, 3 = 49 trials (check.cc:601-606)."""
    s1 = 0
    while (2*s1 + 1)**2 < ntrials:
        s1 += 1
    s2 = s1
    while s2 > 0 and (2*s1 + 1)*(2*s2 - 1) >= ntrials:
        s2 -= 1
    return s1, s2


def verdicts(counters, times, waves, predicate) -> np.ndarray:
    """fail[t] for every trial: predicate(times[:npts], waves[t]) -> True = PASS; a trial that
    ended with E_ITERLIM is a FAIL without asking (check.cc:1702-1707); another error raises,
    as it stops the reference's analysis."""
    ntr = counters.shape[0]
    fail = np.zeros(ntr, dtype=bool)
    for t in range(ntr):
        st = int(counters[t, C_STATUS])
        if st == E_ITERLIM:
            fail[t] = True
        elif st != 0:
            raise RuntimeError(f"trial {t} ended with error {st}: the analysis stops (check.cc:1725)")
        else:
            fail[t] = not bool(predicate(times, waves[t]))
    return fail


def write_report(path, deckfile, fail, alters=None, step1=None, step2=None, program="b200spice", version="0.1"):
    """alters[t] = [(device, parameter, value), ...] of trial t (what `let @dev[param] = value`
    set), printed under its [DATA] line as sFtCirc::printAlter does."""
    fail = np.asarray(fail, dtype=bool)
    n = len(fail)
    if step1 is None or step2 is None:
        step1, step2 = grid_steps(n)
    n1 = 2*step1 + 1
    assert n1*(2*step2 + 1) >= n
    with open(path, "w") as fp:
        fp.write(f"Monte Carlo Analysis from {program}-{version}\n")
        fp.write("Date: %s\n" % _time.strftime("%a %b %d %H:%M:%S %Y"))
        fp.write(f"File: {deckfile}\n")
        for t in range(n):
            j, i = divmod(t, n1)
            i -= step1
            j -= step2
            num = (j + step2)*n1 + i + step1 + 1
            fp.write("[DATA] %3d %3d trial %3d\t\t%s\n" % (i, j, num, "FAIL" if fail[t] else "PASS"))
            for dev, prm, val in (alters[t] if alters is not None else ()):
                fp.write("%-16s %-16s %s\n" % (dev, prm, f"{val:.15g}" if not isinstance(val, str) else val))
    return path


_DATA = re.compile(r"\[DATA\]\s+(-?\d+)\s+(-?\d+)\s+(\S+)\s+(\S+)\s+(PASS|FAIL)")


def read_report(path):
    """[(i, j, run, passed, [(device, parameter, value-string)])] -- the fields processReturn
    parses, plus the alter lines that follow each trial."""
    out = []
    with open(path) as fp:
        for ln in fp:
            m = _DATA.match(ln)
            if m:
                out.append([int(m.group(1)), int(m.group(2)), int(m.group(4)), m.group(5) == "PASS", []])
            elif out and len(ln.split()) == 3 and not ln.startswith(("Date:", "File:")):
                out[-1][4].append(tuple(ln.split()))
    return [tuple(x) for x in out]


def yield_fraction(fail) -> float:
    fail = np.asarray(fail, dtype=bool)
    return float((~fail).sum())/max(1, len(fail))
