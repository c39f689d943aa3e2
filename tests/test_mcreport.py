"""The Monte-Carlo verdict file (check.cc:1676-1722, 2120-2195): layout, run numbering over the
(i, j) grid, E_ITERLIM = FAIL, and that the reference's reader pattern ("[DATA] d1 d2 s1 s2
PASS|FAIL", check.cc:2334-2391) parses what is written."""
import numpy as np
import pytest

from xictools_b200.deck import C_STATUS
from xictools_b200.mcreport import grid_steps, read_report, verdicts, write_report, yield_fraction


def test_grid_of_the_reference_default():
    assert grid_steps(49) == (3, 3)            # check -m default: (2*3+1)^2 runs
    assert grid_steps(1) == (0, 0)
    s1, s2 = grid_steps(20)
    assert (2*s1 + 1)*(2*s2 + 1) >= 20


def test_report_layout_and_numbering(tmp_path):
    fail = np.zeros(49, dtype=bool)
    fail[[0, 7, 48]] = True
    alters = [[("b0.x0", "ics", 1.4e-4*(1 + 0.01*t)), ("r0.x0", "r", 40.0 + t)] for t in range(49)]
    p = write_report(tmp_path/"nor_mc.m00", "nor_mc.cir", fail, alters)
    lines = open(p).read().splitlines()
    assert lines[0].startswith("Monte Carlo Analysis from ") and lines[1].startswith("Date: ")
    assert lines[2] == "File: nor_mc.cir"
    assert lines[3] == "[DATA]  -3  -3 trial   1\t\tFAIL"
    assert lines[4] == "%-16s %-16s %s" % ("b0.x0", "ics", "0.00014")
    rep = read_report(p)
    assert len(rep) == 49
    assert [r[2] for r in rep] == list(range(1, 50))                  # run numbers
    assert (rep[0][0], rep[0][1]) == (-3, -3) and (rep[7][0], rep[7][1]) == (-3, -2)
    assert (rep[48][0], rep[48][1]) == (3, 3)
    assert [not r[3] for r in rep] == fail.tolist()
    assert rep[5][4][1] == ("r0.x0", "r", "45")
    assert abs(yield_fraction(fail) - 46/49) < 1e-15


def test_verdicts_from_counters_and_waves():
    ntr, npts = 5, 11
    cnt = np.zeros((ntr, 9), dtype=np.int32)
    cnt[3, C_STATUS] = 103                                           # E_ITERLIM: a FAIL without asking
    t = np.linspace(0, 1e-9, npts)
    w = np.zeros((ntr, npts, 2))
    w[:, :, 0] = np.linspace(0, 1, npts)[None, :]*np.array([1.0, 0.4, 1.2, 9.9, 0.9])[:, None]
    fail = verdicts(cnt, t, w, lambda tt, ww: ww[-1, 0] >= 0.8)      # final level reached?
    assert fail.tolist() == [False, True, False, True, False]
    cnt[2, C_STATUS] = 101
    with pytest.raises(RuntimeError):
        verdicts(cnt, t, w, lambda tt, ww: True)
