"""SPICE rawfile output of batch results, so WRspice's own
post-processing (`load`, plots, `.measure` on loaded vectors) can consume a batched run unchanged.

Layout follows the reference's writer, cRawOut (wrspice/src/fte/rawfile.cc):
  file_head   :153-262   Title / Date / Plotname / Flags / No. Variables / No. Points
                         [/ Dimensions] / Command: version
  file_vars   :264-305   "Variables:" then one line per vector " <i> <name> <unit>", the scale
                         (time) first; then "Values:" (ASCII) or "Binary:"
  file_points :308-388   ASCII: " <index>\\t<value>\\n" then "\\t<value>\\n" per further vector, values
                         as %.*e with `rawfileprec` digits (default 15); binary: doubles in point-major
                         order, native endianness
A Monte-Carlo / sweep run writes one plot per trial, one after the other in the same file -- what
`check` leaves when it is told to keep every trial's data, and what `load` reads back as a list
of plots.  Vector names and units follow sCKT's output naming (ckt.cc:680-708): node voltages
"v(name)" in V, branch currents "name#branch" in A, the scale "time" in S.

This module writes and (for the tests) re-reads the format; it holds no simulation arithmetic.
"""
from __future__ import annotations

import struct
import time as _time

import numpy as np

DEF_PREC = 15            # rawfileprec default (fte/rawfile.cc: ro_prec)


def vector_name_unit(unknown: str) -> tuple:
    """Reference-style vector name and unit string of one unknown of the MNA system."""
    if unknown.endswith("#branch"):
        return unknown, "A"
    if "#" in unknown:                       # internal nodes keep their composed name
        return unknown, "V"
    return f"v({unknown})", "V"


def _head(fp, title, date, plotname, nvars, npoints, version):
    fp.write(f"Title: {title}\n".encode())
    fp.write(f"Date: {date}\n".encode())
    fp.write(f"Plotname: {plotname}\n".encode())
    fp.write(b"Flags: real\n")
    fp.write(f"No. Variables: {nvars}\n".encode())
    fp.write(f"No. Points: {npoints:<8d}\n".encode())
    fp.write(f"Command: version {version}\n".encode())


def write_plot(fp, title, plotname, names, times, values, *, binary=False, prec=DEF_PREC,
               scale=("time", "S"), date=None, version="b200spice"):
    """One plot: times[npoints], values[npoints][len(names)] (unknown names as the engine
    reports them)."""
    times = np.asarray(times, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    npts = len(times)
    assert values.shape == (npts, len(names))
    _head(fp, title, date or _time.strftime("%a %b %d %H:%M:%S %Y"), plotname, len(names) + 1, npts,
          version)
    fp.write(b"Variables:\n")
    fp.write(f" 0 {scale[0]} {scale[1]}\n".encode())
    for i, nm in enumerate(names):
        vn, unit = vector_name_unit(nm)
        fp.write(f" {i+1} {vn} {unit}\n".encode())
    if binary:
        fp.write(b"Binary:\n")
        rows = np.concatenate([times[:, None], values], axis=1)
        fp.write(np.ascontiguousarray(rows).tobytes())
    else:
        fp.write(b"Values:\n")
        for p in range(npts):
            fp.write(f" {p}\t{times[p]:.{prec}e}\n".encode())
            for v in values[p]:
                fp.write(f"\t{v:.{prec}e}\n".encode())


def write_rawfile(path, title, names, times, waves, counters=None, *, binary=False, prec=DEF_PREC,
                  plotname="Transient Analysis", scale=("time", "S")):
    """waves[ntrials][npoints][len(names)] (engine.Batch.waves / MonteCarlo.waves): one plot per
    trial; with `counters` (the C_NPOINTS column) each plot holds only the points that trial
    wrote (a failed trial's plot is short, as the reference's would be)."""
    from .deck import C_NPOINTS
    waves = np.asarray(waves)
    with open(path, "wb") as fp:
        for t in range(waves.shape[0]):
            npt = int(counters[t, C_NPOINTS]) if counters is not None else waves.shape[1]
            write_plot(fp, title, f"{plotname} trial {t}" if waves.shape[0] > 1 else plotname, names,
                       times[:npt], waves[t, :npt], binary=binary, prec=prec, scale=scale)
    return path


def read_rawfile(path):
    """Parser of the same format (tests): returns [dict(title, plotname, names, units, data)]
    with data[npoints][nvars] including the scale as column 0."""
    plots = []
    with open(path, "rb") as fp:
        buf = fp.read()
    pos = 0

    def line():
        nonlocal pos
        e = buf.index(b"\n", pos)
        s = buf[pos:e].decode()
        pos = e + 1
        return s
    while pos < len(buf):
        hd = {}
        while True:
            ln = line()
            if ln.lower().startswith("variables:"):          # the reader is case-blind (rawfile.cc:560-700)
                break
            k, _, v = ln.partition(":")
            hd[k.strip().lower()] = v.strip()
        nv, npts = int(hd["no. variables"]), int(hd["no. points"])
        names, units = [], []
        for _ in range(nv):
            tok = line().split()
            names.append(tok[1]); units.append(tok[2] if len(tok) > 2 else "")      # the unit may be absent
        kind = line().strip()
        data = np.zeros((npts, nv))
        if kind == "Binary:":
            n = npts*nv
            data = np.array(struct.unpack(f"{n}d", buf[pos:pos + 8*n])).reshape(npts, nv)
            pos += 8*n
        else:
            for p in range(npts):
                tok = line().split()
                assert int(tok[0]) == p
                data[p, 0] = float(tok[1])
                for k in range(1, nv):
                    data[p, k] = float(line())
        plots.append(dict(title=hd["title"], plotname=hd["plotname"], flags=hd["flags"], names=names,
                          units=units, data=data))
    return plots
