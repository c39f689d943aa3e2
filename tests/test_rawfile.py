"""Rawfile output: the writer's layout is the reference's cRawOut
(fte/rawfile.cc:153-388); the parser used to check it also reads the rawfiles the reference tree
itself holds (examples/bsim430-benchmarks.tar.gz, written by Berkeley spice3), which pins the
format both ways where /root/reference is present."""
import io
import os
import tarfile

import numpy as np
import pytest

from xictools_b200.rawfile import read_rawfile, vector_name_unit, write_plot, write_rawfile

REF_TAR = "/root/reference/wrspice/examples/bsim430-benchmarks.tar.gz"


def test_header_lines_are_the_references(tmp_path):
    buf = io.BytesIO()
    write_plot(buf, "a deck", "Transient Analysis", ["out", "vdd#branch", "m1#internal#drain"],
               [0.0, 1e-9], [[1.0, -2e-3, 0.5], [1.5, -1e-3, 0.25]], date="Thu May  8 16:35:13  2003")
    lines = buf.getvalue().decode().splitlines()
    assert lines[:7] == ["Title: a deck", "Date: Thu May  8 16:35:13  2003", "Plotname: Transient Analysis",
                         "Flags: real", "No. Variables: 4", "No. Points: 2       ",
                         "Command: version b200spice"]
    assert lines[7:13] == ["Variables:", " 0 time S", " 1 v(out) V", " 2 vdd#branch A",
                           " 3 m1#internal#drain V", "Values:"]
    assert lines[13] == " 0\t0.000000000000000e+00" and lines[14] == "\t1.000000000000000e+00"
    assert vector_name_unit("n3") == ("v(n3)", "V")


@pytest.mark.parametrize("binary", [False, True])
def test_round_trip_of_a_batch(tmp_path, binary):
    rng = np.random.default_rng(1)
    ntr, npts, names = 3, 11, ["a", "b", "v1#branch"]
    times = np.linspace(0, 1e-9, npts)
    waves = rng.standard_normal((ntr, npts, len(names)))
    cnt = np.zeros((ntr, 9), dtype=np.int32)
    cnt[:, 8] = [npts, 7, npts]                  # C_NPOINTS: trial 1 stopped early
    path = write_rawfile(str(tmp_path / "mc.raw"), "mc run", names, times, waves, cnt, binary=binary)
    plots = read_rawfile(path)
    assert [p["plotname"] for p in plots] == [f"Transient Analysis trial {t}" for t in range(ntr)]
    assert plots[0]["names"] == ["time", "v(a)", "v(b)", "v1#branch"] and plots[0]["units"] == ["S", "V", "V", "A"]
    assert [len(p["data"]) for p in plots] == [npts, 7, npts]
    for t, p in enumerate(plots):
        n = len(p["data"])
        assert np.array_equal(p["data"][:, 0], times[:n])
        if binary:
            assert np.array_equal(p["data"][:, 1:], waves[t, :n])
        else:
            assert np.allclose(p["data"][:, 1:], waves[t, :n], rtol=1e-15, atol=0)


@pytest.mark.skipif(not os.path.exists(REF_TAR), reason="reference tree absent")
def test_parser_reads_the_rawfiles_the_reference_holds(tmp_path):
    with tarfile.open(REF_TAR) as tf:
        members = [m for m in tf.getmembers() if m.name.endswith(".raw") and "/tran/" in m.name]
        assert members
        for m in members[:4]:
            data = tf.extractfile(m).read()
            p = tmp_path / os.path.basename(m.name)
            p.write_bytes(data)
            plots = read_rawfile(str(p))
            assert len(plots) == 1 and plots[0]["names"][0] == "time" and plots[0]["units"][0] == "S"
            d = plots[0]["data"]
            assert d.shape[1] == len(plots[0]["names"]) and (np.diff(d[:, 0]) > 0).all()


@pytest.mark.gpu
def test_engine_waves_to_rawfile(tmp_path):
    from xictools_b200 import engine
    from xictools_b200.deck import mc_variation, ring_oscillator
    d = ring_oscillator(5)
    b = engine.Batch(d, 4)
    b.set_params(mc_variation(d, 4, seed=2))
    b.run_tran()
    t, w = b.waves()
    path = write_rawfile(str(tmp_path / "ro5.raw"), d.title, b.names(), t, w, b.counters(), binary=True)
    plots = read_rawfile(path)
    assert len(plots) == 4 and "vdd#branch" in plots[0]["names"]
    k = plots[2]["names"].index("v(n0)") - 1
    npt = len(plots[2]["data"])
    assert np.array_equal(plots[2]["data"][:, 1 + k], w[2, :npt, b.names().index("n0")])
    b.close()
