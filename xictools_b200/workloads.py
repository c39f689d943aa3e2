"""Deck builders of the benchmark workloads (BASELINE.json configs) beside the ring oscillator
of deck.py: the Josephson transmission line of config 3 and the opamp of config 5.  Plain deck
construction, no numerics."""
from __future__ import annotations

from .deck import Deck

JJ_CARD = dict(rtype=1, cct=1, vg=2.8e-3, delv=0.1e-3, icrit=1e-3, rn=1.7, rsub=17.0, cap=0.7e-12)


def jj_single(**card):
    """One shunted junction driven over its critical current by a current ramp: the junction
    switches and runs (phase slips of 2 pi, step bound tsfact*Phi0/|vj| from JJdev::accept)."""
    d = Deck(title="single jj")
    d.isource("i1", "0", "1", pulse=(0, 1.5e-3, 5e-12, 10e-12, 10e-12, 60e-12, 200e-12))
    d.jj("b1", "1", "0", "jj1")
    d.resistor("r1", "1", "0", 1.0)
    d.jj_model("jj1", **dict(JJ_CARD, **card))
    d.tran(0.5e-12, 100e-12, uic=True)
    return d


def jtl(n=6, phase_nodes=False, vshunt=None, ics=0.25e-3, tjm=None, dc_bias=False, sfq_in=False, **card):
    """Josephson transmission line (the `jtl` subcircuit of examples/JJexamples/nor_mc.cir
    flattened and repeated): biased junctions coupled by inductors, an SFQ pulse launched at
    the input propagates down the line.  Bias currents ramp up from zero (uic start)."""
    d = Deck(title=f"jtl{n}")
    if tjm:
        # tunnel-junction model (JJ level 3): `tjm` names a set-up card of data/tjm_setup.json
        d.tjm_model("jj1", tjm)
    else:
        mc = dict(JJ_CARD, **card)
        if vshunt is not None:
            mc["vshunt"] = vshunt
        d.jj_model("jj1", **mc)
    if sfq_in:
        # a voltage pulse holding one flux quantum (GPULSE with no amplitude: inptran.cc:1695-1712)
        d.vsource("vin", "in", "0", gpulse=(0, 0, 40e-12, 2e-12))
    else:
        d.isource("iin", "0", "in", pulse=(0, 0.6e-3, 40e-12, 3e-12, 3e-12, 2e-12, 200e-12))
    d.inductor("lin", "in", "a0", 2.0e-12)
    for i in range(n):
        a = f"a{i}"
        kw = {"phase": f"ph{i}"} if phase_nodes else {}
        (d.tjm if tjm else d.jj)(f"b{i}", a, "0", "jj1", ics=ics*(1.0 + 0.05*(i % 3)), **kw)
        if vshunt is None and not (tjm and "vshunt" in tjm):
            d.resistor(f"rs{i}", a, "0", 3.0)
        if dc_bias:
            # biased from a 10 mV rail through resistors, as the cells of nor_mc.cir are: the run
            # starts from the phase-mode DC operating point (no `uic`)
            d.resistor(f"rb{i}", "vb", a, 10e-3/(0.7*ics))
        else:
            d.isource(f"ib{i}", "0", a, pulse=(0, 0.7*ics, 0.0, 20e-12, 1e-12, 1.0, 2.0))
        if i + 1 < n:
            d.inductor(f"l{i}", a, f"a{i+1}", 4.0e-12)
    d.resistor("rl", f"a{n-1}", "0", 5.0)
    if dc_bias:
        d.vsource("vdd", "vb", "0", dc=10e-3)
    d.tran(0.5e-12, 120e-12, uic=not dc_bias)
    return d




def opamp(vinp: float = 1.65, vinm: float = 1.65, vdd: float = 3.3):
    """Two-stage CMOS opamp (MOS level 1): PMOS-mirror-loaded NMOS differential pair, common
    source second stage with Miller capacitor, resistor-biased current mirrors -- the shape of
    examples/bsim465-benchmarks `opamp.b4` with level-1 cards.  Config 5 sweeps its inputs."""
    d = Deck(title="two-stage opamp")
    d.vsource("vdd", "vdd", "0", dc=vdd)
    d.vsource("vinp", "inp", "0", dc=vinp)
    d.vsource("vinm", "inm", "0", dc=vinm)
    com = dict(gamma=0.45, phi=0.7, cgso=2e-10, cgdo=2e-10, cbd=5e-15, cbs=5e-15, tox=2e-8,
               **{"lambda": 0.04})
    d.mos_model("nm", "nmos", vto=0.7, kp=1.1e-4, **com)
    d.mos_model("pm", "pmos", vto=-0.8, kp=4e-5, **com)
    # bias: resistor into a diode-connected NMOS, mirrored to the tail and the output stage
    d.resistor("rb", "vdd", "nb", 60e3)
    d.mosfet("m8", "nb", "nb", "0", "0", "nm", w=10e-6, l=2e-6)
    d.mosfet("m5", "tail", "nb", "0", "0", "nm", w=20e-6, l=2e-6)
    # differential pair + PMOS mirror
    d.mosfet("m1", "x", "inm", "tail", "0", "nm", w=30e-6, l=2e-6)
    d.mosfet("m2", "y", "inp", "tail", "0", "nm", w=30e-6, l=2e-6)
    d.mosfet("m3", "x", "x", "vdd", "vdd", "pm", w=30e-6, l=2e-6)
    d.mosfet("m4", "y", "x", "vdd", "vdd", "pm", w=30e-6, l=2e-6)
    # second stage
    d.mosfet("m6", "out", "y", "vdd", "vdd", "pm", w=120e-6, l=2e-6)
    d.mosfet("m7", "out", "nb", "0", "0", "nm", w=40e-6, l=2e-6)
    d.capacitor("cc", "y", "out", 1e-12)
    d.capacitor("cl", "out", "0", 2e-12)
    d.resistor("rl", "out", "0", 1e6)
    d.tran(1e-9, 20e-9)
    return d
