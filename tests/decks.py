"""Decks shared by the test modules (BASELINE.json config shapes at test size)."""
from xictools_b200.deck import Deck, ring_oscillator

COMMON = dict(gamma=0.4, phi=0.65, cgso=2e-10, cgdo=2e-10, cbd=2e-15, cbs=2e-15, tox=2e-8,
              **{"lambda": 0.02})


def inv_chain(n=4, load_res=True, **kw):
    """Pulse-driven CMOS inverter chain (examples/mos6inv.cir shape): exercises PULSE
    breakpoints, step rejection and the resistor preload."""
    d = Deck(title="inverter chain")
    d.vsource("vdd", "vdd", "0", dc=3.3)
    d.vsource("vin", "in", "0", dc=0.0, pulse=(0, 3.3, 0.2e-9, 0.1e-9, 0.1e-9, 1e-9, 2.5e-9))
    prev = "in"
    for i in range(n):
        out = f"o{i}"
        d.mosfet(f"mp{i}", out, prev, "vdd", "vdd", "pm", w=2e-6, l=1e-6)
        d.mosfet(f"mn{i}", out, prev, "0", "0", "nm", w=1e-6, l=1e-6)
        d.capacitor(f"c{i}", out, "0", 20e-15)
        prev = out
    if load_res:
        d.resistor("rl", prev, "0", 1e5)
    d.mos_model("nm", "nmos", vto=0.7, kp=1e-4, **COMMON)
    d.mos_model("pm", "pmos", vto=-0.7, kp=4e-5, **COMMON)
    d.tran(0.05e-9, 6e-9, **kw)
    return d


def rc_sin():
    """RC low-pass driven by SIN + PWL sources (source evaluators, SIN step bound)."""
    d = Deck(title="rc sin")
    d.vsource("v1", "a", "0", sin=(0.0, 1.0, 1e9, 0.1e-9, 0.0))
    d.resistor("r1", "a", "b", 1e3)
    d.capacitor("c1", "b", "0", 1e-12)
    d.vsource("v2", "c", "0", pwl=(0.0, 0.0, 1e-9, 1.0, 2e-9, 0.5, 3e-9, 0.5))
    d.resistor("r2", "c", "b", 2e3)
    d.tran(0.05e-9, 4e-9)
    return d


def rc_fn(**fn):
    """RC low-pass driven by one of the remaining analytic source functions (EXP / SFFM / AM:
    inptran.cc:2985-3082, 3194-3260, 3373-3442): evaluation, EXP breakpoints and the step bound
    each of them imposes through IFtranData::time_limit."""
    d = Deck(title="rc " + "/".join(fn))
    d.vsource("v1", "a", "0", **fn)
    d.resistor("r1", "a", "b", 1e3)
    d.capacitor("c1", "b", "0", 1e-12)
    d.tran(0.05e-9, 5e-9)
    return d


def float_cap():
    """A node that only capacitors reach: the DC matrix is singular whatever the pivot order.
    The reference ends with E_SINGULAR after trying dynamic gmin and source stepping
    (niiter.cc:303-331 reorder + retry, ckt.cc:1492-1776): 25 Newton iterations, no time point."""
    d = Deck(title="floating node")
    d.vsource("v1", "a", "0", dc=0.0, pulse=(0, 1.0, 0.1e-9, 0.05e-9, 0.05e-9, 1e-9, 3e-9))
    d.resistor("r1", "a", "b", 1e3)
    d.capacitor("c1", "b", "0", 1e-12)
    d.capacitor("c2", "b", "f", 1e-12)
    d.capacitor("c3", "f", "0", 2e-12)
    d.tran(0.05e-9, 3e-9)
    return d


def rlc_series():
    """Series RLC driven by a pulse: inductor branch equation, flux LTE, underdamped ringing."""
    d = Deck(title="series rlc")
    d.vsource("v1", "a", "0", dc=0.0, pulse=(0, 1.0, 0.1e-9, 0.05e-9, 0.05e-9, 2e-9, 5e-9))
    d.resistor("r1", "a", "b", 10.0)
    d.inductor("l1", "b", "c", 2e-9)
    d.capacitor("c1", "c", "0", 1e-12)
    d.tran(0.02e-9, 4e-9)
    return d


def lc_ladder(n=5):
    """LC ladder (lumped transmission line) between source and load resistors: nodes that
    touch only inductors and capacitors, several coupled branch equations."""
    d = Deck(title="lc ladder")
    d.vsource("vs", "in", "0", dc=0.0, pulse=(0, 1.0, 0.2e-9, 0.1e-9, 0.1e-9, 1e-9, 4e-9))
    d.resistor("rs", "in", "n0", 50.0)
    for i in range(n):
        d.inductor(f"l{i}", f"n{i}", f"n{i+1}", 2.5e-9)
        d.capacitor(f"c{i}", f"n{i+1}", "0", 1e-12)
    d.resistor("rl", f"n{n}", "0", 50.0)
    d.tran(0.02e-9, 3e-9)
    return d


def rl_isrc():
    """Current pulse into a parallel R-L with a series L-R branch (inductor to ground,
    inductor in series, current source rhs)."""
    d = Deck(title="rl isrc")
    d.isource("i1", "0", "a", dc=0.0, pulse=(0, 1e-3, 0.1e-9, 0.1e-9, 0.1e-9, 1e-9, 3e-9))
    d.resistor("r1", "a", "0", 100.0)
    d.inductor("l1", "a", "0", 5e-9)
    d.inductor("l2", "a", "b", 1e-9)
    d.resistor("r2", "b", "0", 20.0)
    d.capacitor("c1", "b", "0", 0.2e-12)
    d.tran(0.02e-9, 3e-9)
    return d


def inv_lsupply():
    """CMOS inverter chain fed through a supply inductance with a decoupling capacitor:
    MOS + inductor in one matrix (DCOP shorts the inductor)."""
    d = inv_chain(3, load_res=False)
    for e in d.elems:
        if e.kind == 5:         # MOSFETs: move the p-channel sources/bulks to the internal rail
            e.nodes = tuple("vddi" if nd == "vdd" else nd for nd in e.nodes)
    d.inductor("lvdd", "vdd", "vddi", 1e-9)
    d.capacitor("cdec", "vddi", "0", 0.5e-12)
    return d


def rlc_uic():
    """UIC start with an inductor current and a capacitor voltage as initial conditions."""
    d = Deck(title="rlc uic")
    d.resistor("r1", "a", "0", 30.0)
    d.inductor("l1", "a", "b", 3e-9, ic=2e-3)
    d.capacitor("c1", "b", "0", 1e-12, ic=0.5)
    d.tran(0.02e-9, 2e-9, uic=True)
    return d


def transformer():
    """Two coupled LC tanks (k = 0.6) plus a third winding coupled to both: mutual inductors,
    several couplings onto one inductor (ordered flux accumulation)."""
    d = Deck(title="coupled tanks")
    d.vsource("v1", "in", "0", dc=0.0, pulse=(0, 1.0, 0.1e-9, 0.05e-9, 0.05e-9, 1e-9, 3e-9))
    d.resistor("r1", "in", "a", 25.0)
    d.inductor("l1", "a", "0", 4e-9)
    d.capacitor("c1", "a", "0", 0.5e-12)
    d.inductor("l2", "b", "0", 9e-9)
    d.capacitor("c2", "b", "0", 0.3e-12)
    d.resistor("r2", "b", "0", 200.0)
    d.inductor("l3", "c", "d", 2e-9)
    d.resistor("r3", "c", "0", 50.0)
    d.resistor("r4", "d", "0", 50.0)
    d.mutual("k12", "l1", "l2", 0.6)
    d.mutual("k13", "l1", "l3", 0.3)
    d.mutual("k23", "l2", "l3", 0.2)
    d.tran(0.02e-9, 3e-9)
    return d


L6 = dict(kv=0.9, nv=0.85, kc=3.8e-5, nc=1.25, phi=0.65, gamma=0.45, gamma1=0.02, sigma=0.04,
          lambda0=0.03, lambda1=0.01, cgso=2e-10, cgdo=2e-10, cbd=2e-15, cbs=2e-15, tox=2e-8)


def ro_level6(n=5):
    """Ring oscillator with MOS level 6 (n-th power law, moseq.cc:1091) devices, in the style of
    examples/mos6inv.cir."""
    d = ring_oscillator(n)
    d.mos_models.clear()
    d.mos_model("nm", "nmos", level=6, vto=0.7, **L6)
    d.mos_model("pm", "pmos", level=6, vto=-0.7, **dict(L6, kc=1.6e-5))
    return d


def inv_level6():
    d = inv_chain(4)
    d.mos_models.clear()
    d.mos_model("nm", "nmos", level=6, vto=0.7, **L6)
    d.mos_model("pm", "pmos", level=6, vto=-0.7, **dict(L6, kc=1.6e-5))
    return d


L3 = dict(tox=2e-8, nsub=2e16, xj=2.5e-7, ld=8e-8, u0=550, vmax=1.6e5, theta=0.08, eta=0.06,
          kappa=0.6, delta=0.9, nfs=2e11, cgso=2e-10, cgdo=2e-10, cbd=2e-15, cbs=2e-15)
L3S = dict(tox=2e-8, nsub=2e16, u0=550, theta=0.05, eta=0.03, kappa=0.4, kp=9e-5,
           cgso=2e-10, cgdo=2e-10, cbd=2e-15, cbs=2e-15)


def ro_level3(n=5, card=None):
    """Ring oscillator with MOS level 3 devices (moseq.cc:693): short-channel factor, velocity
    saturation, weak inversion (nfs), channel-length modulation and punch-through branches."""
    card = L3 if card is None else card
    d = ring_oscillator(n, tstop=8e-9, tstep=0.08e-9)
    d.mos_models.clear()
    d.mos_model("nm", "nmos", level=3, vto=0.7, **card)
    d.mos_model("pm", "pmos", level=3, vto=-0.7, **dict(card, u0=card["u0"]*0.4))
    return d


def inv_level3():
    d = inv_chain(4)
    d.mos_models.clear()
    d.mos_model("nm", "nmos", level=3, vto=0.7, **L3S)
    d.mos_model("pm", "pmos", level=3, vto=-0.7, **dict(L3S, kp=3.5e-5))
    return d


L2 = dict(tox=2.5e-8, nsub=1.5e16, xj=3e-7, ld=1e-7, u0=580, ucrit=5e4, uexp=0.12, neff=2.5,
          vmax=1.2e5, nfs=1.5e11, delta=1.2, cgso=2e-10, cgdo=2e-10, cbd=2e-15, cbs=2e-15)
L2S = dict(tox=2.5e-8, kp=8e-5, gamma=0.45, phi=0.68, ucrit=4e4, uexp=0.1,
           cgso=2e-10, cgdo=2e-10, cbd=2e-15, cbs=2e-15, **{"lambda": 0.03})


def ro_level2(n=5, card=None):
    """Ring oscillator with MOS level 2 devices (moseq.cc:147): short/narrow channel threshold,
    mobility degradation, Baum velocity saturation (quartic), weak inversion, punch-through."""
    card = L2 if card is None else card
    d = ring_oscillator(n, tstop=8e-9, tstep=0.08e-9)
    d.mos_models.clear()
    d.mos_model("nm", "nmos", level=2, vto=0.7, **card)
    d.mos_model("pm", "pmos", level=2, vto=-0.7, **dict(card, **({"u0": card["u0"]*0.4} if "u0" in card else {"kp": card["kp"]*0.4})))
    return d


def inv_level2():
    d = inv_chain(4)
    d.mos_models.clear()
    d.mos_model("nm", "nmos", level=2, vto=0.7, **L2S)
    d.mos_model("pm", "pmos", level=2, vto=-0.7, **dict(L2S, kp=3.2e-5))
    return d


def ro_rd(n=3):
    """Ring oscillator whose MOSFETs have series rd/rs: internal drain/source nodes."""
    d = ring_oscillator(n)
    for m in d.mos_models.values():
        m["params"]["rd"] = 50.0
        m["params"]["rs"] = 30.0
    return d


def inv_dc(n=3):
    """CMOS inverter chain with a DC input source: the DC operating point / transfer-curve
    sweep deck (BASELINE config 5 shape: every sweep point is one trial)."""
    d = Deck(title="inverter chain dc")
    d.vsource("vdd", "vdd", "0", dc=3.3)
    d.vsource("vin", "in", "0", dc=1.2)
    prev = "in"
    for i in range(n):
        out = f"o{i}"
        d.mosfet(f"mp{i}", out, prev, "vdd", "vdd", "pm", w=2e-6, l=1e-6)
        d.mosfet(f"mn{i}", out, prev, "0", "0", "nm", w=1e-6, l=1e-6)
        d.capacitor(f"c{i}", out, "0", 20e-15)
        prev = out
    d.resistor("rl", prev, "0", 1e5)
    d.mos_model("nm", "nmos", vto=0.7, kp=1e-4, **COMMON)
    d.mos_model("pm", "pmos", vto=-0.7, kp=4e-5, **COMMON)
    d.tran(0.05e-9, 1e-9)
    return d


def op_decks():
    from xictools_b200.workloads import opamp
    return {"op_opamp": opamp(), "op_opamp_unbalanced": opamp(vinp=1.68),
            "op_inv_dc": inv_dc(3), "op_inv4": inv_chain(4), "op_ro5": ring_oscillator(5),
            "op_ro3_rd": ro_rd(3), "op_inv_dc_direct": inv_dc(3).options(gminsteps=-1, srcsteps=-1),
            "op_inv_dc_gminfirst": inv_dc(3).options(gminfirst=1),
            "op_inv7_nogmin": inv_chain(7).options(gminsteps=-1),
            "op_inv_lsupply": inv_lsupply(), "op_inv4_level6": inv_level6(), "op_inv4_level3": inv_level3(),
            "op_ro5_level3": ro_level3(5), "op_inv4_level2": inv_level2(),
            "op_ro5_level2": ro_level2(5),
            # option dcmu: the operating point mixes in some of the last solution (niiter.cc:480-497)
            "op_opamp_dcmu": opamp().options(dcmu=0.3), "op_inv7_dcmu": inv_chain(7).options(dcmu=0.4, gminsteps=-1)}


def variants():
    out = {
        "rlc_series": rlc_series(),
        "lc_ladder": lc_ladder(5),
        "rl_isrc": rl_isrc(),
        "inv_lsupply": inv_lsupply(),
        "rlc_uic": rlc_uic(),
        "transformer": transformer(),
        "ro5_level6": ro_level6(5),
        "ro5_level3": ro_level3(5),
        "ro5_level2": ro_level2(5),
        "ro3_level2_simple": ro_level2(3, L2S),
        "inv4_level2": inv_level2(),
        "ro3_level3_simple": ro_level3(3, L3S),
        "inv4_level3": inv_level3(),
        # transient iteration limit hit: non-convergence -> delta/8 cuts (tranan.cc:744-762)
        "ro5_itl4_4": ring_oscillator(5).options(itl4=4),
        "inv4_itl4_4": inv_chain(4).options(itl4=4),
        "inv4_level6": inv_level6(),
        "ro3": ring_oscillator(3),
        "ro5": ring_oscillator(5),
        "ro17": ring_oscillator(17),
        "inv4": inv_chain(4),
        "inv7": inv_chain(7),
        "ro7_long": ring_oscillator(7, tstep=0.2e-9, tstop=20e-9),
        "ro9_one_ic": _one_ic(),
        "ro5_ord1": ring_oscillator(5).options(maxord=1),
        "ro5_trapcheck": ring_oscillator(5).options(trapcheck=1),
        "ro5_gear3": ring_oscillator(5).options(method=1, maxord=3),
        "ro5_nobypass": ring_oscillator(5).options(bypass=0),
        "ro5_reltol": ring_oscillator(5).options(reltol=1e-4),
        "ro5_interp2": ring_oscillator(5).options(interplev=2),
        "ro3_rd": ro_rd(3),
        "rc_sin": rc_sin(),
        "rc_exp": rc_fn(exp=(0.0, 1.0, 0.5e-9, 0.3e-9, 2.5e-9, 0.6e-9)),
        "rc_exp_defaults": rc_fn(exp=(0.2, 1.0)),
        "rc_sffm": rc_fn(sffm=(0.5, 0.4, 1.0e9, 2.0, 0.2e9)),
        "rc_am": rc_fn(am=(0.8, 1.0, 0.3e9, 2.0e9, 0.4e-9)),
        "float_cap": float_cap(),
        # PULSE / GPULSE bit patterns (inptran.cc:285-520), expanded by deck.expand_bit_pattern
        "rc_pulse_bits": rc_fn(pulse=(0, 1, 0.1e-9, 0.05e-9, 0.05e-9, 0.15e-9, 0.5e-9), pattern="b1011 rb=2 r=-1"),
        "rc_pulse_prbs": rc_fn(pulse=(0.2, 1, 0.05e-9, 0.05e-9, 0.05e-9, 0.1e-9, 0.4e-9), pattern="bprbs7 r=-1"),
        "rc_pulse_two": rc_fn(pulse=(0, 1, 0.1e-9, 0.04e-9, 0.04e-9, 0.1e-9, 0.3e-9), pattern="b101 r=1 b0011 rb=3 r=2"),
        "rc_gpulse_bits": rc_fn(gpulse=(0, 1, 0.2e-9, 0.06e-9, 0.4e-9), pattern="b110 r=2"),
        "rc_gpulse_prbs": rc_fn(gpulse=(0, 1, 0.2e-9, 0.05e-9, 0.3e-9), pattern="bprbs6 rb=11 r=-1"),
        "ro5_dcmu": ring_oscillator(5).options(dcmu=0.35),
        # GPULSE (inptran.cc:1567-1827): FWHM width + two more pulse delays; repeating; old-style width
        "rc_gpulse": rc_fn(gpulse=(0.1, 1.0, 1.0e-9, 0.3e-9, 0, 2.2e-9, 3.1e-9)),
        "rc_gpulse_rpt": rc_fn(gpulse=(0.0, 0.8, 0.6e-9, -0.12e-9, 1.1e-9)),
        # SPULSE (inptran.cc:2708-2870): raised cosine after a delay, plain and damped
        "rc_spulse": rc_fn(spulse=(0.1, 1.0, 1.5e-9, 0.7e-9)),
        "rc_spulse_theta": rc_fn(spulse=(0.0, 1.0, 0.4e-9, 0.3e-9, 1.2e9)),
    }
    out.update(jj_variants())
    return out


def _one_ic():
    d = ring_oscillator(9)
    d.ics.clear()
    d.ic(n0=3.3)
    return d


from xictools_b200.workloads import JJ_CARD, jj_single, jtl, opamp      # noqa: E402,F401


def squid_coupled(uic=False):
    """A two-junction SQUID loop biased through a resistor from a DC rail, flux-coupled (K) to an
    input coil that a current pulse drives: mutual-inductor terms of the phase-mode DC operating
    point (mutload.cc:75-87) and of the transient (flux through the coupled pair)."""
    d = Deck(title="coupled squid")
    d.jj_model("jj1", **JJ_CARD)
    d.vsource("vdd", "vb", "0", dc=10e-3)
    d.resistor("rb", "vb", "m", 10e-3/0.35e-3)
    d.inductor("l1", "m", "a", 4.0e-12)
    d.inductor("l2", "m", "b", 4.0e-12)
    d.jj("b1", "a", "0", "jj1", ics=0.25e-3)
    d.jj("b2", "b", "0", "jj1", ics=0.26e-3)
    d.resistor("rs1", "a", "0", 3.0)
    d.resistor("rs2", "b", "0", 3.0)
    d.isource("iin", "0", "c", pulse=(0, 0.5e-3, 20e-12, 5e-12, 5e-12, 30e-12, 200e-12))
    d.inductor("lin", "c", "0", 6.0e-12)
    d.mutual("k1", "lin", "l1", 0.6)
    d.mutual("k2", "lin", "l2", -0.4)
    d.tran(0.5e-12, 100e-12, uic=uic)
    return d


def jj_variants():
    return {
        "jj_single": jj_single(),
        "jj_single_rtype2": jj_single(rtype=2),
        "jj_single_rtype3": jj_single(rtype=3),
        "jj_single_cct0": jj_single(cct=0),
        "jj_single_pi": jj_single(pijj=1),
        "jtl6": jtl(6),
        "jtl6_phase": jtl(6, phase_nodes=True),
        "jtl6_vshunt": jtl(6, vshunt=0.6857e-3),
        "jtl6_gear": jtl(6).options(method=1),
        "jtl12": jtl(12),
        # tunnel-junction model (JJ level 3); set-up values from xictools_b200/data/tjm_setup.json
        "jtl6_tjm": jtl(6, tjm="tjm1"),
        "jtl6_tjm_vshunt": jtl(6, tjm="tjm1_vshunt"),
        "jtl6_tjm2_phase": jtl(6, tjm="tjm2", phase_nodes=True),
        # no `uic`: the transient starts from a phase-mode DC operating point (jjload.cc:154-250)
        "jtl6_dc": jtl(6, dc_bias=True),
        "jtl6_dc_phase": jtl(6, dc_bias=True, phase_nodes=True),
        "jtl6_dc_vshunt": jtl(6, dc_bias=True, vshunt=0.6857e-3),
        "jtl12_dc": jtl(12, dc_bias=True),
        "jtl6_dc_tjm": jtl(6, dc_bias=True, tjm="tjm1"),
        # option jjaccel (niiter.cc:456,535-539, tranan.cc:590,787-792): the junctions' own test ends
        # the iteration, no LTE
        "jtl6_accel": jtl(6).options(jjaccel=1),
        "jtl6_dc_accel": jtl(6, dc_bias=True).options(jjaccel=1),
        "jtl12_vshunt_accel": jtl(12, vshunt=0.6857e-3).options(jjaccel=1),
        "jtl6_tjm_accel": jtl(6, tjm="tjm1").options(jjaccel=1),
        "squid_coupled": squid_coupled(),
        "squid_coupled_uic": squid_coupled(uic=True),
        # driven by a single-flux-quantum GPULSE voltage source
        "jtl6_dc_sfq": jtl(6, dc_bias=True, sfq_in=True),
    }
