"""Host-side deck description for the batched engine.

Mirrors what WRspice's front end hands to the numerical core: a flat list of device cards
plus .model/.ic/.options/.tran, and reproduces the *orderings* the reference derives from the
deck text, because they decide floating-point summation order in the matrix and RHS:

  * node numbers: order of first appearance while parsing (src/inp/inpdeck.cc pass2,
    ckt/symtab.cc term_insert);
  * load order: model-TYPE list in order of first model creation (include/circuit.h:1199-1223
    sCKTmodHead::insert appends a new type), models of one type newest-first
    (ckt.cc:1418-1421 newModl prepends), instances of one model newest-first
    (ckt.cc:1388-1389 newInst prepends), mutual inductors before everything
    (ckt.cc:991-1005).

`Deck.to_spice()` renders the same circuit as SPICE text for the compiled reference
(oracle/_ref, tests only); `Deck.to_c()` builds the C-ABI `b200_deck` (include/b200spice.h).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

NAN = float("nan")

DEV_RES, DEV_CAP, DEV_IND, DEV_MUT, DEV_SRC, DEV_MOS, DEV_JJ, DEV_TJM = range(8)
SRC_V, SRC_I = 0, 1
FN_NONE, FN_PULSE, FN_PWL, FN_SIN, FN_EXP, FN_SFFM, FN_AM, FN_GAUSS, FN_GPULSE, FN_SPULSE = range(10)

# parameter ids (include/b200spice.h)
P_MOS_VTO, P_MOS_KP, P_MOS_L, P_MOS_W, P_MOS_GAMMA, P_MOS_LAMBDA, P_MOS_PHI = range(1, 8)
P_CAP_C, P_RES_R, P_IND_L, P_SRC_DC, P_JJ_ICS, P_TJM_ICS = 32, 48, 56, 64, 80, 88

# counters
(C_STATUS, C_ACCEPTED, C_REJECTED, C_NUMITER, C_TIMEPTS, C_TRANITER, C_ITERCUT, C_TRAPCUT,
 C_NPOINTS, NCOUNTERS) = range(10)


class COptions(C.Structure):
    _fields_ = ([(n, C.c_double) for n in (
        "abstol", "reltol", "vntol", "chgtol", "trtol", "gmin", "gmax", "pivrel", "pivtol",
        "temp", "tnom", "xmu", "dcmu", "trapratio", "delmin", "minbreak", "resmin",
        "defl", "defw", "defad", "defas")] +
        [(n, C.c_int) for n in (
            "bypass", "maxord", "itl1", "itl2", "itl4", "itl_gmin", "itl_src",
            "gminsteps", "srcsteps", "gminfirst", "method", "trapcheck", "spice3",
            "oldsteplim", "oldlimit", "forcegmin", "interplev", "jjaccel", "nojjtp",
            "noopiter", "steptype")])


class CRes(C.Structure):
    _fields_ = [("pos", C.c_int), ("neg", C.c_int), ("r", C.c_double), ("m", C.c_double),
                ("tc1", C.c_double), ("tc2", C.c_double), ("temp", C.c_double)]


class CCap(C.Structure):
    _fields_ = [("pos", C.c_int), ("neg", C.c_int), ("c", C.c_double), ("m", C.c_double),
                ("ic", C.c_double)]


class CInd(C.Structure):
    _fields_ = [("pos", C.c_int), ("neg", C.c_int), ("l", C.c_double), ("m", C.c_double),
                ("ic", C.c_double)]


class CMut(C.Structure):
    _fields_ = [("ind1", C.c_int), ("ind2", C.c_int), ("k", C.c_double)]


class CSrc(C.Structure):
    _fields_ = [("kind", C.c_int), ("pos", C.c_int), ("neg", C.c_int), ("dc", C.c_double),
                ("fn", C.c_int), ("ncoef", C.c_int), ("coef", C.POINTER(C.c_double))]


MOS_MODEL_DBL = ("vto", "kp", "gamma", "phi", "lambda_", "rd", "rs", "cbd", "cbs", "is_", "pb",
                 "cgso", "cgdo", "cgbo", "rsh", "cj", "mj", "cjsw", "mjsw", "js", "tox", "ld",
                 "u0", "fc", "nsub", "tpg", "nss", "tnom", "nfs", "delta", "uexp", "ucrit",
                 "vmax", "xj", "neff", "eta", "theta", "kappa",
                 "kv", "nv", "kc", "nc", "lamda0", "lamda1", "sigma", "gamma1")


class CMosModel(C.Structure):
    _fields_ = [("type", C.c_int), ("level", C.c_int)] + [(n, C.c_double) for n in MOS_MODEL_DBL]


class CMos(C.Structure):
    _fields_ = ([("d", C.c_int), ("g", C.c_int), ("s", C.c_int), ("b", C.c_int),
                 ("model", C.c_int)] +
                [(n, C.c_double) for n in ("l", "w", "ad", "as_", "pd", "ps", "nrd", "nrs", "m",
                                           "temp", "icvds", "icvgs", "icvbs")] +
                [("off", C.c_int)])


JJ_MODEL_DBL = ("criti", "vg", "delv", "r0", "rn", "cap", "icfactor", "cmu", "gmu", "vshunt",
                "vdpbak", "tsaccl", "tsfact")


class CJJModel(C.Structure):
    _fields_ = [("rtype", C.c_int), ("ictype", C.c_int), ("pi", C.c_int)] + \
               [(n, C.c_double) for n in JJ_MODEL_DBL]


class CJJ(C.Structure):
    _fields_ = [("pos", C.c_int), ("neg", C.c_int), ("phase_node", C.c_int), ("model", C.c_int),
                ("area", C.c_double), ("ics", C.c_double), ("vshunt", C.c_double),
                ("vg", C.c_double), ("tcf", C.c_double), ("icv", C.c_double), ("icphi", C.c_double)]


class CTJMModel(C.Structure):
    _fields_ = [("rtype", C.c_int), ("ictype", C.c_int), ("nterms", C.c_int)] + \
               [(n, C.c_double) for n in ("criti", "cap", "cmu", "gmu", "r0", "rsint", "vshunt", "tsaccl",
                                          "tsfact", "vgnom")]


class CTJM(C.Structure):
    _fields_ = [("pos", C.c_int), ("neg", C.c_int), ("phase_node", C.c_int), ("model", C.c_int),
                ("area", C.c_double), ("ics", C.c_double), ("vshunt", C.c_double), ("tcf", C.c_double),
                ("vdpbak", C.c_double), ("omegaJ", C.c_double), ("kgap", C.c_double),
                ("kgap_rejpt", C.c_double), ("alphaN", C.c_double), ("coef", C.POINTER(C.c_double)),
                ("icv", C.c_double), ("icphi", C.c_double)]


def tjm_setup(name: str) -> dict:
    """A level-3 (TJM) model card as the reference's setup leaves it: xictools_b200/data/
    tjm_setup.json, generated from the compiled reference by tools/make_tjm_fixture.py."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tjm_setup.json")
    return json.load(open(path))[name]


PHI0 = 2.067833848e-15              # wrsCONSTphi0 (include/circuit.h:123-124)
PHI0_2PI = PHI0/(2*3.14159265358979323846)


def expand_bit_pattern(spec: str):
    """(bits, repeat_start) of a PULSE / GPULSE bit-pattern specification as the reference expands
    it: pbitList::parse (inptran.cc:312-381: a sequence of `b<string>` items, each with optional
    `r[=N]` repeat count, -1 = forever, and `rb=M` repeat-from bit, 1-based) and pbitAry::add /
    addPRBS (:412-518; `bprbsN`, N = 6..12, the LFSR sequences listed there, rotated by rb-1)."""
    toks = spec.replace("=", " ").split()
    items = []
    i = 0
    while i < len(toks):
        tk = toks[i]
        i += 1
        if tk[0] not in "bB":
            continue                                # unknown token: ignored
        it = dict(s=tk, rb=1, r=0)
        while i < len(toks) and toks[i].lower() in ("r", "rb"):
            key = toks[i].lower()
            if i + 1 < len(toks) and toks[i+1].lstrip("-").isdigit():
                val = int(toks[i+1])
                i += 2
            else:                                   # a bare `r`
                val = 0
                i += 1
            if key == "rb":
                it["rb"] = max(1, val)
            else:
                it["r"] = 0 if val < -1 else val
        items.append(it)
    bits, rst = [], 0

    def prbs(n, bit):
        n1, n3, n4 = n, 1, 1
        if n1 <= 6:
            n1, n2 = 6, 5
        elif n1 == 7:
            n2 = 6
        elif n1 == 8:
            n2, n3, n4 = 6, 5, 4
        elif n1 == 9:
            n2 = 5
        elif n1 == 10:
            n2 = 7
        elif n1 == 11:
            n2 = 9
        else:
            n1, n2, n3, n4 = 12, 11, 8, 6
        mask = (1 << n1) - 1
        start = a = 0x02
        while True:
            if not bit:
                start = a
            if n1 in (8, 12):
                nb = ((a >> (n1-1)) + (a >> (n2-1)) + (a >> (n3-1)) + (a >> (n4-1))) & 1
            else:
                nb = ((a >> (n1-1)) + (a >> (n2-1))) & 1
            a = ((a << 1) | nb) & mask
            bit -= 1
            if bit < 0:
                bits.append(a & 1)
                if a == start:
                    break

    for it in items:
        strt = len(bits)
        body = it["s"][1:]
        if body[:4].lower() == "prbs":
            n = int(body[4:] or 0)
            prbs(n, it["rb"] - 1)
            if it["r"] == -1:
                rst = strt
                if rst == 0:
                    bits.append(bits[0] & 1)
                    rst = 1
                break
            for _ in range(it["r"]):
                prbs(n, it["rb"] - 1)
        else:
            zero = "0fnFN"
            bits.extend(0 if ch in zero else 1 for ch in body)
            if it["r"] == -1:
                rst = strt + it["rb"] - 1
                if rst == 0:
                    bits.append(bits[0] & 1)
                    rst = 1
                break
            for _ in range(it["r"]):
                bits.extend(0 if ch in zero else 1 for ch in it["s"][it["rb"]:])
    return bits, rst


def jj_model_setup(card: dict, dphimax: float = 0.2*3.14159265358979323846) -> dict:
    """What sJJmodel holds after the model pass of JJdev::setup for a `.model NAME jj(...)` card
    whose values are inside the documented ranges (jjset.cc:196-545; out-of-range values are
    reset there with a warning and are not restated here).  `vg` must be on the card: without
    it the reference integrates the BCS gap equation (devlib/tempr.cc:177), which only the
    harvest path (shim/harvest.cc) carries."""
    import math
    c = {k.lower(): v for k, v in card.items()}
    if "vg" not in c:
        raise ValueError("hand-built JJ model needs vg (the BCS default comes through the harvest path)")
    m = dict(rtype=int(c.get("rtype", 1)), ictype=int(c.get("cct", c.get("ictype", 1))),
             pi=int(c.get("pijj", 0)))
    m["vg"] = float(c["vg"])
    m["delv"] = float(c.get("delv", 0.08e-3))
    m["criti"] = float(c.get("icrit", 1e-3))
    cpic = float(c.get("cpic", 1e-9*70.0/100.0))
    m["cap"] = float(c["cap"]) if "cap" in c else cpic*m["criti"]
    cpic = m["cap"]/m["criti"] if "cap" in c else cpic
    m["icfactor"] = float(c.get("icfct", c.get("icfactor", math.pi/4)))
    icrn = float(c.get("icrn", 1.65e-3))
    m["rn"] = float(c["rn"]) if "rn" in c else icrn/m["criti"]
    vm = float(c.get("vm", 16.5e-3))
    m["r0"] = float(c["rsub"]) if "rsub" in c else vm/m["criti"]
    if m["r0"] < m["rn"]:
        m["rn"] = m["r0"]
    m["cmu"] = float(c.get("cmu", 0.0))
    m["gmu"] = float(c.get("gmu", 0.0))
    m["vshunt"] = float(c.get("vshunt", 0.0))
    m["tsfact"] = float(c.get("tsfactor", dphimax/(2*math.pi)))
    m["tsaccl"] = float(c.get("tsaccel", 1.0))
    halfvg = m["vg"]/2
    m["vdpbak"] = min(math.sqrt(PHI0_2PI/cpic), halfvg) if m["cap"] > 0.0 else halfvg
    return m


class CDeck(C.Structure):
    _fields_ = [
        ("nnodes", C.c_int), ("opt", COptions),
        ("tstep", C.c_double), ("tstop", C.c_double), ("tstart", C.c_double),
        ("tmax", C.c_double), ("uic", C.c_int),
        ("nic", C.c_int), ("ic_node", C.POINTER(C.c_int)), ("ic_val", C.POINTER(C.c_double)),
        ("nnodeset", C.c_int), ("ns_node", C.POINTER(C.c_int)), ("ns_val", C.POINTER(C.c_double)),
        ("nres", C.c_int), ("res", C.POINTER(CRes)),
        ("ncap", C.c_int), ("cap", C.POINTER(CCap)),
        ("nind", C.c_int), ("ind", C.POINTER(CInd)),
        ("nmut", C.c_int), ("mut", C.POINTER(CMut)),
        ("nsrc", C.c_int), ("src", C.POINTER(CSrc)),
        ("nmosmod", C.c_int), ("mos_models", C.POINTER(CMosModel)),
        ("nmos", C.c_int), ("mos", C.POINTER(CMos)),
        ("njjmod", C.c_int), ("jj_models", C.POINTER(CJJModel)),
        ("njj", C.c_int), ("jj", C.POINTER(CJJ)),
        ("norder", C.c_int), ("order_kind", C.POINTER(C.c_int)),
        ("order_index", C.POINTER(C.c_int)),
        # TJM junctions come through the harvest path only (their coefficient sets are made by
        # the reference's setup); a hand-built deck has none
        ("ntjmmod", C.c_int), ("tjm_models", C.POINTER(CTJMModel)), ("ntjm", C.c_int),
        ("tjm", C.POINTER(CTJM)),
    ]


# reference defaults: wrspice/include/circuit.h:461-631
DEFAULT_OPTIONS = dict(
    abstol=1e-12, reltol=1e-3, vntol=1e-6, chgtol=1e-14, trtol=7.0, gmin=1e-12, gmax=1e6,
    pivrel=1e-3, pivtol=1e-13, temp=298.15, tnom=298.15, xmu=0.5, dcmu=0.5, trapratio=10.0,
    delmin=0.0, minbreak=0.0, resmin=1e-3, defl=0.0, defw=0.0, defad=0.0, defas=0.0,
    bypass=1, maxord=2, itl1=400, itl2=100, itl4=20, itl_gmin=20, itl_src=20,
    gminsteps=0, srcsteps=0, gminfirst=0, method=0, trapcheck=0, spice3=0, oldsteplim=0,
    oldlimit=0, forcegmin=0, interplev=1, jjaccel=0, nojjtp=0, noopiter=0, steptype=0)

# our option name -> WRspice .options keyword (analysis/optsetp.cc:104-...)
_OPT_SPICE = dict(abstol="abstol", reltol="reltol", vntol="vntol", chgtol="chgtol",
                  trtol="trtol", gmin="gmin", gmax="gmax", pivrel="pivrel", pivtol="pivtol",
                  xmu="xmu", dcmu="dcmu", trapratio="trapratio", delmin="delmin",
                  minbreak="minbreak", resmin="resmin", bypass="bypass", maxord="maxord",
                  itl1="itl1", itl2="itl2", itl4="itl4", itl_gmin="itl2gmin", itl_src="itl2src",
                  gminsteps="gminsteps", srcsteps="srcsteps", interplev="interplev")


def _fmt(x: float) -> str:
    return repr(float(x))


@dataclass
class _Elem:
    kind: int
    name: str
    nodes: tuple
    params: dict
    model: str | None = None


@dataclass
class Deck:
    title: str = "b200 deck"
    elems: list = field(default_factory=list)
    mos_models: dict = field(default_factory=dict)     # name -> dict(type, level, params)
    jj_models: dict = field(default_factory=dict)
    tjm_models: dict = field(default_factory=dict)
    ics: dict = field(default_factory=dict)
    nodesets: dict = field(default_factory=dict)
    opts: dict = field(default_factory=lambda: dict(DEFAULT_OPTIONS))
    tran_args: tuple | None = None

    # ---- construction -------------------------------------------------------------
    def vsource(self, name, p, n, dc=None, pulse=None, pwl=None, sin=None, exp=None, sffm=None,
                am=None, gpulse=None, spulse=None, pattern=None):
        self.elems.append(_Elem(DEV_SRC, name.lower(), (str(p).lower(), str(n).lower()),
                                dict(kind=SRC_V, dc=dc, pulse=pulse, pwl=pwl, sin=sin, exp=exp,
                                     sffm=sffm, am=am, gpulse=gpulse, spulse=spulse, pattern=pattern)))
        return self

    def isource(self, name, p, n, dc=None, pulse=None, pwl=None, sin=None, exp=None, sffm=None,
                am=None, gpulse=None, spulse=None, pattern=None):
        self.elems.append(_Elem(DEV_SRC, name.lower(), (str(p).lower(), str(n).lower()),
                                dict(kind=SRC_I, dc=dc, pulse=pulse, pwl=pwl, sin=sin, exp=exp,
                                     sffm=sffm, am=am, gpulse=gpulse, spulse=spulse, pattern=pattern)))
        return self

    def resistor(self, name, p, n, r):
        self.elems.append(_Elem(DEV_RES, name.lower(), (str(p).lower(), str(n).lower()),
                                dict(r=float(r))))
        return self

    def capacitor(self, name, p, n, c, ic=None, m=None):
        self.elems.append(_Elem(DEV_CAP, name.lower(), (str(p).lower(), str(n).lower()),
                                dict(c=float(c), ic=ic, m=m)))
        return self

    def inductor(self, name, p, n, l, ic=None):
        self.elems.append(_Elem(DEV_IND, name.lower(), (str(p).lower(), str(n).lower()),
                                dict(l=float(l), ic=ic)))
        return self

    def mutual(self, name, l1, l2, k):
        """Mutual inductor coupling `k` between inductors named l1 and l2 (mutprse.cc)."""
        self.elems.append(_Elem(DEV_MUT, name.lower(), (), dict(l1=l1.lower(), l2=l2.lower(),
                                                                  k=float(k))))
        return self

    def jj_model(self, name, **card):
        """`.model NAME jj(card)`; card keys as in the reference (rtype, cct, vg, delv, icrit,
        cap/cpic, icrn/rn, vm/rsub, icfct, cmu, gmu, vshunt, tsfactor, tsaccel, pijj)."""
        self.jj_models[name.lower()] = dict(card)
        return self

    def tjm_model(self, name, setup: str):
        """`.model NAME jj(level=3 ...)` with the card and set-up values of tjm_setup(setup)."""
        self.tjm_models[name.lower()] = tjm_setup(setup)
        return self

    def tjm(self, name, p, n, model, phase=None, **params):
        """Tunnel-junction-model junction `Bname p n [phase] model [ics= area= vshunt=]`."""
        nodes = (str(p).lower(), str(n).lower()) + ((str(phase).lower(),) if phase is not None else ())
        self.elems.append(_Elem(DEV_TJM, name.lower(), nodes, dict(params), model.lower()))
        return self

    def jj(self, name, p, n, model, phase=None, **params):
        """Josephson junction `Bname p n [phase] model [area= ics= vshunt= ic=vj,phi]`."""
        nodes = (str(p).lower(), str(n).lower()) + ((str(phase).lower(),) if phase is not None else ())
        self.elems.append(_Elem(DEV_JJ, name.lower(), nodes, dict(params), model.lower()))
        return self

    def mos_model(self, name, mtype, level=1, **params):
        self.mos_models[name.lower()] = dict(type=+1 if mtype.lower().startswith("n") else -1,
                                             level=level, params=dict(params))
        return self

    def mosfet(self, name, d, g, s, b, model, **params):
        self.elems.append(_Elem(DEV_MOS, name.lower(),
                                tuple(str(x).lower() for x in (d, g, s, b)), dict(params),
                                model.lower()))
        return self

    def ic(self, **kv):
        for k, v in kv.items():
            self.ics[k.lower()] = float(v)
        return self

    def nodeset(self, **kv):
        for k, v in kv.items():
            self.nodesets[k.lower()] = float(v)
        return self

    # optsetp.cc CHECKSET: out-of-range option values are clamped (circuit.h:579-596)
    _OPT_RANGE = {"itl1": (10, 1000), "itl4": (4, 100)}

    def options(self, **kv):
        for k, v in kv.items():
            if k not in DEFAULT_OPTIONS:
                raise KeyError(k)
            if k in self._OPT_RANGE:
                lo, hi = self._OPT_RANGE[k]
                v = min(max(v, lo), hi)
            self.opts[k] = v
        return self

    def tran(self, tstep, tstop, tstart=0.0, tmax=0.0, uic=False):
        self.tran_args = (float(tstep), float(tstop), float(tstart), float(tmax), bool(uic))
        return self

    # ---- derived orderings ----------------------------------------------------------
    def node_numbers(self) -> dict:
        """name -> number, ground names map to 0 (inpdeck.cc:150-162)."""
        nums = {"0": 0, "gnd": 0}
        nxt = 1
        for e in self.elems:
            for nd in e.nodes:
                if nd not in nums:
                    nums[nd] = nxt
                    nxt += 1
        return nums

    def kind_tables(self) -> dict:
        """kind -> list of elems in deck order (the table index used by the C ABI)."""
        tabs = {k: [] for k in range(8)}
        for e in self.elems:
            tabs[e.kind].append(e)
        return tabs

    def load_order(self) -> list:
        """[(kind, index)] in the reference's sCKT::load order."""
        tabs = self.kind_tables()
        index_of = {id(e): i for k in tabs for i, e in enumerate(tabs[k])}
        type_order = []            # model types in order of first model creation
        models_of = {}             # kind -> [model key newest-first]
        insts_of = {}              # (kind, model) -> [elem newest-first]
        for e in self.elems:
            mkey = (e.kind, e.model if e.model is not None else "<default>")
            if e.kind not in models_of:
                models_of[e.kind] = []
                type_order.append(e.kind)
            if mkey not in insts_of:
                insts_of[mkey] = []
                models_of[e.kind].insert(0, mkey)
            insts_of[mkey].insert(0, e)
        # Resistors that touch a phase node (a terminal of a junction with supercurrent, or of
        # an inductor once such a junction exists: jjset.cc:656-681, indset.cc:104-120) are not
        # preloaded; RESdev::temperature moves each of them to the head of its model's list as
        # it walks the list (restemp.cc:240-248), which leaves them in front, in deck order.
        phase = self.phase_nodes()
        if phase:
            for mkey in models_of.get(DEV_RES, []):
                lst = insts_of[mkey]
                moved = [k > 0 and any(nd in phase for nd in e.nodes) for k, e in enumerate(lst)]
                # each moved element lands in front of the earlier ones: reverse visiting order
                insts_of[mkey] = ([e for e, mv in zip(reversed(lst), reversed(moved)) if mv] +
                                  [e for e, mv in zip(lst, moved) if not mv])
        out = []
        if DEV_MUT in models_of:
            for mkey in models_of[DEV_MUT]:
                out += [(DEV_MUT, index_of[id(e)]) for e in insts_of[mkey]]
        for k in type_order:
            if k == DEV_MUT:
                continue
            for mkey in models_of[k]:
                out += [(k, index_of[id(e)]) for e in insts_of[mkey]]
        return out

    def phase_nodes(self) -> set:
        """Names of the nodes flagged `phase` by the reference's setup (phase-mode DC support)."""
        ph = set()
        for e in self.elems:
            if e.kind == DEV_JJ and int(self.jj_models[e.model].get("cct", 1)) > 0:
                ph.update(nd for nd in e.nodes[:2] if nd not in ("0", "gnd"))
            if e.kind == DEV_TJM:
                ph.update(nd for nd in e.nodes[:2] if nd not in ("0", "gnd"))
        if ph:
            for e in self.elems:
                if e.kind == DEV_IND:
                    ph.update(nd for nd in e.nodes if nd not in ("0", "gnd"))
        return ph

    def load_order_names(self) -> list:
        tabs = self.kind_tables()
        return [tabs[k][i].name for k, i in self.load_order()]

    # ---- SPICE text for the compiled reference (tests only) -------------------------
    def to_spice(self, overrides: dict | None = None, analysis: str = "tran") -> str:
        """overrides: {(param_id, table_index): value} for one Monte-Carlo trial; varied MOS
        devices get a private .model, as the reference's MC decks do.  analysis: "tran" or "op"."""
        overrides = overrides or {}
        tabs = self.kind_tables()
        idx = {id(e): i for k in tabs for i, e in enumerate(tabs[k])}
        lines = [self.title]
        extra_models = []
        for e in self.elems:
            i = idx[id(e)]
            if e.kind == DEV_SRC:
                p = e.params
                s = f"{e.name} {e.nodes[0]} {e.nodes[1]}"
                dc = overrides.get((P_SRC_DC, i), p["dc"])
                if dc is not None:
                    s += f" dc {_fmt(dc)}"
                pat = (" " + p["pattern"]) if p.get("pattern") else ""
                if p["pulse"] is not None:
                    s += " pulse(" + " ".join(_fmt(x) for x in p["pulse"]) + pat + ")"
                if p["pwl"] is not None:
                    s += " pwl(" + " ".join(_fmt(x) for x in p["pwl"]) + ")"
                if p["sin"] is not None:
                    s += " sin(" + " ".join(_fmt(x) for x in p["sin"]) + ")"
                for fn in ("exp", "sffm", "am", "gpulse", "spulse"):
                    if p.get(fn) is not None:
                        s += f" {fn}(" + " ".join(_fmt(x) for x in p[fn]) + (pat if fn == "gpulse" else "") + ")"
                lines.append(s)
            elif e.kind == DEV_RES:
                r = overrides.get((P_RES_R, i), e.params["r"])
                lines.append(f"{e.name} {e.nodes[0]} {e.nodes[1]} {_fmt(r)}")
            elif e.kind == DEV_CAP:
                cval = overrides.get((P_CAP_C, i), e.params["c"])
                s = f"{e.name} {e.nodes[0]} {e.nodes[1]} {_fmt(cval)}"
                if e.params.get("ic") is not None:
                    s += f" ic={_fmt(e.params['ic'])}"
                if e.params.get("m") is not None:
                    s += f" m={_fmt(e.params['m'])}"
                lines.append(s)
            elif e.kind == DEV_IND:
                lval = overrides.get((P_IND_L, i), e.params["l"])
                s = f"{e.name} {e.nodes[0]} {e.nodes[1]} {_fmt(lval)}"
                if e.params.get("ic") is not None:
                    s += f" ic={_fmt(e.params['ic'])}"
                lines.append(s)
            elif e.kind == DEV_MUT:
                lines.append(f"{e.name} {e.params['l1']} {e.params['l2']} {_fmt(e.params['k'])}")
            elif e.kind == DEV_MOS:
                mname = e.model
                mod_over = {pid: overrides[(pid, i)] for pid in
                            (P_MOS_VTO, P_MOS_KP, P_MOS_GAMMA, P_MOS_LAMBDA, P_MOS_PHI)
                            if (pid, i) in overrides}
                if mod_over:
                    base = self.mos_models[e.model]
                    prm = dict(base["params"])
                    key = {P_MOS_VTO: "vto", P_MOS_KP: "kp", P_MOS_GAMMA: "gamma",
                           P_MOS_LAMBDA: "lambda", P_MOS_PHI: "phi"}
                    if base["level"] == 6:
                        key[P_MOS_KP] = "kc"
                    for pid, v in mod_over.items():
                        prm[key[pid]] = v
                    mname = f"{e.model}_{e.name}"
                    extra_models.append((mname, base["type"], base["level"], prm))
                prm = dict(e.params)
                if (P_MOS_L, i) in overrides:
                    prm["l"] = overrides[(P_MOS_L, i)]
                if (P_MOS_W, i) in overrides:
                    prm["w"] = overrides[(P_MOS_W, i)]
                s = f"{e.name} {' '.join(e.nodes)} {mname}"
                for k, v in prm.items():
                    if k == "off":
                        if v:
                            s += " off"
                    else:
                        s += f" {k}={_fmt(v)}"
                lines.append(s)
            elif e.kind in (DEV_JJ, DEV_TJM):
                prm = dict(e.params)
                pid = P_JJ_ICS if e.kind == DEV_JJ else P_TJM_ICS
                if (pid, i) in overrides:
                    prm["ics"] = overrides[(pid, i)]
                    prm.pop("area", None)
                sline = f"{e.name} {' '.join(e.nodes)} {e.model}"
                for k, v in prm.items():
                    if k == "ic":
                        sline += f" ic={_fmt(v[0])},{_fmt(v[1])}"
                    else:
                        sline += f" {k}={_fmt(v)}"
                lines.append(sline)
            else:
                raise NotImplementedError(e.kind)
        for name, card in self.jj_models.items():
            lines.append(f".model {name} jj(" + " ".join(f"{k}={_fmt(v) if isinstance(v, float) else v}"
                                                          for k, v in card.items()) + ")")
        for name, su in self.tjm_models.items():
            lines.append(f".model {name} {su['card']}")
        for name, m in self.mos_models.items():
            lines.append(self._model_line(name, m["type"], m["level"], m["params"]))
        for name, t, lev, prm in extra_models:
            lines.append(self._model_line(name, t, lev, prm))
        if self.ics:
            lines.append(".ic " + " ".join(f"v({k})={_fmt(v)}" for k, v in self.ics.items()))
        if self.nodesets:
            lines.append(".nodeset " + " ".join(f"v({k})={_fmt(v)}"
                                                 for k, v in self.nodesets.items()))
        optl = [f"{_OPT_SPICE[k]}={v!r}" for k, v in self.opts.items()
                if v != DEFAULT_OPTIONS[k] and k in _OPT_SPICE]
        if self.opts["method"] == 1:
            optl.append("method=gear")
        for flag in ("trapcheck", "spice3", "oldsteplim", "forcegmin", "jjaccel", "gminfirst"):
            if self.opts[flag]:
                optl.append(flag)
        for k, v in self.opts.items():
            if v != DEFAULT_OPTIONS[k] and k not in _OPT_SPICE and k not in (
                    "method", "trapcheck", "spice3", "oldsteplim", "forcegmin", "jjaccel",
                    "gminfirst"):
                raise NotImplementedError(f"option {k} has no reference .options mapping")
        if optl:
            lines.append(".options " + " ".join(optl))
        if analysis == "op":
            lines.append(".op")
            return "\n".join(lines) + "\n"
        ts, tp, t0, tm, uic = self.tran_args
        s = f".tran {_fmt(ts)} {_fmt(tp)}"
        if t0 or tm:
            s += f" {_fmt(t0)}"
        if tm:
            s += f" {_fmt(tm)}"
        if uic:
            s += " uic"
        lines.append(s)
        return "\n".join(lines) + "\n"

    @staticmethod
    def _model_line(name, mtype, level, prm):
        t = "nmos" if mtype > 0 else "pmos"
        return (f".model {name} {t} level={level} " +
                " ".join(f"{k}={_fmt(v)}" for k, v in prm.items()))

    # ---- C ABI --------------------------------------------------------------------
    def to_c(self):
        """Returns (CDeck, keepalive list)."""
        keep = []
        nums = self.node_numbers()
        tabs = self.kind_tables()
        d = CDeck()
        d.nnodes = max(nums.values()) + 1
        for k, v in self.opts.items():
            setattr(d.opt, k, v)
        ts, tp, t0, tm, uic = self.tran_args
        d.tstep, d.tstop, d.tstart, d.tmax, d.uic = ts, tp, t0, tm, int(uic)

        def arr(ctype, vals):
            a = (ctype * max(1, len(vals)))(*vals)
            keep.append(a)
            return a

        d.nic = len(self.ics)
        d.ic_node = arr(C.c_int, [nums[k] for k in self.ics])
        d.ic_val = arr(C.c_double, list(self.ics.values()))
        d.nnodeset = len(self.nodesets)
        d.ns_node = arr(C.c_int, [nums[k] for k in self.nodesets])
        d.ns_val = arr(C.c_double, list(self.nodesets.values()))

        res = (CRes * max(1, len(tabs[DEV_RES])))()
        for i, e in enumerate(tabs[DEV_RES]):
            res[i] = CRes(nums[e.nodes[0]], nums[e.nodes[1]], e.params["r"], NAN, NAN, NAN, NAN)
        keep.append(res)
        d.nres, d.res = len(tabs[DEV_RES]), res

        cap = (CCap * max(1, len(tabs[DEV_CAP])))()
        for i, e in enumerate(tabs[DEV_CAP]):
            ic = e.params.get("ic")
            mm = e.params.get("m")
            cap[i] = CCap(nums[e.nodes[0]], nums[e.nodes[1]], e.params["c"], NAN if mm is None else float(mm),
                          NAN if ic is None else ic)
        keep.append(cap)
        d.ncap, d.cap = len(tabs[DEV_CAP]), cap

        src = (CSrc * max(1, len(tabs[DEV_SRC])))()
        for i, e in enumerate(tabs[DEV_SRC]):
            p = e.params
            fn, coef = FN_NONE, []
            if p["pulse"] is not None:
                fn, coef = FN_PULSE, list(p["pulse"])
            elif p["pwl"] is not None:
                fn, coef = FN_PWL, list(p["pwl"])
            elif p["sin"] is not None:
                fn, coef = FN_SIN, list(p["sin"])
            elif p.get("exp") is not None:
                fn, coef = FN_EXP, list(p["exp"])
            elif p.get("sffm") is not None:
                fn, coef = FN_SFFM, list(p["sffm"])
            elif p.get("am") is not None:
                fn, coef = FN_AM, list(p["am"])
            elif p.get("gpulse") is not None:
                fn, coef = FN_GPULSE, list(p["gpulse"])
            elif p.get("spulse") is not None:
                fn, coef = FN_SPULSE, list(p["spulse"])
            if p.get("pattern") and fn in (FN_PULSE, FN_GPULSE):
                bits, rst = expand_bit_pattern(p["pattern"])
                if bits:                            # after a NaN: length, repeat start, bits (b200spice.h)
                    coef = coef + [NAN, float(len(bits)), float(rst)] + [float(b) for b in bits]
            ca = arr(C.c_double, [float(x) for x in coef])
            src[i] = CSrc(p["kind"], nums[e.nodes[0]], nums[e.nodes[1]],
                          NAN if p["dc"] is None else float(p["dc"]), fn, len(coef), ca)
        keep.append(src)
        d.nsrc, d.src = len(tabs[DEV_SRC]), src

        mnames = list(self.mos_models)
        mm = (CMosModel * max(1, len(mnames)))()
        alias = {"lambda": "lambda_", "is": "is_", "lambda0": "lamda0", "lambda1": "lamda1"}
        for i, nme in enumerate(mnames):
            m = self.mos_models[nme]
            mm[i].type, mm[i].level = m["type"], m["level"]
            for f in MOS_MODEL_DBL:
                setattr(mm[i], f, NAN)
            for k, v in m["params"].items():
                setattr(mm[i], alias.get(k, k), float(v))
        keep.append(mm)
        d.nmosmod, d.mos_models = len(mnames), mm

        mos = (CMos * max(1, len(tabs[DEV_MOS])))()
        for i, e in enumerate(tabs[DEV_MOS]):
            x = mos[i]
            x.d, x.g, x.s, x.b = (nums[n] for n in e.nodes)
            x.model = mnames.index(e.model)
            for f in ("l", "w", "ad", "as_", "pd", "ps", "nrd", "nrs", "m", "temp", "icvds",
                      "icvgs", "icvbs"):
                setattr(x, f, NAN)
            x.off = 0
            for k, v in e.params.items():
                if k == "off":
                    x.off = int(bool(v))
                else:
                    setattr(x, {"as": "as_"}.get(k, k), float(v))
        keep.append(mos)
        d.nmos, d.mos = len(tabs[DEV_MOS]), mos

        ind = (CInd * max(1, len(tabs[DEV_IND])))()
        for i, e in enumerate(tabs[DEV_IND]):
            ic = e.params.get("ic")
            ind[i] = CInd(nums[e.nodes[0]], nums[e.nodes[1]], e.params["l"], NAN,
                          NAN if ic is None else ic)
        keep.append(ind)
        d.nind, d.ind = len(tabs[DEV_IND]), ind
        inames = [e.name for e in tabs[DEV_IND]]
        mut = (CMut * max(1, len(tabs[DEV_MUT])))()
        for i, e in enumerate(tabs[DEV_MUT]):
            mut[i] = CMut(inames.index(e.params["l1"]), inames.index(e.params["l2"]), e.params["k"])
        keep.append(mut)
        d.nmut, d.mut = len(tabs[DEV_MUT]), mut
        jnames = list(self.jj_models)
        jm = (CJJModel * max(1, len(jnames)))()
        for i, nme in enumerate(jnames):
            for k, v in jj_model_setup(self.jj_models[nme]).items():
                setattr(jm[i], k, v)
        keep.append(jm)
        d.njjmod, d.jj_models = len(jnames), jm
        jj = (CJJ * max(1, len(tabs[DEV_JJ])))()
        for i, e in enumerate(tabs[DEV_JJ]):
            x = jj[i]
            x.pos, x.neg = nums[e.nodes[0]], nums[e.nodes[1]]
            x.phase_node = nums[e.nodes[2]] if len(e.nodes) > 2 else 0
            x.model = jnames.index(e.model)
            for f in ("area", "ics", "vshunt", "vg", "tcf", "icv", "icphi"):
                setattr(x, f, NAN)
            for k, v in e.params.items():
                if k == "ic":
                    x.icv, x.icphi = float(v[0]), float(v[1])
                else:
                    setattr(x, k, float(v))
        keep.append(jj)
        d.njj, d.jj = len(tabs[DEV_JJ]), jj
        tnames = list(self.tjm_models)
        tm = (CTJMModel * max(1, len(tnames)))()
        for i, nme in enumerate(tnames):
            for k, v in self.tjm_models[nme]["model"].items():
                setattr(tm[i], k, v)
        keep.append(tm)
        d.ntjmmod, d.tjm_models = len(tnames), tm
        tj = (CTJM * max(1, len(tabs[DEV_TJM])))()
        for i, e in enumerate(tabs[DEV_TJM]):
            x = tj[i]
            su = self.tjm_models[e.model]
            x.pos, x.neg = nums[e.nodes[0]], nums[e.nodes[1]]
            x.phase_node = nums[e.nodes[2]] if len(e.nodes) > 2 else 0
            x.model = tnames.index(e.model)
            for f in ("area", "ics", "vshunt", "icv", "icphi"):
                setattr(x, f, NAN)
            for k, v in su["instance"].items():
                setattr(x, k, v)
            x.coef = arr(C.c_double, su["coef"])
            for k, v in e.params.items():
                if k == "ic":
                    x.icv, x.icphi = float(v[0]), float(v[1])
                else:
                    setattr(x, k, float(v))
        keep.append(tj)
        d.ntjm, d.tjm = len(tabs[DEV_TJM]), tj
        order = self.load_order()
        d.norder = len(order)
        d.order_kind = arr(C.c_int, [k for k, _ in order])
        d.order_index = arr(C.c_int, [i for _, i in order])
        return d, keep

    def unknown_names(self, generic: list) -> list:
        """Map the engine's generic unknown names ('n3', 'src0#branch', 'mos2#internal#drain')
        to the reference's ('out', 'vdd#branch', 'm2#internal#drain')."""
        nums = self.node_numbers()
        inv = {v: k for k, v in nums.items() if not (k == "gnd")}
        tabs = self.kind_tables()
        out = []
        for g in generic:
            if g.startswith("n") and g[1:].isdigit():
                out.append(inv[int(g[1:])])
            elif g.startswith("src"):
                i, rest = g[3:].split("#", 1)
                out.append(f"{tabs[DEV_SRC][int(i)].name}#{rest}")
            elif g.startswith("mos"):
                i, rest = g[3:].split("#", 1)
                out.append(f"{tabs[DEV_MOS][int(i)].name}#{rest}")
            elif g.startswith("ind"):
                i, rest = g[3:].split("#", 1)
                out.append(f"{tabs[DEV_IND][int(i)].name}#{rest}")
            else:
                out.append(g)
        return out


def ring_oscillator(nstages: int, level: int = 1, vdd: float = 3.3, cload: float = 10e-15,
                    tstep: float = 0.05e-9, tstop: float = 5e-9, uic: bool = False) -> Deck:
    """N-stage CMOS ring oscillator (BASELINE.json configs[0]/[1] shape): 2 MOSFETs + 1 node
    capacitor per stage, one Vdd source; alternating-rail .ic as in
    examples/bsim465-benchmarks ro_17.b4.  Model cards are MOS level-1 cards in the style of
    wrspice/examples/mos6inv.cir."""
    assert nstages % 2 == 1
    d = Deck(title=f"ro{nstages} cmos ring oscillator")
    d.vsource("vdd", "vdd", "0", dc=vdd)
    for i in range(nstages):
        a, b = f"n{i}", f"n{(i + 1) % nstages}"
        d.mosfet(f"mp{i}", b, a, "vdd", "vdd", "pm", w=2e-6, l=1e-6)
        d.mosfet(f"mn{i}", b, a, "0", "0", "nm", w=1e-6, l=1e-6)
        d.capacitor(f"c{i}", b, "0", cload)
    common = dict(gamma=0.4, phi=0.65, cgso=2e-10, cgdo=2e-10, cbd=2e-15, cbs=2e-15,
                  tox=2e-8, **{"lambda": 0.02})
    d.mos_model("nm", "nmos", level=level, vto=0.7, kp=1e-4, **common)
    d.mos_model("pm", "pmos", level=level, vto=-0.7, kp=4e-5, **common)
    d.ic(**{f"n{i}": (vdd if i % 2 == 0 else 0.0) for i in range(nstages)})
    d.tran(tstep, tstop, uic=uic)
    return d


def mc_variation(deck: Deck, ntrials: int, seed: int = 1, sigma_vto: float = 0.03,
                 sigma_l: float = 0.02):
    """Per-trial, per-MOSFET Vth0 and L variation (BASELINE.json configs[1]);
    trial 0 nominal.  The deviates are the reference's own: trial t = 1, 2, ... draws, device by
    device in deck order, gauss() for vto and then gauss() for L from ONE sRnd stream seeded with
    `seed` (xt_base/miscutil/random.cc:51-72 over randval.cc; restated in csrc/mcgen.cpp).
    Returns {(param_id, mos_index): np.ndarray[ntrials]}."""
    import numpy as np
    from .refrand import gauss_stream
    tabs = deck.kind_tables()
    nm = len(tabs[DEV_MOS])
    g = gauss_stream(seed, max(0, ntrials - 1)*nm*2).reshape(max(0, ntrials - 1), nm, 2)
    out = {}
    for i, e in enumerate(tabs[DEV_MOS]):
        vto = np.full(ntrials, float(deck.mos_models[e.model]["params"]["vto"]))
        ln = np.full(ntrials, float(e.params["l"]))
        vto[1:] *= 1.0 + sigma_vto*g[:, i, 0]
        ln[1:] *= 1.0 + sigma_l*g[:, i, 1]
        out[(P_MOS_VTO, i)] = vto
        out[(P_MOS_L, i)] = ln
    return out
