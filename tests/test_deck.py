"""Host-side deck logic: node numbering, reference load order, C struct marshalling."""
import numpy as np

from tests.decks import inv_chain, variants
from xictools_b200.deck import (DEV_CAP, DEV_MOS, DEV_RES, DEV_SRC, Deck, mc_variation,
                                ring_oscillator)


def test_load_order_rule():
    d = Deck()
    d.vsource("v1", "a", "0", dc=1.0)
    d.resistor("r1", "a", "b", 1e3)
    d.mosfet("m1", "b", "a", "0", "0", "nm", w=1e-6, l=1e-6)
    d.resistor("r2", "b", "0", 1e3)
    d.mosfet("m2", "b", "a", "0", "0", "pm", w=1e-6, l=1e-6)
    d.mosfet("m3", "b", "a", "0", "0", "nm", w=1e-6, l=1e-6)
    d.mos_model("nm", "nmos", vto=0.7)
    d.mos_model("pm", "pmos", vto=-0.7)
    d.tran(1e-9, 1e-8)
    # types in order of first model creation (src, res, mos); models newest-first;
    # instances newest-first within a model
    assert d.load_order_names() == ["v1", "r2", "r1", "m2", "m3", "m1"]


def test_node_numbers_first_appearance():
    d = ring_oscillator(3)
    nums = d.node_numbers()
    assert nums["0"] == 0 and nums["vdd"] == 1 and nums["n1"] == 2 and nums["n0"] == 3


def test_c_struct_roundtrip():
    d = inv_chain(3)
    cd, keep = d.to_c()
    assert cd.nmos == 6 and cd.ncap == 3 and cd.nres == 1 and cd.nsrc == 2
    assert cd.norder == 12
    kinds = [cd.order_kind[i] for i in range(cd.norder)]
    assert kinds.count(DEV_MOS) == 6 and kinds.count(DEV_SRC) == 2
    assert kinds.count(DEV_CAP) == 3 and kinds.count(DEV_RES) == 1
    assert cd.src[1].fn == 1 and cd.src[1].ncoef == 7
    assert np.isnan(cd.mos[0].ad) and cd.mos[0].l == 1e-6


def test_mc_variation_trial0_is_nominal():
    d = ring_oscillator(5)
    ov = mc_variation(d, 6, seed=1)
    for (pid, i), v in ov.items():
        assert v.shape == (6,)
        assert len(set(v)) > 1
    assert ov[(1, 0)][0] == -0.7 and ov[(3, 0)][0] == 1e-6


def test_spice_text_mentions_every_element():
    for name, d in variants().items():
        txt = d.to_spice()
        for e in d.elems:
            assert txt.count("\n" + e.name + " ") == 1, (name, e.name)
        assert ".tran" in txt
