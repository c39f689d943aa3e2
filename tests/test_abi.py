"""The C-ABI library loads and exports every symbol include/b200spice.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

from xictools_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b200spice.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_expected_groups():
    syms = declared_symbols()
    for s in ("b200_batch_create", "b200_batch_run_tran", "b200_batch_run_op", "b200_batch_get_waves",
              "b200_klu_analyze", "b200_klu_refactor_solve_host", "b200_batch_debug_load"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_version_string():
    lib = engine.load_library()
    assert b"sm_100a" in lib.b200_version()


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device every compute entry point must refuse, loudly."""
    lib = engine.load_library()
    if lib.b200_cuda_device_count() > 0:
        pytest.skip("GPU present")
    from xictools_b200.deck import ring_oscillator
    with pytest.raises(engine.B200Error) as e:
        engine.Batch(ring_oscillator(3), 2)
    assert e.value.code == engine.E_NOGPU


def test_integration_method_is_validated_before_anything_runs():
    """TSKintegrateMethod is TRAPEZOIDAL (0) or GEAR (1) in the reference (circuit.h); any other
    value is a caller error, reported before a device is even looked for."""
    from xictools_b200.deck import ring_oscillator
    d = ring_oscillator(3)
    d.options(method=2)
    with pytest.raises(engine.B200Error) as e:
        engine.Batch(d, 2)
    assert e.value.code == 11          # B200_E_BADPARM
