"""The C-ABI library loads on a CPU-only box and exports every symbol include/sdr_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "sdr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sdr_[a-z0-9_]+)\s*\(", txt)))


def test_header_matches_binding_table():
    from cubicsdr_b200 import SYMBOLS
    assert header_symbols() == sorted(SYMBOLS)


def test_product_library_exports_every_symbol(product_lib):
    for name in header_symbols():
        assert hasattr(product_lib.L, name), name
    assert product_lib.is_device_build
    assert b"sm_100a" in product_lib.L.sdr_version()


def test_product_has_no_cpu_path(product_lib):
    """Without a GPU, engine creation must fail loudly (never fall back)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = product_lib.L.sdr_engine_create(0, 1024, None, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in product_lib.L.sdr_last_error() or b"cuda" in product_lib.L.sdr_last_error().lower()


def test_package_refuses_emulator():
    """cubicsdr_b200.lib() only ever binds the device build."""
    import cubicsdr_b200
    assert cubicsdr_b200.PRODUCT_LIB.endswith("cubicsdr_b200/libsdr_b200.so")
    src = open(os.path.join(ROOT, "cubicsdr_b200", "__init__.py")).read() + open(os.path.join(ROOT, "cubicsdr_b200", "_capi.py")).read()
    assert "oracle" not in src.replace("no CPU fallback", "") or "import oracle" not in src
    assert "emu" not in src.lower().replace("emulator", "")
