"""Shared fixtures.  `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200 and goes through the C-ABI."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The oracle: reference liquid-dsp 1.5.0 binary + CubicSDR glue restatement (oracle/_ref/libsdr_ref.so)."""
    from oracle import refbind
    if not os.path.exists(refbind.REF_LIB) and not os.path.exists("/root/reference"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return refbind.Ref()


@pytest.fixture(scope="session")
def product_lib():
    """libsdr_b200.so (nvcc, sm_100a).  Loading it needs no GPU; creating an engine does."""
    from cubicsdr_b200 import build, Lib
    return Lib(build.build())


@pytest.fixture(scope="session")
def emu_lib():
    """Test-only serial replay of the kernels' phases (tests/emu); never used by the package."""
    from tests.emu import build_emu
    from cubicsdr_b200 import Lib
    return Lib(build_emu.build())


@pytest.fixture(scope="session")
def gpu_lib(product_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    assert product_lib.is_device_build
    return product_lib
