"""cubicsdr_b200 -- B200-native spectrum + demodulator engine behind CubicSDR's hot-path seams.

The product is cubicsdr_b200/libsdr_b200.so (CUDA, sm_100a) with the C-ABI of include/sdr_b200.h.  This package is the
thin Python host side used by the tests and bench.py: ctypes bindings plus classes that mirror the reference's own
interfaces for the path (cubicsdr_b200.host).  There is no CPU fallback: `lib()` raises if the CUDA library is
missing, and engine creation fails if no GPU is present.
"""
from ._capi import Channel, Channelizer, Engine, Group, Lib, Mixer, PRODUCT_LIB, Scope, SdrError, Spectrum, SYMBOLS, fetch_pcm16  # noqa: F401
from . import synth  # noqa: F401

_lib = None


def lib():
    """The product library (loaded once)."""
    global _lib
    if _lib is None:
        _lib = Lib(PRODUCT_LIB)
        if not _lib.is_device_build:
            raise SdrError("%s is not a device build" % PRODUCT_LIB)
    return _lib
