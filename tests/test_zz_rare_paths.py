"""GPU tier of the seeded fuzz tests (tests/test_fuzz.py), in a file that sorts last on purpose: broad random coverage
after the targeted parity tests, and the rare paths only short blocks reach (history mixed with several oscillator
settings, dsp_core.h premix_thread)."""
import pytest

from tests.test_fuzz import _case, _channelizer_case


@pytest.mark.gpu
def test_short_blocks_around_retunes_gpu(gpu_lib, ref):
    for seed in range(9100, 9130):
        _case(gpu_lib, ref, seed, tiny=True)


@pytest.mark.gpu
def test_random_configurations_gpu(gpu_lib, ref):
    for seed in range(7000, 7040):
        _case(gpu_lib, ref, seed)


@pytest.mark.gpu
def test_random_channelizer_configurations_gpu(gpu_lib, ref):
    for seed in range(300, 330):
        _channelizer_case(gpu_lib, ref, seed)
