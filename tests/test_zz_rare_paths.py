"""GPU tier, sorted last on purpose: rare paths that only short blocks reach (history mixed with several oscillator
settings, dsp_core.h premix_thread)."""
import pytest

from tests.test_fuzz import _case


@pytest.mark.gpu
def test_short_blocks_around_retunes_gpu(gpu_lib, ref):
    for seed in range(9100, 9130):
        _case(gpu_lib, ref, seed, tiny=True)
