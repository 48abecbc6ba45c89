"""The guard-gap mechanism of the native runners (tests/guard_util.py) through the emulation harness: gaps are recorded,
survive a run, and a deliberately damaged gap is reported."""
import numpy as np
import pytest
import torch

from emul_util import emulated_device
from guard_util import guarded
from skan_b200 import _pipeline
from skan_b200.synth import walks


def test_guard_gaps_survive_and_damage_is_seen():
    img = walks((24, 28, 32), 10, 60, seed=3, isolated=4, rings=3)
    with emulated_device():
        for fill in (0x00, 0xA5):
            with guarded(fill) as g:
                _pipeline.run_native(torch.from_numpy(img), spacing=(0.5, 0.1, 0.1), device='cpu')
                n = g.check_gaps()
                assert n > 10
                offs = np.zeros(512, dtype=np.int64)
                k = g.lib._dll.skb_debug_guard_offsets(0, offs.ctypes.data, 512)
                _pipeline.LAST_ARENAS[0][int(offs[k // 2]) + 7] = fill ^ 0x5A
                with pytest.raises(AssertionError):
                    g.check_gaps()
