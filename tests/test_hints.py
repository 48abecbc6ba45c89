"""Capacity hints of the native runner (see hint_util.py) through the host emulation harness and on the GPU."""
import pytest
import torch

from emul_util import emulated_device
from hint_util import run_sequence


@pytest.mark.parametrize('shape,spacing', [((48, 64), 1), ((20, 24, 28), (0.5, 0.1, 0.1))])
def test_hints_emulated(shape, spacing):
    with emulated_device():
        run_sequence(shape, spacing)


@pytest.mark.gpu
@pytest.mark.parametrize('shape,spacing', [((480, 640), 1), ((96, 112, 128), (0.5, 0.1, 0.1))])
def test_hints_gpu(shape, spacing):
    run_sequence(shape, spacing, device_image=lambda a: torch.from_numpy(a).cuda())
