"""A stack of independent 2-D images in one pass (skan_b200.batch) against the oracle run image
by image -- through the host emulation harness here, through the CUDA library in test_gpu_batch."""
import numpy as np
import pytest

from emul_util import emulated_device
from batch_util import check_stack_against_oracle, make_stack


@pytest.fixture(scope='module', autouse=True)
def _emulated():
    with emulated_device():
        yield


@pytest.mark.parametrize('shape,spacing', [((7, 48, 64), 1), ((5, 40, 40), (1.5, 0.5)), ((3, 128, 128), 2)])
def test_stack_matches_per_image_oracle(shape, spacing):
    check_stack_against_oracle(make_stack(shape, seed=sum(shape)), spacing)


def test_valued_stack():
    stack = make_stack((4, 64, 64), seed=3)
    rng = np.random.default_rng(0)
    check_stack_against_oracle((stack * (rng.random(stack.shape) + 0.5)).astype(np.float32), 1)
