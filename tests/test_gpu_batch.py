"""Stacks of independent 2-D images on the GPU against the per-image oracle."""
import numpy as np
import pytest

from batch_util import check_stack_against_oracle, make_stack

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape,spacing', [((9, 96, 128), 1), ((6, 100, 60), (1.5, 0.5)), ((4, 512, 512), 2)])
def test_stack_matches_per_image_oracle(shape, spacing):
    check_stack_against_oracle(make_stack(shape, seed=sum(shape)), spacing)


def test_valued_stack():
    stack = make_stack((4, 128, 128), seed=3)
    rng = np.random.default_rng(0)
    check_stack_against_oracle((stack * (rng.random(stack.shape) + 0.5)).astype(np.float32), 1)
