"""The partition-independent result digest (skan_b200/digest.py): a single-device run through the emulation
harness, the CPU oracle and (tests/slab_worker.py) every Z-slab run must agree; a changed array must not."""
import numpy as np
import torch

from emul_util import emulated_device
from oracle import skan_oracle as orc
from skan_b200 import _pipeline, digest
from skan_b200.synth import walks


def _case(valued):
    img = walks((24, 28, 32), 10, 60, seed=3, isolated=4, rings=3)
    if valued:
        img = (img * (np.random.default_rng(1).random(img.shape) + 0.5)).astype(np.float32)
    return img


def test_digest_matches_oracle_and_detects_changes():
    for valued in (False, True):
        img = _case(valued)
        sk, summ = orc.skeleton_and_summary(img, spacing=(0.5, 0.1, 0.1))
        want = digest.digest_oracle(sk, summ)
        with emulated_device():
            r = _pipeline.run_native(torch.from_numpy(img), spacing=(0.5, 0.1, 0.1), device='cpu')
            got = digest.digest_single(r)
            assert digest.same(got, want), (got, want)
            assert got['digest'] == want['digest']
            r.paths.indices[3] += 1
            assert not digest.same(digest.digest_single(r), want)


def test_linear_hash_is_additive_over_partitions():
    rng = np.random.default_rng(0)
    a = rng.integers(-2**40, 2**40, size=1000)
    whole = digest.linear_hash(a, 0, 5)
    parts = (digest.linear_hash(a[:300], 0, 5) + digest.linear_hash(a[300:], 300, 5)) % 2**64
    assert whole == parts
    assert digest.linear_hash(a, 1, 5) != whole and digest.linear_hash(a, 0, 6) != whole
