"""Memory checks of our own on a real GPU (compute-sanitizer is closed on the pool): see tests/guard_util.py."""
import numpy as np
import pytest
import torch

from guard_util import guarded
from skan_b200 import _native, _pipeline, digest
from skan_b200.synth import walks

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(4)
    vol = walks((72, 80, 96), 40, 120, seed=3, isolated=8, rings=12)
    img = walks((300, 260), 40, 150, seed=5, isolated=6, rings=30)
    yield 'volume', vol, (0.5, 0.1, 0.1), {}
    yield 'image', img, 1, {}
    yield 'valued', (img * (rng.random(img.shape) + 0.5)).astype(np.float32), (2.0, 0.5), {}
    yield 'dense', rng.random((40, 40, 40)) < 0.12, 1, {}
    yield 'stack', np.stack([walks((64, 64), 6, 50, seed=20 + i, isolated=2) for i in range(12)]), 1, dict(image_stack=True)
    yield 'ragged', walks((37, 41, 53), 20, 60, seed=9, isolated=3, rings=4), 1, {}


@pytest.mark.parametrize('variant', ['tma', 'ldg', 'fused'])
def test_no_out_of_bounds_write_and_no_read_of_garbage(variant):
    dev = torch.device('cuda', torch.cuda.current_device())
    old = _pipeline.default_pack_variant
    _pipeline.default_pack_variant = variant
    lib = _native.library()
    try:
        for name, img, spacing, kw in _cases():
            seen = {}
            for small_cycles in (1, 0):
                lib.call('skb_set_option', b'small_cycles', small_cycles)
                lib.call('skb_set_option', b'junction_append', small_cycles)   # both forms of the junction edge emission
                for fill in (0x00, 0xA5, 0xFF):
                    with guarded(fill) as g:
                        for rep in range(2):          # the second call runs with capacity hints
                            r = _pipeline.run_native(torch.from_numpy(img).to(dev), spacing=spacing, device=dev, **kw)
                            torch.cuda.synchronize()
                            assert g.check_gaps() > 10
                            d = digest.digest_single(r)
                            seen.setdefault('first', d)
                            assert d['components'] == seen['first']['components'], (name, variant, fill, rep, small_cycles)
                            assert d['fp_sums'] == seen['first']['fp_sums'], (name, variant, fill, rep, small_cycles)
    finally:
        _pipeline.default_pack_variant = old
        lib.call('skb_set_option', b'small_cycles', 1)
        lib.call('skb_set_option', b'junction_append', 1)


def test_repeated_runs_are_bit_identical():
    """Atomics and inter-CTA protocols (look-back scan, Boruvka rounds, union-find) must not make results depend on timing."""
    dev = torch.device('cuda', torch.cuda.current_device())
    img = torch.from_numpy(walks((160, 192, 224), 400, 200, seed=13, isolated=40, rings=60)).to(dev)
    first = None
    for rep in range(25):
        r = _pipeline.run_native(img, spacing=(0.5, 0.1, 0.1), device=dev)
        keep = [t.clone() for t in (r.graph.indices, r.paths.indices, r.paths.indptr, r.table[0], r.table[1], r.dist, r.table[2])]
        if first is None:
            first = keep
        else:
            for a, b in zip(first, keep):
                assert torch.equal(a.view(torch.uint8), b.view(torch.uint8)), rep
