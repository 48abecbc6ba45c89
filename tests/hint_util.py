"""Shared body of the capacity-hint tests (emulated and GPU): skb_run_image sizes arrays by hints taken from
the previous image of the same shape and overlaps the read-back of the exact counts with the kernels that fill
them; results must not depend on the hints, whether they are exact, too large or too small."""
import numpy as np

from parity_util import compare_with_oracle
from skan_b200 import _pipeline
from skan_b200.synth import walks


def images(shape):
    rng = np.random.default_rng(7)
    sparse = walks(shape, 6, 40, seed=1, isolated=2, rings=1)
    dense = walks(shape, 60, 120, seed=2, isolated=30, rings=6)
    blob = rng.random(shape) < 0.35             # big junction clusters: many junction edges
    ring_only = np.zeros(shape, bool)           # only junction-free cycles: no regular path at all
    if len(shape) == 2:
        ring_only[4:9, 4] = ring_only[4:9, 8] = True
        ring_only[4, 4:9] = ring_only[8, 4:9] = True
    else:
        ring_only[2, 4:9, 4] = ring_only[2, 4:9, 8] = True
        ring_only[2, 4, 4:9] = ring_only[2, 8, 4:9] = True
    return dict(sparse=sparse, dense=dense, blob=blob, ring_only=ring_only)


def last_counts():
    (cache,) = [c for c in _pipeline._arena_cache.values() if c.get('counts')][-1:]
    return cache['counts']


def run_sequence(shape, spacing, device_image=lambda a: a):
    imgs = images(shape)
    order = ['sparse', 'sparse', 'dense', 'dense', 'blob', 'sparse', 'ring_only', 'dense', 'ring_only', 'ring_only',
             'blob', 'blob']
    seen = dict(overlapped=0, retaken=0)
    _pipeline._arena_cache.clear()
    slack, _pipeline.HINT_SLACK = _pipeline.HINT_SLACK, 1   # small test images must be able to outgrow a hint
    try:
        _run(imgs, order, seen, shape, spacing, device_image)
    finally:
        _pipeline.HINT_SLACK = slack


def _run(imgs, order, seen, shape, spacing, device_image):
    for i, name in enumerate(order):
        compare_with_oracle(imgs[name], spacing, image_for_device=device_image(imgs[name]))
        c = last_counts()
        if i == 0:
            assert c['overlapped'] <= 1 and c['retaken'] == 0     # no hints yet (only the entry count overlaps)
        if i > 0 and order[i - 1] == name:
            assert c['overlapped'] >= 5 and c['retaken'] == 0, (name, c)   # exact hints: every read-back overlapped
        seen['overlapped'] += c['overlapped']
        seen['retaken'] += c['retaken']
    assert seen['retaken'] >= 3        # sparse -> dense, dense -> blob, ... : hints too small were detected
    # valued image: node values travel through the same hinted arrays
    rng = np.random.default_rng(3)
    valued = (imgs['dense'] * (rng.random(shape) + 0.5)).astype(np.float32)
    for _ in range(2):
        compare_with_oracle(valued, spacing, image_for_device=device_image(valued))
    assert last_counts()['overlapped'] >= 5
    # hints switched off: the plain path
    _pipeline.USE_HINTS = False
    try:
        compare_with_oracle(imgs['dense'], spacing, image_for_device=device_image(imgs['dense']))
        assert last_counts()['overlapped'] <= 1
    finally:
        _pipeline.USE_HINTS = True
