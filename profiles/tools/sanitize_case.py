#!/usr/bin/env python
"""Small end-to-end workload for compute-sanitizer (one tool per run):

    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python profiles/tools/sanitize_case.py

Runs Skeleton + summarize through the native runner for the three sweep variants, a valued image, a stack of images, ring
images on both junction-free-cycle paths, and the stage-level path, and compares with the CPU oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import skan_b200  # noqa: E402
from skan_b200 import _native, _pipeline, batch  # noqa: E402
from skan_b200.synth import walks  # noqa: E402
from oracle import skan_oracle as orc  # noqa: E402


def check(img, spacing):
    sk = skan_b200.Skeleton(torch.from_numpy(img).cuda(), spacing=spacing)
    df = skan_b200.summarize(sk, separator='_')
    ref = orc.OracleSkeleton(img, spacing=spacing)
    rs = orc.summarize(ref)
    assert np.array_equal(sk.graph.indices, ref.graph.indices) and np.array_equal(sk.paths.indices, ref.paths.indices)
    assert np.array_equal(df['skeleton_id'].to_numpy(), rs['skeleton_id'])
    np.testing.assert_allclose(df['branch_distance'].to_numpy(), rs['branch_distance'], rtol=1e-12)


def main():
    torch.cuda.set_device(0)
    lib = _native.library()
    vol = walks((72, 80, 96), 40, 120, seed=3, isolated=8, rings=12)
    for variant in ('tma', 'ldg', 'fused'):
        _pipeline.default_pack_variant = variant
        check(vol, (0.5, 0.1, 0.1))
    _pipeline.default_pack_variant = 'tma'
    img = walks((300, 260), 40, 150, seed=5, isolated=6, rings=30)
    for small in (1, 0):
        lib.call('skb_set_option', b'small_cycles', small)
        check(img, 1)
    lib.call('skb_set_option', b'small_cycles', 1)
    check((img * (np.random.default_rng(1).random(img.shape) + 0.5)).astype(np.float32), (2.0, 0.5))
    dense = np.random.default_rng(2).random((40, 40, 40)) < 0.12     # large junction clusters: many Boruvka rounds
    check(dense, 1)
    stack = np.stack([walks((64, 64), 6, 50, seed=20 + i, isolated=2) for i in range(12)])
    df = batch.summarize_images(torch.from_numpy(stack).cuda(), spacing=1)
    assert len(df) > 0
    g = _pipeline.skeleton_and_summary_device(torch.from_numpy(vol).cuda(), spacing=(0.5, 0.1, 0.1), device=torch.device('cuda', 0))
    assert g.paths.n_paths > 0
    torch.cuda.synchronize()
    print('SANITIZE-CASE-OK')


if __name__ == '__main__':
    main()
