#!/usr/bin/env python
"""A/B of the chain-walk scheduling and of the capacity hints on one GPU, one process, one resident volume:

    python profiles/tools/ab_walks.py [workload] > gpurun_out/ab_walks.txt

For every variant (one walk per thread; lanes refilled from the ticket queue at several refill thresholds) it
prints the step time (CUDA events around 10 steps after 3 warm-ups) and the runner's stage clock.  Results do
not depend on the variant (tests/test_gpu_parity.py runs both)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from skan_b200 import _native, _pipeline  # noqa: E402
from skan_b200.synth import walks_device  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'volume_2048'
    shape, nw, steps, seed, iso, rings, spacing, valued = bench.WORKLOADS[workload]
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    img = walks_device(shape, nw, steps, seed, dev, isolated=iso, rings=rings)
    lib = _native.library()
    V = int(np.prod(shape, dtype=np.int64))
    # (name, dynamic walks, refill threshold, capacity hints)
    variants = [('static, no hints', 0, 8, False), ('static, hints', 0, 8, True),
                ('static, no hints (2)', 0, 8, False), ('static, hints (2)', 0, 8, True)]
    for name, dyn, refill, hints in variants:
        lib.call('skb_set_option', b'dynamic_walks', dyn)
        lib.call('skb_set_option', b'walk_refill', refill)
        _pipeline.USE_HINTS = hints
        for _ in range(3):
            _pipeline.run_native(img, spacing=spacing, device=dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            _pipeline.run_native(img, spacing=spacing, device=dev)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        _pipeline.TIME_SWEEP = 2
        r = _pipeline.run_native(img, spacing=spacing, device=dev)
        torch.cuda.synchronize()
        _pipeline.TIME_SWEEP = False
        st = r.counts['stage_ns']
        keys = ('sweep', 'nodes', 'raw_stencil', 'junction_count', 'junction_mst', 'degrees', 'csr_records', 'walk',
                'rank', 'select', 'heads', 'emit', 'cycles', 'skeleton_ids', 'stats', 'table')
        print(f'{name:22s} {ms:7.3f} ms/step  {V / ms / 1e3:11.0f} Mvox/s  syncs={r.counts["syncs"]} '
              f'overlapped={r.counts["overlapped"]} retaken={r.counts["retaken"]} | ' +
              ' '.join(f'{k}={st.get(k, 0) * 1e-6:.3f}' for k in keys), flush=True)
    lib.call('skb_set_option', b'dynamic_walks', 0)
    lib.call('skb_set_option', b'walk_refill', 8)


if __name__ == '__main__':
    main()
