#!/usr/bin/env python
"""A/B of the walk options (both-ways emit, splitter density) and of the capacity hints on one GPU,
one process, one resident volume:

    python profiles/tools/ab_walks.py [workload] > gpurun_out/ab_walks.txt

For every variant it prints the step time (CUDA events around 10 steps after 3 warm-ups), the read-back counters and
the runner's stage clock.  Results do not depend on the variant (tests/test_gpu_parity.py and
tests/test_zy_pixel_graph.py compare every option with the oracle)."""
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
    # (name, options of skb_set_option, capacity hints)
    base = dict(emit_both_ways=1, splitter_bits=0, mst_keys=1, mst_blocks_per_sm=0, emit_blocks_per_sm=16, junction_append=1,
                pdl=1)
    variants = [('default (1/16, both ways)', {}, True),
                ('splitters 1/32, one way (r01)', dict(splitter_bits=5, emit_both_ways=0), True),
                ('splitters 1/32, both ways', dict(splitter_bits=5), True),
                ('splitters 1/16, one way', dict(emit_both_ways=0), True),
                ('splitters 1/8, both ways', dict(splitter_bits=3), True),
                ('splitters 1/16, both ways', dict(splitter_bits=4), True),
                ('no hints', {}, False),
                ('plain launches (no pdl)', dict(pdl=0), True),
                ('mst two-phase rounds (r01)', dict(mst_keys=0), True),
                ('junction edges: count + scan + emit', dict(junction_append=0), True),
                ('default (again)', {}, True)]
    only = os.environ.get('AB_ONLY')
    if only:
        variants = [v for v in variants if any(k in v[0] for k in only.split(','))]
    for name, opts, hints in variants:
        for k, v in {**base, **opts}.items():
            lib.call('skb_set_option', k.encode(), v)
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
        print(f'{name:28s} {ms:7.3f} ms/step  {V / ms / 1e3:11.0f} Mvox/s  syncs={r.counts["syncs"]} '
              f'overlapped={r.counts["overlapped"]} retaken={r.counts["retaken"]} | ' +
              ' '.join(f'{k}={st.get(k, 0) * 1e-6:.3f}' for k in keys), flush=True)
    for k, v in base.items():
        lib.call('skb_set_option', k.encode(), v)
    _pipeline.USE_HINTS = True


if __name__ == '__main__':
    main()
