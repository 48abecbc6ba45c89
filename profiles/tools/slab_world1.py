#!/usr/bin/env python
"""The Z-slab runner with ONE rank on a whole volume, next to the single-image runner on the same volume: what the slab
machinery (exchanges with nobody, its synchronisations, the sums carried through the ranking, the replicated stages)
costs before any inter-rank skew.   python profiles/tools/slab_world1.py [workload]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from skan_b200 import _pipeline, slab  # noqa: E402
from skan_b200.synth import walks_device  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'volume_1024'
    shape, nw, steps, seed, iso, rings, spacing, valued = bench.WORKLOADS[workload]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT='29533', RANK='0', WORLD_SIZE='1')
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    dist.init_process_group('nccl', rank=0, world_size=1, device_id=dev)
    img = walks_device(shape, nw, steps, seed, dev, isolated=iso, rings=rings)
    t_single = timed(lambda: _pipeline.run_native(img, spacing=spacing, device=dev))
    for mode in ('input', 'exchange'):
        kw = dict(ze0=0) if mode == 'input' else dict(halo='exchange')
        f = lambda: slab.skeleton_and_summary_slab(img, global_shape=shape, z0=0, z1=shape[0], spacing=spacing, device=dev, **kw)
        t = timed(f)
        slab.TIME_SWEEP = 2
        r = f()
        torch.cuda.synchronize()
        slab.TIME_SWEEP = False
        print(f'{workload}: single-image runner {t_single:.3f} ms; slab runner, one rank, halo {mode}: {t:.3f} ms '
              f'(syncs {r.counts["syncs"]}, exchanges {r.counts["exchanges"]})')
        for k, v in r.counts['host_us'].items():
            print(f'    {k:24s} {v / 1e3:7.3f} ms')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
