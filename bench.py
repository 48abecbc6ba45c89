#!/usr/bin/env python
"""bench.py -- skeleton Mvox/s of Skeleton + summarize on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W         # CPU port of the reference path
    torchrun ... bench.py --gpus N ...                            # one rank per GPU

A "step" is one pass of the whole hot path (mask sweep -> pixel graph -> junction MST -> CSR ->
paths -> statistics -> branch table) over one synthetic skeleton volume.  `value` is measured
with the image already resident in HBM (CUDA events, max over ranks); `e2e` is the same job
through the public API (skan_b200.Skeleton + skan_b200.summarize) from a pinned HOST image,
H2D and D2H copies inside the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'skeleton_mvox_per_s'
UNIT = 'Mvox/s'

# name -> (shape, walkers, steps, seed, isolated, rings, spacing, valued)
WORKLOADS = {
    # BASELINE.json configs[4] / north_star target volume (whole volume on one GPU at N=1)
    'volume_2048': ((2048, 2048, 2048), 16384, 512, 5, 4000, 256, (0.5, 0.1, 0.1), False),
    # configs[3]
    'volume_1024': ((1024, 1024, 1024), 2048, 512, 4, 1000, 64, (0.5, 0.1, 0.1), False),
    # configs[1]: float32-valued 2-D skeleton (path_means over the image's own values)
    'image_16384': ((16384, 16384), 8192, 512, 2, 1000, 64, 1, True),
    # configs[2]: 8192 independent 512 x 512 images pushed through the kernels as one stack; ranks take
    # contiguous shares of the stack, no collective
    'batch_512': ((8192, 512, 512), 24, 192, 1000, 4, 0, 1, False),
    # small case for quick checks
    'volume_256': ((256, 256, 256), 256, 256, 6, 100, 16, (0.5, 0.1, 0.1), False),
}


def ncu_traffic(workload, kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of the same workload (profiles/r01_ncu_full_<workload>.csv), or None."""
    import csv
    path = os.path.join(ROOT, 'profiles', f'r01_ncu_full_{workload}.csv')
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ni, ri, wi = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
        for d in rows[2:]:
            if kernel_prefix in d[ni]:
                return int(float(d[ri]) * scale[units[ri]] + float(d[wi]) * scale[units[wi]])
    except Exception:
        pass
    return None


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(',')]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)
        return False

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.samples)}


def cpu_baseline(workload, seconds_hint=15.0):
    """The oracle (oracle/skan_oracle.py, a NumPy/SciPy/numba port of the reference path) timed on
    this box's host cores on a bounded sample of the workload: the first Z-planes (rows in 2-D)."""
    from oracle import skan_oracle as orc
    from skan_b200.synth import walks
    shape, nw, steps, seed, iso, rings, spacing, valued = WORKLOADS[workload]
    # warm numba
    orc.skeleton_and_summary(walks((32, 32, 32), 8, 24, seed=1), spacing=(0.5, 0.1, 0.1))
    frac = {'volume_2048': 2, 'volume_1024': 1, 'image_16384': 1, 'volume_256': 1, 'batch_512': 64}[workload]
    sub = (shape[0] // frac,) + tuple(shape[1:])
    if workload.startswith('batch'):
        from skan_b200.synth import stack_walks_points
        img = np.zeros(int(np.prod(sub)), bool)
        img[stack_walks_points(sub[0], sub[1], sub[2], nw, steps, seed, isolated=iso)] = True
        img = img.reshape(sub)
        t0 = time.perf_counter()
        n_paths = 0
        for i in range(sub[0]):
            try:
                sk, _ = orc.skeleton_and_summary(img[i], spacing=spacing)
                n_paths += sk.n_paths
            except ValueError:
                pass
        dt = time.perf_counter() - t0
        return {'value': img.size / dt / 1e6, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                'sample': f'first {sub[0]} of {shape[0]} images of {workload}, one by one ({n_paths} paths), {dt:.2f} s, '
                          'single thread like the reference', 'seconds': dt}
    img = walks(sub, max(8, nw // frac), steps, seed, isolated=iso // frac, rings=rings // frac)
    if valued:
        rng = np.random.default_rng(seed)
        img = (img * (rng.random(sub, dtype=np.float32) * 0.9 + 0.1)).astype(np.float32)
    t0 = time.perf_counter()
    sk, _ = orc.skeleton_and_summary(img, spacing=spacing)
    dt = time.perf_counter() - t0
    return {'value': img.size / dt / 1e6, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': f'first {sub[0]} of {shape[0]} planes of {workload} ({img.size / 1e6:.0f} Mvox, '
                      f'{sk.graph.shape[0]} nodes, {sk.n_paths} paths), {dt:.2f} s, single thread like the reference',
            'seconds': dt}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: the reference is pure Python and
    cannot travel to the GPU box) on the host cores, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(args.workload)
        if i >= args.warmup:
            vals.append(r)
        if sum(v['seconds'] for v in vals) > 240:
            break
    v = float(np.mean([x['value'] for x in vals]))
    ms = float(np.mean([x['seconds'] for x in vals])) * 1e3
    shape = WORKLOADS[args.workload][0]
    line = {'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(vals), 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'u8',
            'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': args.workload, 'shape': list(shape)},
            'cpu_baseline': {k: vals[-1][k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    line['cpu_baseline']['value'] = v
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='skan_b200')
    ap.add_argument('--workload', default='volume_2048', choices=sorted(WORKLOADS))
    ap.add_argument('--pack-variant', default=None, choices=['ldg', 'tma', 'fused'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--halo', type=int, default=2, help='halo planes of the Z-slab decomposition (N > 1)')
    ap.add_argument('--timeline', action='store_true', help='N > 1: print the host timeline of one step to stderr')
    ap.add_argument('--stages', action='store_true', help='also print a per-entry-point timing table to stderr')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}')
    if rank == 0:
        entry.build()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')  # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()
    import skan_b200
    from skan_b200 import _native, _pipeline, slab
    from skan_b200.synth import walks_device, walks_points

    if args.pack_variant:
        _pipeline.default_pack_variant = args.pack_variant
    shape, nw, steps, seed, iso, rings, spacing, valued = WORKLOADS[args.workload]
    V = int(np.prod(shape, dtype=np.int64))
    is_stack = args.workload.startswith('batch')
    if world > 1 and not is_stack and (len(shape) != 3 or valued):
        raise SystemExit('multi-GPU runs use the Z-slab decomposition of a 3-D volume (volume_* workloads) '
                         'or a stack of independent images (batch_512)')

    # ---- synthetic input, resident in HBM ------------------------------------------------------
    esize, dtype_name = (4, 'f32') if valued else (1, 'u8')
    if is_stack:
        # this rank's contiguous share of the stack of independent images
        from skan_b200 import batch
        from skan_b200.synth import stack_walks_points
        b0, b1 = rank * shape[0] // world, (rank + 1) * shape[0] // world
        pts = stack_walks_points(b1 - b0, shape[1], shape[2], nw, steps, seed + rank, isolated=iso)
        img = torch.zeros((b1 - b0) * shape[1] * shape[2], dtype=torch.bool, device=dev)
        img[torch.from_numpy(pts).to(dev)] = True
        img = img.reshape(b1 - b0, shape[1], shape[2])
        V_local = int(img.numel())
        del pts
    elif world == 1:
        if valued:
            mask = walks_device(shape, nw, steps, seed, dev, isolated=iso, rings=rings)
            gen = torch.Generator(device=dev).manual_seed(seed)
            img = (torch.rand(shape, device=dev, generator=gen) * 0.9 + 0.1) * mask
            del mask
        else:
            img = walks_device(shape, nw, steps, seed, dev, isolated=iso, rings=rings)
        V_local = V
    else:
        # this rank's Z-slab: owned planes [z0, z1) plus `halo` planes on either side; the same
        # seeded point set on every rank, scattered into the slab only
        z0, z1, ze0, ze1 = slab.slab_bounds(shape[0], world, rank, args.halo)
        pts = walks_points(shape, nw, steps, seed, isolated=iso, rings=rings)
        plane = shape[1] * shape[2]
        pts = pts[(pts >= ze0 * plane) & (pts < ze1 * plane)] - ze0 * plane
        img = torch.zeros((ze1 - ze0) * plane, dtype=torch.bool, device=dev)
        img[torch.from_numpy(pts).to(dev)] = True
        img = img.reshape(ze1 - ze0, shape[1], shape[2])
        V_local = int(img.numel())
        del pts
    torch.cuda.synchronize()
    lib = _native.library()

    def step(image=None):
        image = img if image is None else image
        if is_stack:
            return batch.stack_device(image, spacing=spacing, device=dev)
        if world == 1:
            return _pipeline.run_native(image, spacing=spacing, device=dev)
        return slab.skeleton_and_summary_slab(image, global_shape=shape, z0=z0, z1=z1, ze0=ze0, spacing=spacing,
                                              device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    clocks = ClockSampler(local)
    clocks.__enter__()   # sampled from the warm-up on: the timed region alone is ~0.1 s
    for _ in range(max(1, args.warmup)):
        r = step()
    sync_all()
    if world == 1 or is_stack:
        g, p = r.graph, r.paths
        counts = dict(nodes=g.n_nodes, edges=g.n_edges, paths=p.n_paths, path_entries=p.n_entries,
                      junction_edges=g.n_junction_edges, mst_rounds=g.mst_rounds, rank_rounds=p.rank_rounds)
        if is_stack and world > 1:
            counts = {k: int(sum_over_ranks(v)) for k, v in counts.items() if k in ('nodes', 'edges', 'paths', 'path_entries')}
        del g, p
    else:
        counts = dict(nodes=r.n_nodes_total, edges=r.n_edges_total, paths=r.n_paths_total,
                      path_entries=r.n_entries_total, halo=args.halo, rank_rounds=r.rank_rounds,
                      table_records=r.n_table, table_rounds=r.table_rounds)
        if getattr(r, 'counts', None):   # native runner: exchanges through peer-mapped arenas, no collective library
            counts.update(runner='native (skb_run_slab)', host_syncs_per_step=r.counts['syncs'],
                          exchanges_per_step=r.counts['exchanges'], nccl_collectives_per_step=0)
        else:
            counts.update(runner='python (slab.py)', nccl_collectives_per_step=8)
    alg = _pipeline.algorithmic_bytes(V, esize, len(shape), counts['nodes'], counts['edges'], counts['paths'],
                                      counts['path_entries'], valued)
    del r

    # ---- timed region: K steps, CUDA events, max over ranks; pack kernel timed by its own events ----
    hook = {'skb_pack_count': []}
    _pipeline.EVENT_HOOK = hook
    _pipeline.TIME_SWEEP = True      # the native runner times the sweep kernel with its own CUDA events
    slab.TIME_SWEEP = True
    sweep_ns = []
    if world == 1 and not is_stack:
        counts.update(runner='native (skb_run_image); capacity hints = counts of the previous step + 1/16, read-backs '
                             'overlap the kernels they size, every count verified every step')
    launches0 = lib.kernel_launches()
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    sync_all()
    start.record()
    for _ in range(args.steps):
        r = step()
        if getattr(r, 'counts', None):
            sweep_ns.append(r.counts.get('sweep_ns', 0))
            if world == 1 and 'overlapped' in r.counts:
                counts.update(host_read_backs_per_step=r.counts['syncs'], overlapped_read_backs=r.counts['overlapped'],
                              relaunched_for_small_hint=r.counts['retaken'])
        del r
    stop.record()
    sync_all()
    clocks.__exit__()
    _pipeline.EVENT_HOOK = None
    _pipeline.TIME_SWEEP = False
    slab.TIME_SWEEP = False
    total_ms = max_over_ranks(start.elapsed_time(stop))
    launches = int(sum_over_ranks(lib.kernel_launches() - launches0))
    ms = total_ms / args.steps
    value = V / (ms * 1e-3) / 1e6
    if hook['skb_pack_count']:
        pack_ms = float(np.mean([a.elapsed_time(b) for a, b in hook['skb_pack_count']]))
    else:
        pack_ms = float(np.mean(sweep_ns)) * 1e-6
    pack_ms = max_over_ranks(pack_ms)
    peak, peak_src = load_peaks()
    achieved = V_local * esize / (pack_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': f'skb_pack_{_pipeline.default_pack_variant}_kernel<{esize}>',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': ncu_traffic(args.workload, f'skb_pack_{_pipeline.default_pack_variant}_kernel<{esize},')
                if world == 1 else None,
                'peak_source': peak_src, 'kernel_ms': pack_ms, 'kernel_share_of_step': pack_ms / ms,
                'algorithmic_bytes_per_launch': V_local * esize,
                'job': {'algorithmic_bytes': alg, 'achieved': alg / (ms * 1e-3) / 1e9,
                        'frac_of_n_gpus': alg / (ms * 1e-3) / 1e9 / (peak * world)}}

    if args.timeline and world > 1 and slab.NATIVE:
        sync_all()
        r = step()
        torch.cuda.synchronize()
        if rank == 0:
            for k, v in r.counts['host_us'].items():
                print(f'  [host timeline, rank 0] {k:24s} +{v / 1e3:7.3f} ms', file=sys.stderr)
            print(f'  [host timeline, rank 0] syncs {r.counts["syncs"]} exchanges {r.counts["exchanges"]}', file=sys.stderr)
        del r
    elif args.timeline and world > 1:
        sync_all()
        slab.TIMELINE = []
        t0 = time.perf_counter()
        r = step()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        tl, slab.TIMELINE = slab.TIMELINE, None
        if rank == 0:
            prev = t0
            for label, t in tl:
                print(f'  [timeline] {label:24s} +{(t - prev) * 1e3:7.3f} ms', file=sys.stderr)
                prev = t
            print(f'  [timeline] {"final sync":24s} +{(t1 - prev) * 1e3:7.3f} ms   total {(t1 - t0) * 1e3:.3f} ms', file=sys.stderr)
        del r

    # ---- optional per-entry-point table (one extra step, not part of any reported number) -----
    if args.stages:
        hook = {'*': True}
        _pipeline.EVENT_HOOK = hook
        _pipeline.TIME_SWEEP = 2
        r = step()
        sync_all()
        _pipeline.EVENT_HOOK = None
        _pipeline.TIME_SWEEP = False
        rows = [(k, sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in hook.items() if k != '*']
        if getattr(r, 'counts', None) and 'stage_ns' in r.counts:   # native runner: its own stage clock
            rows = [(k, v * 1e-6, 1) for k, v in r.counts['stage_ns'].items() if k != 'start']
            rows.append(('syncs', float(r.counts['syncs']), 0))
        if rank == 0:
            for k, t, n in sorted(rows, key=lambda x: -x[1]):
                print(f'  {k:28s} {t:9.3f} ms  ({n} calls)', file=sys.stderr)
        del r

    # ---- end to end through the public API from a pinned host image ---------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
        host.copy_(img)
        torch.cuda.synchronize()
        del img
        torch.cuda.empty_cache()
        n_e2e = max(2, min(args.steps, 5))
        d2h = 0
        times = []
        for i in range(1 + n_e2e):
            sync_all()
            t0 = time.perf_counter()
            if is_stack:
                df = batch.summarize_images(host, spacing=spacing, device=dev)
                d2h = int(df.memory_usage(index=False).sum())
                del df
            elif world == 1:
                sk = skan_b200.Skeleton(host, spacing=spacing, keep_images=False)
                df = skan_b200.summarize(sk, separator='_')
                d2h = int(df.memory_usage(index=False).sum())
                del sk, df
            else:
                r = step(host.to(dev, non_blocking=True))
                tabs = [t.cpu() for t in r.table] + [r.dist.cpu(), r.mean.cpu(), r.stdev.cpu()]
                d2h = int(sum(t.numel() * t.element_size() for t in tabs))
                del r, tabs
            sync_all()
            dt = time.perf_counter() - t0
            if i:
                times.append(dt)
        e2e_s = max_over_ranks(float(np.mean(times)))
        e2e = {'value': V / e2e_s / 1e6, 'unit': UNIT, 'h2d_bytes_per_step': int(sum_over_ranks(V_local * esize)),
               'd2h_bytes_per_step': int(sum_over_ranks(d2h)), 'ms_per_step': e2e_s * 1e3, 'steps': n_e2e}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': dtype_name, 'data': 'synthetic',
            'config': {'workload': args.workload, 'shape': list(shape), 'spacing': [float(x) for x in np.atleast_1d(spacing)],
                       'parallelism': 'single GPU' if world == 1 else
                       (f'{world} shares of the stack, no collective' if is_stack else f'{world} Z-slabs, halo {args.halo}'),
                       'l2': 'input larger than L2 (no flush needed)' if V_local * esize > 256e6 else 'input fits L2',
                       'pack_variant': _pipeline.default_pack_variant, **counts},
            'roofline': roofline, 'clocks': clocks.summary(), 'gpu_launches': int(launches),
            'e2e': e2e}
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        cb = cpu_baseline(args.workload)
        cb.pop('seconds', None)
        line['cpu_baseline'] = cb
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
