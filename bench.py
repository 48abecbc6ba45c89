#!/usr/bin/env python
"""bench.py -- skeleton Mvox/s of Skeleton + summarize on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W         # the reference's own CPU path (oracle/_ref)
    torchrun ... bench.py --gpus N ...                            # one rank per GPU

A "step" is one pass of the whole hot path (mask sweep -> pixel graph -> junction MST -> CSR ->
paths -> statistics -> branch table) over one synthetic skeleton volume.  `value` is measured
with the image already resident in HBM (CUDA events, max over ranks); `e2e` is the same job
through the public API (skan_b200.Skeleton + skan_b200.summarize) from a pinned HOST image,
H2D and D2H copies inside the timed region.  `parity` compares the CUDA path with the reference's
own CPU implementation (oracle/_ref; the NumPy port of oracle/ when that copy is absent) on a
bounded sample of the same workload (ints bit-exact, fp64 to 1e-12) and carries a
partition-independent digest of the FULL result, identical at 1, 2, 4 and 8 GPUs when the results
are.  Prints ONE JSON line on rank 0 and exits non-zero when a parity check fails.
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
# the bounded sample the CPU implementation runs on: leading planes / rows / images = shape[0] // SAMPLE_FRAC
SAMPLE_FRAC = {'volume_2048': 2, 'volume_1024': 1, 'image_16384': 1, 'volume_256': 1, 'batch_512': 64}
OTHER_WORKLOADS = ('volume_1024', 'image_16384', 'batch_512')


def ncu_traffic(workload, kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of the same workload (profiles/r0N_ncu_full_<workload>.csv, newest round first)."""
    import csv
    for rnd in ('r02', 'r01'):
        path = os.path.join(ROOT, 'profiles', f'{rnd}_ncu_full_{workload}.csv')
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


# ------------------------------------------------------------------------------------------------------------
# the CPU side: the reference itself (oracle/_ref, an unmodified copy made by oracle/make_ref.py) or, when that
# copy is absent, the NumPy/SciPy/numba port (oracle/skan_oracle.py).  Only used as the checker and as the
# reported baseline -- never on the product path.
# ------------------------------------------------------------------------------------------------------------
_cpu_impl = None


def cpu_impl():
    """-> (kind, run(img, spacing) -> (skeleton, summary columns), map_array_seconds(reset) or None)."""
    global _cpu_impl
    if _cpu_impl is not None:
        return _cpu_impl
    from skan_b200.synth import walks
    tiny = walks((32, 32, 32), 8, 24, seed=1)
    want = os.environ.get('SKAN_B200_CPU_IMPL', 'reference')
    if want == 'reference':
        try:
            from oracle import make_ref
            make_ref.make()
            csr = make_ref.import_reference()

            def run(img, spacing):
                sk = csr.Skeleton(img, spacing=spacing)
                return sk, csr.summarize(sk, separator='_')
            run(tiny, (0.5, 0.1, 0.1))     # numba JIT of the reference's loops
            run(walks((64, 64), 6, 48, seed=1), 1)
            _cpu_impl = ('reference', run, make_ref.map_array_seconds)
            return _cpu_impl
        except Exception as e:  # the copy is absent or does not import here: say so and use the port
            print(f'[bench] reference copy unavailable ({e}); timing the port instead', file=sys.stderr)
    from oracle import skan_oracle as orc

    def run(img, spacing):
        return orc.skeleton_and_summary(img, spacing=spacing)
    run(tiny, (0.5, 0.1, 0.1))
    _cpu_impl = ('port', run, None)
    return _cpu_impl


def sample_image(workload, frac=None):
    """The bounded sample of a workload as a host array: a volume of the leading shape[0] // frac planes (rows,
    images) with the walker / ring / isolated counts scaled alike.  Deterministic, so every rank can rebuild it."""
    from skan_b200.synth import walks, stack_walks_points
    shape, nw, steps, seed, iso, rings, spacing, valued = WORKLOADS[workload]
    frac = frac or SAMPLE_FRAC[workload]
    sub = (shape[0] // frac,) + tuple(shape[1:])
    if workload.startswith('batch'):
        img = np.zeros(int(np.prod(sub)), bool)
        img[stack_walks_points(sub[0], sub[1], sub[2], nw, steps, seed, isolated=iso)] = True
        return img.reshape(sub), sub
    img = walks(sub, max(8, nw // frac), steps, seed, isolated=iso // frac, rings=rings // frac)
    if valued:
        rng = np.random.default_rng(seed)
        img = (img * (rng.random(sub, dtype=np.float32) * 0.9 + 0.1)).astype(np.float32)
    return img, sub


def sample_points(workload, frac=None):
    """Raveled foreground indices of sample_image() for a boolean volume (a rank scatters its slab from them)."""
    from skan_b200.synth import walks_points
    shape, nw, steps, seed, iso, rings, spacing, valued = WORKLOADS[workload]
    frac = frac or SAMPLE_FRAC[workload]
    sub = (shape[0] // frac,) + tuple(shape[1:])
    return walks_points(sub, max(8, nw // frac), steps, seed, isolated=iso // frac, rings=rings // frac), sub


def cpu_run(workload, img, sub, keep=False):
    """Time the CPU implementation on one sample.  -> (baseline dict, (skeleton, summary) or None)."""
    kind, run, map_s = cpu_impl()
    shape, spacing = WORKLOADS[workload][0], WORKLOADS[workload][6]
    if map_s:
        map_s(True)
    t0 = time.perf_counter()
    if workload.startswith('batch'):
        n_paths, result = 0, None
        per_image = []
        for i in range(sub[0]):
            try:
                sk, summ = run(img[i], spacing)
                n_paths += sk.n_paths
                if keep:
                    per_image.append((sk, summ))
            except ValueError:
                if keep:
                    per_image.append(None)
        dt = time.perf_counter() - t0
        what = f'first {sub[0]} of {shape[0]} images of {workload}, one by one ({n_paths} paths)'
        result = per_image if keep else None
    else:
        sk, summ = run(img, spacing)
        dt = time.perf_counter() - t0
        what = (f'first {sub[0]} of {shape[0]} planes of {workload} ({img.size / 1e6:.0f} Mvox, {sk.graph.shape[0]} nodes, '
                f'{sk.n_paths} paths)')
        result = (sk, summ) if keep else None
    base = {'value': img.size / dt / 1e6, 'unit': UNIT, 'cores': 1, 'kind': kind,
            'sample': f'{what}, {dt:.2f} s, single thread like the reference', 'seconds': dt}
    if map_s:
        base['map_array_share'] = round(map_s(False) / dt, 3)
        base['note'] = ('unmodified reference source (oracle/_ref); skimage.util.map_array is a NumPy searchsorted '
                        'stand-in (scikit-image is not in this image), its share of the time is map_array_share')
    return base, result


def _pool_init():
    cpu_impl()


def _pool_one(args):
    img, spacing = args
    try:
        sk, _ = cpu_impl()[1](img, spacing)
        return sk.n_paths
    except ValueError:
        return 0


def cpu_pool_run(workload, img, sub):
    """The analogue of the reference's pipe.process_images (pipe.py:140-148): independent images over a pool of
    os.cpu_count() workers (capped at 64: every worker JIT-compiles the reference's numba loops first)."""
    import multiprocessing as mp
    spacing = WORKLOADS[workload][6]
    cores = max(1, min(os.cpu_count() or 1, 64, sub[0]))
    with mp.get_context('fork').Pool(cores, initializer=_pool_init) as pool:
        pool.map(_pool_one, [(img[0], spacing)] * cores)     # every worker warm
        t0 = time.perf_counter()
        n_paths = sum(pool.map(_pool_one, [(img[i], spacing) for i in range(sub[0])], chunksize=max(1, sub[0] // (4 * cores))))
        dt = time.perf_counter() - t0
    return {'value': img.size / dt / 1e6, 'unit': UNIT, 'cores': cores, 'seconds': dt,
            'sample': f'{sub[0]} images over multiprocessing.Pool({cores}) ({n_paths} paths), {dt:.2f} s'}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, each step
    a bounded sample of the workload.  The reference is single-threaded on this path (cores = 1); for the stack of
    independent images the multiprocessing.Pool figure is added."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    kind = cpu_impl()[0]
    img, sub = sample_image(args.workload, args.sample_frac)
    vals, t_all = [], time.perf_counter()
    for i in range(args.warmup + args.steps):
        r, _ = cpu_run(args.workload, img, sub)
        if i >= min(args.warmup, 1):       # the JIT warm-up is done in cpu_impl(); one extra sample is enough
            vals.append(r)
        if time.perf_counter() - t_all > 200:
            break
    v = float(np.median([x['value'] for x in vals]))
    ms = float(np.median([x['seconds'] for x in vals])) * 1e3
    shape = WORKLOADS[args.workload][0]
    cb = {k: vals[-1][k] for k in vals[-1] if k != 'seconds'}
    cb['value'] = v
    cb['best'] = float(max(x['value'] for x in vals))
    cb['repetitions'] = len(vals)
    if args.workload.startswith('batch'):
        cb['pool'] = cpu_pool_run(args.workload, img, sub)
    line = {'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(vals), 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32' if WORKLOADS[args.workload][7] else 'u8', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': args.workload, 'shape': list(shape), 'sample_shape': list(sub),
                       'cpu_implementation': kind, 'statistic': 'median over the repetitions'},
            'cpu_baseline': cb,
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# parity of a device result against the CPU implementation's result on the same input
# ------------------------------------------------------------------------------------------------------------
def _fp_columns(summ):
    col = {k.replace('-', '_'): np.asarray(v) for k, v in dict(summ).items()}
    return {k: col[k].astype(np.float64) for k in ('branch_distance', 'mean_pixel_value', 'stdev_pixel_value',
                                                    'euclidean_distance')}


def _fp_check(got, want):
    """max relative error of the per-path fp64 columns (tolerance 1e-12, BASELINE.json north_star); the standard
    deviation is compared through its square (sqrt of a difference of nearly equal numbers, csr.py:813-816)."""
    worst = 0.0
    for k, w in want.items():
        g = np.asarray(got[k], np.float64)
        if g.shape != w.shape:
            return float('inf')
        if k.startswith('stdev'):
            m2 = np.maximum(np.asarray(want['mean_pixel_value'])**2, 1e-300)
            err = np.abs(g**2 - w**2) / (m2 + w**2)
        else:
            err = np.abs(g - w) / np.maximum(np.abs(w), 1e-300)
        err = np.where((g == w) | (np.isnan(g) & np.isnan(w)), 0.0, err)
        if err.size:
            worst = max(worst, float(err.max()))
    return worst


def check_single(r, ref, nvox, what):
    """r: a single-GPU run (_pipeline.run_native); ref = (skeleton, summary) of the CPU implementation."""
    from skan_b200 import digest
    sk, summ = ref
    got, want = digest.digest_single(r), digest.digest_oracle(sk, summ)
    ints_ok = got['components'] == want['components'] and got['counts'] == want['counts']
    tf = r.table[2]
    fp = _fp_check(dict(branch_distance=r.dist.cpu().numpy(), mean_pixel_value=r.mean.cpu().numpy(),
                        stdev_pixel_value=r.stdev.cpu().numpy(), euclidean_distance=tf[-1].cpu().numpy()),
                   _fp_columns(summ))
    bad = [k for k in want['components'] if got['components'].get(k) != want['components'][k]]
    return {'ok': bool(ints_ok and fp <= 1e-12), 'checked_voxels': int(nvox), 'against': what,
            'int_arrays_bit_exact': bool(ints_ok), 'fp64_max_rel_err': fp, 'fp64_tolerance': 1e-12,
            'nodes': want['counts']['nodes'], 'paths': want['counts']['paths'], 'sample_digest': got['digest'],
            **({'mismatch': bad or ['counts']} if not ints_ok else {})}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='skan_b200')
    ap.add_argument('--workload', default='volume_2048', choices=sorted(WORKLOADS))
    ap.add_argument('--pack-variant', default=None, choices=['ldg', 'tma', 'fused'])
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU run (and with it the parity check)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-others', action='store_true', help='N=1 default run: skip the other BASELINE workloads')
    ap.add_argument('--sample-frac', type=int, default=None, help='CPU sample = leading 1/frac of the workload')
    ap.add_argument('--halo', type=int, default=2, help='halo planes of the Z-slab decomposition (N > 1, --halo-mode input)')
    ap.add_argument('--halo-mode', default='exchange', choices=['exchange', 'input'],
                    help='N > 1: exchange = disjoint Z-partition, boundary planes of the packed mask exchanged over NVLink; '
                         'input = every rank is handed its slab with the halo planes')
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
    # the host thread (launches, read-back mailbox, pinned buffers of the end-to-end leg) runs next to its GPU
    from skan_b200._numa import bind_thread_to_gpu
    placement = bind_thread_to_gpu(dev)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')  # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()
    import skan_b200
    from skan_b200 import _native, _pipeline, slab, batch, digest
    from skan_b200.synth import walks_device, walks_points, stack_walks_points

    if args.pack_variant:
        _pipeline.default_pack_variant = args.pack_variant
    lib = _native.library()
    peak, peak_src = load_peaks()
    xhalo = args.halo_mode == "exchange"

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

    def scatter_slab(pts, shp, z_lo, z_hi):
        """planes [z_lo, z_hi) of the boolean volume whose foreground voxels are `pts`, on the device"""
        plane = shp[1] * shp[2]
        sel = pts[(pts >= z_lo * plane) & (pts < z_hi * plane)] - z_lo * plane
        t = torch.zeros((z_hi - z_lo) * plane, dtype=torch.bool, device=dev)
        t[torch.from_numpy(sel).to(dev)] = True
        return t.reshape(z_hi - z_lo, shp[1], shp[2])

    def measure(workload, main_line, steps, warmup):
        """One workload: synthetic input resident in HBM, warm-up, K timed steps, roofline, parity."""
        shape, nw, wsteps, seed, iso, rings, spacing, valued = WORKLOADS[workload]
        V = int(np.prod(shape, dtype=np.int64))
        is_stack = workload.startswith('batch')
        if world > 1 and not is_stack and (len(shape) != 3 or valued):
            raise SystemExit('multi-GPU runs use the Z-slab decomposition of a 3-D volume (volume_* workloads) '
                             'or a stack of independent images (batch_512)')
        esize, dtype_name = (4, 'f32') if valued else (1, 'u8')
        z0 = z1 = ze0 = ze1 = 0
        if is_stack:
            b0, b1 = rank * shape[0] // world, (rank + 1) * shape[0] // world
            pts = stack_walks_points(b1 - b0, shape[1], shape[2], nw, wsteps, seed + rank, isolated=iso)
            img = torch.zeros((b1 - b0) * shape[1] * shape[2], dtype=torch.bool, device=dev)
            img[torch.from_numpy(pts).to(dev)] = True
            img = img.reshape(b1 - b0, shape[1], shape[2])
            del pts
        elif world == 1:
            if valued:
                mask = walks_device(shape, nw, wsteps, seed, dev, isolated=iso, rings=rings)
                gen = torch.Generator(device=dev).manual_seed(seed)
                img = (torch.rand(shape, device=dev, generator=gen) * 0.9 + 0.1) * mask
                del mask
            else:
                img = walks_device(shape, nw, wsteps, seed, dev, isolated=iso, rings=rings)
        else:
            # this rank's Z-slab: owned planes [z0, z1) plus `halo` planes on either side; the same seeded point
            # set on every rank, scattered into the slab only
            z0, z1, ze0, ze1 = slab.slab_bounds(shape[0], world, rank, 0 if xhalo else args.halo)
            img = scatter_slab(walks_points(shape, nw, wsteps, seed, isolated=iso, rings=rings), shape, ze0, ze1)
        V_local = int(img.numel())
        torch.cuda.synchronize()

        def step(image=None, bounds=None, gshape=None):
            image = img if image is None else image
            if is_stack:
                return batch.stack_device(image, spacing=spacing, device=dev)
            if world == 1:
                return _pipeline.run_native(image, spacing=spacing, device=dev)
            a0, a1, ae0 = bounds or (z0, z1, ze0)
            if xhalo:
                return slab.skeleton_and_summary_slab(image, global_shape=gshape or shape, z0=a0, z1=a1, spacing=spacing,
                                                      device=dev, halo='exchange')
            return slab.skeleton_and_summary_slab(image, global_shape=gshape or shape, z0=a0, z1=a1, ze0=ae0,
                                                  spacing=spacing, device=dev)

        clocks = ClockSampler(local) if main_line else None
        if clocks:
            clocks.__enter__()   # sampled from the warm-up on: the timed region alone is ~0.1 s
        # a cold image: no capacity hints from an image of the same kind (one step, after the allocator is warm)
        r = step()
        sync_all()
        cold_ms = None
        if world == 1:
            del r
            _pipeline.USE_HINTS = False
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                r = step()
                del r
            b.record()
            torch.cuda.synchronize()
            cold_ms = a.elapsed_time(b) / 3
            _pipeline.USE_HINTS = True
        for _ in range(max(3, warmup)):
            r = step()
        sync_all()
        if world == 1 or is_stack:
            g, p = r.graph, r.paths
            counts = dict(nodes=g.n_nodes, edges=g.n_edges, paths=p.n_paths, path_entries=p.n_entries,
                          junction_edges=g.n_junction_edges)
            if is_stack and world > 1:
                counts = {k: int(sum_over_ranks(v)) for k, v in counts.items() if k != 'junction_edges'}
            del g, p
        else:
            counts = dict(nodes=r.n_nodes_total, edges=r.n_edges_total, paths=r.n_paths_total,
                          path_entries=r.n_entries_total, table_records=r.n_table,
                          halo='2 planes of packed mask per face exchanged over NVLink (disjoint partition)' if xhalo
                          else f'{args.halo} planes per face passed with the slab')
            if getattr(r, 'counts', None):   # native runner: exchanges through peer-mapped arenas, no collective library
                counts.update(runner='native (skb_run_slab)', host_syncs_per_step=r.counts['syncs'],
                              exchanges_per_step=r.counts['exchanges'],
                              device_barriers_per_step=r.counts.get('device_barriers', 0), nccl_collectives_per_step=0)
            else:
                counts.update(runner='python (slab.py)', nccl_collectives_per_step=8)
        alg = _pipeline.algorithmic_bytes(V, esize, len(shape), counts['nodes'], counts['edges'], counts['paths'],
                                          counts['path_entries'], valued)
        # digest of the full result (partition independent: the same at every N when the results are the same)
        full = None
        if not is_stack:
            full = digest.digest_single(r) if world == 1 else digest.digest_slab(r)
        del r

        # ---- timed region: K steps, CUDA events (one per step), max over ranks ----
        _pipeline.TIME_SWEEP = True      # the native runner times the sweep kernel with its own CUDA events
        slab.TIME_SWEEP = True
        sweep_ns = []
        if world == 1 and not is_stack:
            counts.update(runner='native (skb_run_image); capacity hints = counts of the previous step + 1/16')
        launches0 = lib.kernel_launches()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        sync_all()
        marks[0].record()
        for i in range(steps):
            r = step()
            if getattr(r, 'counts', None):
                sweep_ns.append(r.counts.get('sweep_ns', 0))
                if world == 1 and 'overlapped' in r.counts:
                    counts.update(host_read_backs_per_step=r.counts['syncs'], overlapped_read_backs=r.counts['overlapped'],
                                  relaunched_for_small_hint=r.counts['retaken'])
            del r
            marks[i + 1].record()
        sync_all()
        if clocks:
            clocks.__exit__()
        _pipeline.TIME_SWEEP = False
        slab.TIME_SWEEP = False
        per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        total_ms = max_over_ranks(marks[0].elapsed_time(marks[steps]))
        launches = int(sum_over_ranks(lib.kernel_launches() - launches0))
        ms = total_ms / steps
        value = V / (ms * 1e-3) / 1e6
        pack_ms = max_over_ranks(float(np.mean(sweep_ns)) * 1e-6 if sweep_ns else 0.0)
        achieved = V_local * esize / (pack_ms * 1e-3) / 1e9 if pack_ms else 0.0
        kern = f'skb_pack_{_pipeline.default_pack_variant}_kernel<{esize}'
        roofline = {'bound': 'hbm', 'kernel': kern + '>', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': ncu_traffic(workload, kern + ',') if world == 1 else None,
                    'peak_source': peak_src, 'kernel_ms': pack_ms, 'kernel_share_of_step': pack_ms / ms,
                    'algorithmic_bytes_per_launch': V_local * esize,
                    'job': {'algorithmic_bytes': alg, 'achieved': alg / (ms * 1e-3) / 1e9,
                            'frac_of_n_gpus': alg / (ms * 1e-3) / 1e9 / (peak * world),
                            'frac_of_8tbs': alg / (ms * 1e-3) / 1e9 / (8000.0 * world)}}
        out = dict(workload=workload, V=V, V_local=V_local, esize=esize, dtype=dtype_name, shape=shape, spacing=spacing,
                   is_stack=is_stack, ms=ms, ms_median=max_over_ranks(float(np.median(per_step))),
                   ms_best=max_over_ranks(float(np.min(per_step))), value=value, counts=counts, roofline=roofline,
                   launches=launches, clocks=clocks.summary() if clocks else None, cold_ms=cold_ms, full_digest=full,
                   img=img, step=step, bounds=(z0, z1, ze0, ze1))
        return out

    def parity_of(m, with_baseline):
        """CPU implementation on the bounded sample (rank 0) against the CUDA path on the same sample."""
        workload = m['workload']
        shape, spacing = m['shape'], m['spacing']
        frac = args.sample_frac or SAMPLE_FRAC[workload]
        ref = base = None
        what = None
        host_img = None
        if rank == 0:
            host_img, sub = sample_image(workload, frac)
            base, ref = cpu_run(workload, host_img, sub, keep=True)
            what = f'{base["kind"]} on the ' + base['sample'].split(', ')[0]
        if m['is_stack']:
            if rank != 0:
                return None, None
            # per-image comparison of the stack path with the CPU results of the same images
            r = batch.stack_device(torch.from_numpy(host_img).to(dev), spacing=spacing, device=dev)
            df = batch.summarize_images(torch.from_numpy(host_img).to(dev), spacing=spacing, device=dev)
            ok, worst, n_img = True, 0.0, 0
            ids = df['image_id'].to_numpy()
            for i, rr in enumerate(ref):
                rows = df[ids == i]
                if rr is None:
                    ok &= len(rows) == 0
                    continue
                col = {k.replace('-', '_'): np.asarray(v) for k, v in dict(rr[1]).items()}
                n_img += 1
                if len(rows) != len(col['skeleton_id']):
                    ok = False
                    continue
                for k in ('skeleton_id', 'node_id_src', 'node_id_dst', 'branch_type', 'image_coord_src_0',
                          'image_coord_src_1', 'image_coord_dst_0', 'image_coord_dst_1'):
                    ok &= bool(np.array_equal(rows[k].to_numpy(), col[k]))
                worst = max(worst, _fp_check({k: rows[k].to_numpy() for k in _fp_columns(rr[1])}, _fp_columns(rr[1])))
            del r
            par = {'ok': bool(ok and worst <= 1e-12), 'checked_voxels': int(host_img.size), 'against': what,
                   'int_arrays_bit_exact': bool(ok), 'fp64_max_rel_err': worst, 'fp64_tolerance': 1e-12,
                   'images': n_img}
            return par, base
        if world == 1:
            dimg = torch.from_numpy(host_img).to(dev)
            r = m['step'](dimg)
            torch.cuda.synchronize()
            par = check_single(r, ref, host_img.size, what)
            del r, dimg
            return par, base
        # N > 1: the sample volume over the same Z-slab decomposition, compared through the digest
        pts, sub = sample_points(workload, frac)
        a0, a1, ae0, ae1 = slab.slab_bounds(sub[0], world, rank, 0 if xhalo else args.halo)
        simg = scatter_slab(pts, sub, ae0, ae1)
        r = m['step'](simg, (a0, a1, ae0), sub)
        got = digest.digest_slab(r)
        nr = r.n_regular
        cols = dict(branch_distance=r.dist, mean_pixel_value=r.mean, stdev_pixel_value=r.stdev,
                    euclidean_distance=r.table[2][-1])
        mine = {k: (v[:nr].cpu().numpy(), v[nr:].cpu().numpy()) for k, v in cols.items()}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        del r, simg
        if rank != 0:
            return None, None
        want = digest.digest_oracle(*ref)
        ints_ok = got['components'] == want['components'] and got['counts'] == want['counts']
        fp_got = {k: np.concatenate([g[k][0] for g in gathered] + [g[k][1] for g in gathered]) for k in cols}
        fp = _fp_check(fp_got, _fp_columns(ref[1]))
        par = {'ok': bool(ints_ok and fp <= 1e-12), 'checked_voxels': int(np.prod(sub)), 'against': what,
               'int_arrays_bit_exact': bool(ints_ok), 'fp64_max_rel_err': fp, 'fp64_tolerance': 1e-12,
               'nodes': want['counts']['nodes'], 'paths': want['counts']['paths'], 'sample_digest': got['digest']}
        return par, base

    m = measure(args.workload, True, args.steps, args.warmup)
    workload, V, V_local, esize = m['workload'], m['V'], m['V_local'], m['esize']
    shape, spacing, is_stack, ms = m['shape'], m['spacing'], m['is_stack'], m['ms']
    img, step = m['img'], m['step']

    # ---- N > 1: section clock of the slab runner (one extra step, not part of any reported number): device time of
    # every section of rank 0 goes into the JSON line (what limits the scaling is read from it), --timeline prints the
    # host clock of every rank as well
    sections = None
    if world > 1 and slab.NATIVE and not is_stack:
        sync_all()
        slab.TIME_SWEEP = 2
        r = step()
        torch.cuda.synchronize()
        slab.TIME_SWEEP = False
        tl = [None] * world
        dist.all_gather_object(tl, (r.counts.get('host_us', {}), r.counts.get('dev_us', {})))
        sections = {k: round(max(t[1].get(k, 0.0) for t in tl) / 1e3, 4) for k in r.counts.get('dev_us', {})}
        if rank == 0 and args.timeline:
            for k, v in r.counts['host_us'].items():
                print(f'  [host timeline, rank 0] {k:24s} +{v / 1e3:7.3f} ms   device {r.counts["dev_us"][k] / 1e3:7.3f} ms   '
                      f'(host, all ranks: ' + ' '.join(f'{t[0][k] / 1e3:.3f}' for t in tl) + ')', file=sys.stderr)
            print(f'  [host timeline, rank 0] syncs {r.counts["syncs"]} exchanges {r.counts["exchanges"]} '
                  f'device barriers {r.counts.get("device_barriers", 0)}', file=sys.stderr)
        del r

    # ---- optional stage clock of the native runner (one extra step, not part of any reported number) -----
    if args.stages:
        _pipeline.TIME_SWEEP = 2
        r = step()
        sync_all()
        _pipeline.TIME_SWEEP = False
        if getattr(r, 'counts', None) and 'stage_ns' in r.counts and rank == 0:
            rows = [(k, v * 1e-6) for k, v in r.counts['stage_ns'].items() if k != 'start']
            for k, t in sorted(rows, key=lambda x: -x[1]):
                print(f'  {k:28s} {t:9.3f} ms', file=sys.stderr)
        del r

    # ---- parity on the bounded sample (the CPU run doubles as cpu_baseline) ---------------------------------
    parity = cpu_base = None
    if not args.no_cpu_baseline:
        parity, cpu_base = parity_of(m, True)

    # ---- end to end through the public API from a pinned host image ---------------------------
    e2e = None
    if not args.no_e2e:
        # the pinned source buffer lives on the NUMA node of this rank's GPU (first touch by the thread bound there)
        host = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
        host.copy_(img)
        torch.cuda.synchronize()
        del img
        m['img'] = None
        torch.cuda.empty_cache()
        n_e2e = max(2, min(args.steps, 5))
        d2h = 0
        times = []
        z0, z1, ze0, ze1 = m['bounds']
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
               'd2h_bytes_per_step': int(sum_over_ranks(d2h)), 'ms_per_step': e2e_s * 1e3, 'steps': n_e2e,
               'host_buffer': f"pinned, rank 0 on NUMA node {placement['numa_node']} ({placement['cpus']} CPUs)"
               if placement['bound'] else 'pinned, no NUMA placement (' + placement.get('error', 'no node information') + ')'}
        del host
    m['img'] = m['step'] = None
    img = step = None
    torch.cuda.empty_cache()

    # ---- the other BASELINE workloads (one GPU, default run): time, share of the roofline, parity -----------
    others = None
    if world == 1 and not args.no_others and workload == 'volume_2048':
        others = {}
        for w in OTHER_WORKLOADS:
            o = measure(w, False, 10, 3)
            po, cbo = (None, None) if args.no_cpu_baseline else parity_of(o, False)
            others[w] = {'ms_per_step': o['ms'], 'ms_median': o['ms_median'], 'ms_best': o['ms_best'], 'value': o['value'],
                         'unit': UNIT, 'dtype': o['dtype'], 'cold_image_ms': o['cold_ms'],
                         'job_frac_of_measured_peak': o['roofline']['job']['frac_of_n_gpus'],
                         'sweep_frac_of_measured_peak': o['roofline']['frac'], 'nodes': o['counts']['nodes'],
                         'paths': o['counts']['paths'], 'parity': po,
                         'cpu': None if cbo is None else {k: cbo[k] for k in ('value', 'kind', 'cores') if k in cbo}}
            o['img'] = o['step'] = None
            del o
            torch.cuda.empty_cache()

    line = {'metric': METRIC, 'value': m['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'ms_median': m['ms_median'], 'ms_best': m['ms_best'],
            'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': m['dtype'], 'data': 'synthetic',
            'config': {'workload': workload, 'shape': list(shape), 'spacing': [float(x) for x in np.atleast_1d(spacing)],
                       'parallelism': 'single GPU' if world == 1 else
                       (f'{world} shares of the stack, no collective' if is_stack else f'{world} Z-slabs, halo {args.halo_mode}'),
                       'l2': 'input larger than L2 (no flush needed)' if V_local * esize > 256e6 else 'input fits L2',
                       'pack_variant': _pipeline.default_pack_variant, **m['counts']},
            'roofline': m['roofline'], 'clocks': m['clocks'], 'gpu_launches': int(m['launches']),
            'e2e': e2e}
    if sections:   # device ms per section of the slab runner, max over ranks (one extra, untimed step)
        line['slab_sections_ms'] = sections
    if m['cold_ms'] is not None:
        line['cold_image_ms'] = m['cold_ms']
    if parity is not None or m['full_digest'] is not None:
        fd = m['full_digest']
        line['parity'] = dict(parity or {}, digest=fd['digest'] if fd else None,
                              digest_of='the full result of this run: equal at every N when the results are equal',
                              digest_components=fd['components'] if fd else None, fp_sums=fd['fp_sums'] if fd else None)
    if others is not None:
        line['other_workloads'] = others
    if cpu_base is not None and rank == 0 and world == 1:
        cpu_base.pop('seconds', None)
        line['cpu_baseline'] = cpu_base
    failed = False
    if rank == 0:
        print(json.dumps(line))
        failed = bool(parity is not None and not parity.get('ok', False))
        if others:
            failed |= any(o['parity'] is not None and not o['parity']['ok'] for o in others.values())
        if failed:
            print('[bench] PARITY FAILURE: the CUDA path and the CPU implementation disagree', file=sys.stderr)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        sys.exit(3)


if __name__ == '__main__':
    main()
