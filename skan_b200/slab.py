"""One volume over several GPUs: Z-slab decomposition with a halo and path stitching.

The reference has no distributed layer; this module is the B200-native
answer to "a volume larger than one device" for the skeleton-to-graph hot path
(BASELINE.json configs[4]).  One process per GPU, ``torch.distributed``
for the plumbing (NCCL on GPUs; gloo + the host emulation harness in the CPU unit tests).

Decomposition
  * rank k owns planes [z0, z1) of the volume and receives them together with ``halo`` planes on
    either side (clipped at the volume ends): the *extended slab*;
  * mask sweep, node emission, raw neighbourhood masks and CSR emission run on the extended slab
    exactly as on a single GPU.  Raster-order node ids of consecutive slabs are consecutive, so a
    global node id is the local id plus a per-rank constant.  Two halo planes make the raw masks
    (and so the "junction" test, raw degree >= 3) exact for the owned voxels and for the adjacent
    plane of each neighbour;
  * junction MST: clusters of junction voxels cross slab boundaries freely, so every rank emits the
    junction-junction edges whose smaller endpoint it owns (with global ids), the edge lists are
    all-gathered in rank order (which is the global (a, b) order the tie-break needs) and every
    rank runs the same Boruvka on them; the edge set is ~3 % of the nodes, so this costs little;
  * path tracing: walk starts are numbered for the owned nodes and for the adjacent plane of each
    neighbour exactly as the neighbour numbers them (all chain voxels of slab-boundary planes are
    forced splitters), so a list that leaves the slab points at a *global* start id.  Every rank
    ranks its own lists, then the ranked starts of the boundary planes of all ranks (a few
    thousand records each) are all-gathered, ranked again (same computation on every rank) and
    substituted back;
  * every path is owned by the rank that holds its first node; per-path sums (length, sum and sum
    of squares of the values) travel with the ranking, so the branch table of a path is computed
    where it is owned, whatever ranks its voxels live on;
  * skeleton ids: one (first key, last key, minimum node) record per path is all-gathered, the
    contracted graph is labelled identically on every rank, each rank ranks the component minima
    that are its own nodes, and one small all-reduce combines the ids.

Collectives per volume: 4 all-gathers of a handful of integers, 1 all-gather of the junction edges,
1 of the boundary table, 1 of the path records, 1 all-reduce of one int32 per path.  No voxel
data crosses NVLink (the halo planes are part of the input partition).
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

from . import _native, _pipeline

AUX_BYTES = 64

# optional host timeline (bench.py --timeline): list of (label, perf_counter) appended at the points
# where the host has just waited for the device or for a collective
TIMELINE = None


def _mark(label):
    if TIMELINE is not None:
        import time
        TIMELINE.append((label, time.perf_counter()))


def slab_bounds(nz, world, rank, halo):
    """Owned planes [z0, z1) and extended planes [ze0, ze1) of `rank` for a volume of nz planes."""
    base, rem = divmod(nz, world)
    z0 = rank * base + min(rank, rem)
    z1 = z0 + base + (1 if rank < rem else 0)
    return z0, z1, max(0, z0 - halo), min(nz, z1 + halo)


class _CountChannel:
    """All-gather of a few host integers between the ranks of ONE node through POSIX shared memory.

    The sizes that drive the next allocation are already on the host when they have to be
    exchanged; sending them through NCCL costs a host->device copy, a collective launch and a
    device->host synchronisation (~100-150 us) for a few bytes.  A sequence-numbered, double
    buffered shared-memory table does it in a few microseconds.  Used only when every rank of the
    group runs on this node; otherwise _allgather_ints falls back to the process group."""
    MAXV = 15

    def __init__(self, group):
        from multiprocessing import shared_memory
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        nbytes = 2 * self.world * (1 + self.MAXV) * 8
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            self.shm.buf[:nbytes] = bytes(nbytes)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:  # only the creating rank owns the segment (python < 3.13 registers attachers too)
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, 'shared_memory')
            except Exception:
                pass
        self.arr = np.ndarray((2, self.world, 1 + self.MAXV), dtype=np.int64, buffer=self.shm.buf)
        self.seq = 0
        dist.barrier(group=group)
        import atexit
        atexit.register(self.close)

    def close(self):
        try:
            self.arr = None
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass

    def allgather(self, vals):
        self.seq += 1
        b = self.seq & 1
        row = self.arr[b, self.rank]
        row[1:1 + len(vals)] = vals
        row[0] = self.seq           # published after the values (x86 keeps store order)
        flags = self.arr[b, :, 0]
        spins = 0
        while not (flags == self.seq).all():
            spins += 1
            if spins > 50_000_000:
                raise RuntimeError('count exchange timed out: a rank of the slab decomposition died')
        return self.arr[b, :, 1:1 + len(vals)].copy()


_channels = {}


def _single_node(group):
    lw = os.environ.get('LOCAL_WORLD_SIZE')
    return lw is None or int(lw) == dist.get_world_size(group)


def _allgather_ints(vals, device, group):
    vals = [int(v) for v in vals]
    if os.environ.get('SKAN_B200_COUNT_CHANNEL', 'shm') == 'shm' and _single_node(group) and len(vals) <= _CountChannel.MAXV:
        key = id(group) if group is not None else 0
        if key not in _channels:
            _channels[key] = _CountChannel(group)
        return _channels[key].allgather(vals)
    t = torch.tensor(vals, dtype=torch.int64, device=device)
    world = dist.get_world_size(group)
    out = torch.empty((world, t.numel()), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out.view(-1), t, group=group)
    return out.cpu().numpy()


def _allgather_padded(t, count, max_count, group):
    """t: (count, ...) tensor -> (world, max_count, ...) tensor (rows beyond each rank's count are -1)."""
    world = dist.get_world_size(group)
    pad = torch.full((max_count,) + tuple(t.shape[1:]), -1, dtype=t.dtype, device=t.device)
    if count:
        pad[:count] = t[:count]
    out = torch.empty((world,) + tuple(pad.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    return out


class _Exchange:
    """One all-gather of several per-rank arrays of `cap` rows each.  The producing kernels write
    straight into the send buffer (padding rows included); after the gather one small kernel
    regroups the rank-major chunks into one contiguous array per field."""

    def __init__(self, ctx, world, cap, row_bytes, dtypes):
        self.ctx, self.world = ctx, world
        self.cap = cap = max(16, -(-int(cap) // 16) * 16)     # every field segment is a multiple of 16 bytes
        self.row_bytes, self.dtypes = row_bytes, dtypes
        self.offs, o = [], 0
        for rb in row_bytes:
            self.offs.append(o)
            o += cap * rb
        self.chunk = o
        self.send = ctx.empty(self.chunk, torch.uint8)
        self.recv = ctx.empty(world * self.chunk, torch.uint8)

    def field(self, f):
        """This rank's rows of field f, as a typed view of the send buffer."""
        rb = self.row_bytes[f]
        return self.send[self.offs[f]: self.offs[f] + self.cap * rb].view(self.dtypes[f])

    def gather(self, group):
        dist.all_gather_into_tensor(self.recv, self.send, group=group)
        outs, desc = [], [self.world * len(self.row_bytes)]
        base = self.recv.data_ptr()
        for f, rb in enumerate(self.row_bytes):
            seg = self.cap * rb
            out = self.ctx.empty(self.world * seg, torch.uint8)
            for k in range(self.world):
                desc += [base + k * self.chunk + self.offs[f], out.data_ptr() + k * seg, seg]
            outs.append(out.view(self.dtypes[f]))
        d = torch.tensor(desc, dtype=torch.int64)
        self.ctx.call('skb_copy_segments', d.data_ptr())
        return outs


# ---- native runner (skb_run_slab): one C call per rank and volume -------------------------------------
NATIVE = os.environ.get('SKAN_B200_SLAB_NATIVE', '1') != '0'   # 0: the per-stage python sequencing below
TIME_SWEEP = False                                             # bench.py: counts['sweep_ns'] of the mask sweep
_comms = {}     # group key -> dict(handle, bytes, lib)
_arenas = {}    # (device, stream) -> dict(work, out_bytes)
_wtabs = {}
_INITIAL_ARENAS = (1 << 22, 1 << 22)   # work, out bytes of the first call (tests shrink them to exercise the retries)
_INITIAL_EXCHANGE = None             # bytes of the exchange arena (None: SKAN_B200_EXCHANGE_MB, default 64)


def _get_comm(ctx, group, min_bytes=0):
    """This group's communicator (control block in shared memory + peer-mapped exchange arenas),
    created collectively on first use and again when a volume needs a larger exchange arena."""
    import ctypes
    key = id(group) if group is not None else 0
    cur = _comms.get(key)
    if cur is not None and cur['lib'] is ctx.lib and cur['bytes'] >= min_bytes:
        return cur
    if cur is not None:
        cur['lib'].call('skb_comm_destroy', cur['handle'])
        del _comms[key]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nbytes = max(int(min_bytes), _INITIAL_EXCHANGE or (int(os.environ.get('SKAN_B200_EXCHANGE_MB', '64')) << 20))
    name = [None]
    if rank == 0:
        import uuid
        name[0] = 'g' + uuid.uuid4().hex[:20]
    dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    handle = ctypes.c_void_p()
    ctx.lib.call('skb_comm_create', rank, world, name[0].encode(), nbytes, ctypes.byref(handle))
    cur = dict(handle=handle, bytes=nbytes, lib=ctx.lib)
    _comms[key] = cur
    import atexit
    atexit.register(_close_comm, key)
    return cur


def _close_comm(key):
    cur = _comms.pop(key, None)
    if cur is not None:
        try:
            cur['lib'].call('skb_comm_destroy', cur['handle'])
        except Exception:
            pass


def _world_rank(group):
    """(world size, rank) of a process group.  Asked every time (a few microseconds): remembering the group object in a
    module global keeps it alive into interpreter shutdown, where its destructor races the backend's threads
    ("terminate called without an active exception" in one rank of a finished run), and remembering only its identity
    goes stale when a group is destroyed and another one created."""
    return dist.get_world_size(group), dist.get_rank(group)


def _slab_native(slab_image, *, global_shape, z0, z1, ze0, spacing, device, group, pack_variant, halo_exchange=False,
                 value_is_height=False):
    """skeleton_and_summary_slab through skb_run_slab: the whole per-rank sequence in one native call."""
    import ctypes
    world, rank = _world_rank(group)
    ctx = _pipeline._Ctx(device)
    img, code, shape, np_dtype = _pipeline._as_device_image(slab_image, ctx.device)
    if len(shape) != 3:
        raise ValueError('slab mode is for 3-D volumes')
    is_bool = np_dtype == np.dtype(bool)
    if value_is_height and is_bool:
        raise TypeError('value_is_height needs a numeric (non-boolean) skeleton image')
    if value_is_height and np_dtype.kind == 'u':
        raise ValueError('value_is_height with an unsigned integer image wraps the height differences '
                         '(reference csr.py:1024); cast the image to a signed or floating dtype')
    sp_arr = np.asarray(spacing)
    wkey = (str(ctx.device), id(ctx.lib), sp_arr.dtype.str, tuple(np.atleast_1d(sp_arr).tolist()))
    if wkey not in _wtabs:
        _wtabs[wkey] = torch.from_numpy(_pipeline.neighbour_weights(3, spacing)).to(ctx.device)
    sp = sp_arr if sp_arr.ndim else np.full(3, sp_arr)
    sp_is_int = sp.dtype.kind in 'iub'
    prm = _native.SlabParams()
    prm.image, prm.dtype = img.data_ptr(), code
    prm.pack_variant = min(1, _pipeline.PACK_VARIANT[pack_variant or _pipeline.default_pack_variant])
    for i in range(3):
        prm.global_shape[i] = int(global_shape[i])
        prm.spacing_f[i] = float(sp[i])
        prm.spacing_i[i] = int(sp[i]) if sp_is_int else 1
    prm.z0, prm.z1, prm.ze0, prm.nz_ext = int(z0), int(z1), int(ze0), int(shape[0])
    if halo_exchange:   # the image holds the owned planes only: the extended slab adds two planes towards every neighbour
        prm.halo_exchange = 1
        prm.ze0 = int(z0) - (2 if z0 > 0 else 0)
        prm.nz_ext = (int(z1) + (2 if z1 < int(global_shape[0]) else 0)) - prm.ze0
    prm.height = int(bool(value_is_height))
    prm.want_values, prm.time_sweep = int(not is_bool), int(TIME_SWEEP)   # 0, 1 (sweep kernel) or 2 (device section clock)
    prm.wtab, prm.spacing_is_int, prm.stream = _wtabs[wkey].data_ptr(), int(sp_is_int), ctx.stream
    V = int(prm.nz_ext) * int(shape[1]) * int(shape[2])
    akey = (str(ctx.device), ctx.stream, id(ctx.lib))
    cache = _arenas.setdefault(akey, dict(work=None, out_bytes=_INITIAL_ARENAS[1]))
    if cache['work'] is None:
        cache['work'] = ctx.empty(max(_INITIAL_ARENAS[0], V // 4 + _INITIAL_ARENAS[0]), torch.uint8)
    comm = _get_comm(ctx, group)
    res = _native.SlabResult()
    for _ in range(12):
        work = cache['work']
        out = ctx.empty(cache['out_bytes'], torch.uint8)
        prm.work, prm.work_bytes = work.data_ptr(), work.numel()
        prm.out, prm.out_bytes = out.data_ptr(), out.numel()
        if _pipeline.DEBUG_FILL is not None:
            work.fill_(_pipeline.DEBUG_FILL)
            out.fill_(_pipeline.DEBUG_FILL)
            _pipeline.LAST_ARENAS = (work, out)
        failure = None
        try:
            with ctx.guard():
                rc = ctx.lib.call('skb_run_slab', ctypes.byref(prm), comm['handle'], ctypes.byref(res))
        except _native.NativeError as e:   # the peers see the abort word and come here too: tell them, then raise
            failure, rc = e, -1
        if rc == 0:
            break
        # rare: some rank needs larger arenas.  The ranks agree on what happened through the process group
        # (they no longer agree on the exchange sequence), then run again.
        rcs = [None] * world
        dist.all_gather_object(rcs, (int(rc), int(res.exchange_need)), group=group)
        if failure is not None:
            raise failure
        if any(r[0] < 0 for r in rcs):
            raise _native.NativeError('skb_run_slab failed on rank ' + str([k for k, r in enumerate(rcs) if r[0] < 0]))
        if rc == 1:
            if res.work_need > work.numel():
                cache['work'] = ctx.empty(int(res.work_need), torch.uint8)
            if res.out_need > out.numel():
                cache['out_bytes'] = int(res.out_need)
        if any(r[0] == 2 for r in rcs):
            comm = _get_comm(ctx, group, min_bytes=max(max(r[1] for r in rcs), 2 * comm['bytes']))
        else:
            dist.barrier(group=group)
            ctx.lib.call('skb_comm_reset', comm['handle'])
            dist.barrier(group=group)
    else:
        raise _native.NativeError('skb_run_slab kept asking for larger arenas')
    cache['out_bytes'] = max(1 << 20, int(res.out_need * 1.05) + (1 << 16))
    counts = res.counts
    C = {k: int(counts[i]) for k, i in _native.SLAB_COUNT.items()}
    if prm.time_sweep:   # host / device clocks of the sections (bench.py --timeline)
        prev = 0
        C['host_us'] = {}
        for i, k in enumerate(_native.SLAB_SECTIONS):
            C['host_us'][k] = (int(res.host_ns[i]) - prev) / 1e3
            prev = int(res.host_ns[i])
        C['dev_us'] = {k: int(res.dev_ns[i]) / 1e3 for i, k in enumerate(_native.SLAB_SECTIONS)}
    if C['paths_total'] == 0:
        raise ValueError('index pointer size 0 should be 1')
    return _SlabResult(ctx, out, res.off, C, rank, world, tuple(shape), np_dtype, img, sp_is_int, int(prm.height))


class _SlabGraph:
    """Graph part of a native slab run (rows of the extended slab, local ids): views created on first access."""

    def __init__(self, v, ctx, shape, np_dtype, n, e, keepalive):
        self._v, self.ctx, self.shape, self.ndim, self.np_dtype = v, ctx, shape, 3, np_dtype
        self.n_nodes, self.n_edges, self._keepalive = n, e, keepalive

    lin = property(lambda self: self._v.lin)
    coords = property(lambda self: self._v.coords)
    values = property(lambda self: self._v.values)
    degrees = property(lambda self: self._v.degrees)
    indptr = property(lambda self: self._v.indptr)
    indices = property(lambda self: self._v.indices)
    data = property(lambda self: self._v.data)


class _SlabResult:
    """This rank's share of a native slab run.  The arrays are typed views of the output arena, created when they are
    first read: a caller that only looks at the branch table does not pay for the rest (the interpreter time between
    two volumes is GPU idle time)."""

    def __init__(self, ctx, out, off, C, rank, world, shape, np_dtype, img, sp_is_int, height=0):
        Ne, E, Pk, L = C['nodes_ext'], C['edges_ext'], C['regular'] + C['cycles'], C['entries_total']
        Pk1 = max(Pk, 1)
        i32, i64, f64 = torch.int32, torch.int64, torch.float64
        O = _native.SLAB_OFF
        spec = {
            'lin': (int(off[O['lin']]), Ne, i64, None), 'coords': (int(off[O['coords']]), Ne * 3, i64, (Ne, 3)),
            'values': (int(off[O['values']]), Ne, f64, None), 'degrees': (int(off[O['degrees']]), Ne, i32, None),
            'indptr': (int(off[O['indptr']]), Ne + 1, i32, None), 'indices': (int(off[O['indices']]), E, i32, None),
            'data': (int(off[O['data']]), E, f64, None), 'plen': (int(off[O['plen']]), Pk, i32, None),
            'pidx': (int(off[O['pidx']]), L, i32, None), 'pdata': (int(off[O['pdata']]), L, f64, None),
            'pw': (int(off[O['pw']]), L, f64, None), 'dist': (int(off[O['dist']]), Pk, f64, None),
            'mean': (int(off[O['mean']]), Pk, f64, None), 'stdev': (int(off[O['stdev']]), Pk, f64, None),
            't32': (int(off[O['t32']]), 3 * Pk, i32, (3, Pk)), 't64': (int(off[O['t64']]), 7 * Pk, i64, (7, Pk)),
            'tf': (int(off[O['tf']]), (7 + 2 * height) * Pk, f64, (7 + 2 * height, Pk)),
        }
        self._v = v = _pipeline._ArenaViews(out, spec)
        self.graph = _SlabGraph(v, ctx, shape, np_dtype, Ne, E, (img, out))
        self.rank, self.world, self.counts, self.spacing_is_int, self.value_is_height = rank, world, C, sp_is_int, bool(height)
        self.own = (C['own0'], C['own1'])
        self.id_shift = C['id_shift']
        self.n_nodes = C['own1'] - C['own0']
        self.n_edges, self.node_offset, self.edge_offset = C['edges_owned'], C['node_offset'], C['edge_offset']
        self.n_nodes_total, self.n_edges_total = C['nodes_total'], C['edges_total']
        self.n_regular, self.n_cycles, self.n_paths_total, self.n_entries_total = C['regular'], C['cycles'], C['paths_total'], L
        self.n_regular_total, self.regular_offset, self.cycle_offset = C['regular_total'], C['regular_offset'], C['cycle_offset']
        self.entry_offset_regular, self.entry_offset_cycles = C['entry_off_regular'], C['entry_off_cycles']
        self.rank_rounds = self.table_rounds = 0
        self.n_starts, self.n_table = C['starts'], C['table_records']

    path_len = property(lambda self: self._v.plen)
    paths_indices = property(lambda self: self._v.pidx)
    paths_data = property(lambda self: self._v.pdata)
    paths_weights = property(lambda self: self._v.pw)
    dist = property(lambda self: self._v.dist)
    mean = property(lambda self: self._v.mean)
    stdev = property(lambda self: self._v.stdev)
    table = property(lambda self: (self._v.t32, self._v.t64, self._v.tf))


def _exchange_image_halo(slab_image, z0, z1, NZ, device, group):
    """Owned planes [z0, z1) of rank k -> (extended slab with two planes of either neighbour, its first plane).
    Ranks are ordered along Z (rank k-1 holds the planes below z0, rank k+1 those from z1 on), as everywhere in this
    module; the planes travel as bytes through torch.distributed point-to-point operations (NCCL on GPUs, gloo on the
    CPU), so every image dtype is served."""
    world, rank = _world_rank(group)
    t = slab_image if isinstance(slab_image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(slab_image))
    t = t.to(device).contiguous()
    dt = t.dtype
    nz, NY, NX = (int(x) for x in t.shape)
    b = t.reshape(nz, -1).view(torch.uint8)
    has_lo, has_hi = z0 > 0 and rank > 0, z1 < NZ and rank + 1 < world
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    lo = torch.empty((2, b.shape[1]), dtype=torch.uint8, device=b.device) if has_lo else None
    hi = torch.empty((2, b.shape[1]), dtype=torch.uint8, device=b.device) if has_hi else None
    ops = []
    if has_lo:
        ops += [dist.P2POp(dist.isend, b[:2].contiguous(), peer(rank - 1), group),
                dist.P2POp(dist.irecv, lo, peer(rank - 1), group)]
    if has_hi:
        ops += [dist.P2POp(dist.isend, b[-2:].contiguous(), peer(rank + 1), group),
                dist.P2POp(dist.irecv, hi, peer(rank + 1), group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    ext = torch.cat([x for x in (lo, b, hi) if x is not None]).view(dt).reshape(-1, NY, NX)
    return ext, z0 - (2 if has_lo else 0)


def skeleton_and_summary_slab(slab_image, *, global_shape, z0, z1, ze0=None, spacing=1, device=None, group=None,
                              pack_variant=None, halo='input', value_is_height=False):
    """Skeleton + summarize of one Z-slab of a 3-D volume, collectively over the process group.

    halo='input' (default): slab_image = planes [ze0, ze0 + slab_image.shape[0]) of the volume (owned planes
    [z0, z1) plus at least two halo planes towards every neighbour): the caller partitions with overlap.
    halo='exchange': slab_image = the owned planes [z0, z1) only, a disjoint partition of the volume.  For boolean
    volumes whose planes are a multiple of 16384 voxels (native runner) the ranks exchange the packed mask of their two
    boundary planes through the peer-mapped exchange arenas (NVLink), 1 bit per voxel; otherwise (valued images,
    value_is_height, other plane sizes, several nodes) the two boundary planes of the image travel through the process
    group and the extended slab goes the halo='input' way.
    value_is_height (Skeleton's argument, csr.py:148-151, 932-948; numeric images, native runner): edge
    weights are hypot(height difference, distance) and the pixel value is an extra coordinate of the branch table; the
    value of a path's far end travels with the ranking records like its coordinates do.
    Returns a namespace with this rank's share of the results (device tensors):
    graph rows of the owned nodes with global column ids, global path arrays holding the entries
    of this rank's voxels, and the branch table of the paths whose first node is owned here.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    NZ, NY, NX = (int(x) for x in global_shape)
    if len(slab_image.shape) != 3:
        raise ValueError('slab mode is for 3-D volumes')
    if halo not in ('input', 'exchange'):
        raise ValueError("halo must be 'input' or 'exchange'")
    if halo == 'exchange':
        if int(slab_image.shape[0]) != z1 - z0:
            raise ValueError("halo='exchange' takes the owned planes [z0, z1) only")
        if z1 - z0 < 2:
            raise ValueError('bad slab bounds (each rank needs at least two owned planes)')
        if device is None:
            device = slab_image.device if isinstance(slab_image, torch.Tensor) and slab_image.device.type == 'cuda' else None
        if device is None:
            from .csr import _default_device
            device = _default_device()
        is_bool = (slab_image.dtype == torch.bool) if isinstance(slab_image, torch.Tensor) else (np.asarray(slab_image).dtype == bool)
        if (NATIVE and _single_node(group) and world <= 16 and is_bool and not value_is_height
                and (NY * NX) % 16384 == 0):   # 16384 = voxels per mask tile (skb_tile_voxels)
            # the packed mask of the boundary planes travels through the peer-mapped arenas inside the native runner
            return _slab_native(slab_image, global_shape=global_shape, z0=z0, z1=z1, ze0=z0, spacing=spacing, device=device,
                                group=group, pack_variant=pack_variant, halo_exchange=True)
        # everything else (valued images, value_is_height, planes that are not whole mask tiles, ranks on several nodes):
        # the two boundary planes of the IMAGE are exchanged with the neighbours through the process group, then the
        # extended slab goes the halo='input' way
        slab_image, ze0 = _exchange_image_halo(slab_image, z0, z1, NZ, device, group)
        halo = 'input'
    if ze0 is None:
        raise ValueError("halo='input' needs ze0, the first plane of slab_image")
    ze1 = ze0 + int(slab_image.shape[0])
    if not (0 <= ze0 <= z0 < z1 <= ze1 <= NZ) or z1 - z0 < 2:
        raise ValueError('bad slab bounds (each rank needs at least two owned planes)')
    has_lo, has_hi = z0 > 0, z1 < NZ
    if (has_lo and z0 - ze0 < 2) or (has_hi and ze1 - z1 < 2):
        raise ValueError('slab mode needs a halo of at least 2 planes')
    sz = NY * NX
    if device is None:
        device = slab_image.device if isinstance(slab_image, torch.Tensor) and slab_image.device.type == 'cuda' else None
    if device is None:
        from .csr import _default_device
        device = _default_device()
    if NATIVE and _single_node(group) and world <= 16:
        return _slab_native(slab_image, global_shape=global_shape, z0=z0, z1=z1, ze0=ze0, spacing=spacing,
                            device=device, group=group, pack_variant=pack_variant, value_is_height=value_is_height)
    if value_is_height:
        raise NotImplementedError('value_is_height in slab mode needs the native runner (all ranks on one node)')
    # ---- voxel and node stages on the extended slab ---------------------------------------------
    Vloc = (ze1 - ze0) * sz
    p_own0, p_own1 = (z0 - ze0) * sz, (z1 - ze0) * sz
    positions = [p_own0, p_own1, p_own0 + sz, p_own1 - sz,
                 p_own0 - sz if has_lo else p_own0, p_own1 + sz if has_hi else p_own1]
    g = _pipeline.graph_nodes(slab_image, spacing=spacing, device=device, pack_variant=pack_variant, z_origin=ze0,
                              rank_positions=positions)
    ctx = g.ctx
    Ne = g.n_nodes
    _mark('graph_nodes')

    def ranks_at(positions):
        pos = [min(int(p), Vloc) for p in positions]
        inside = [p for p in pos if p < Vloc]
        out = {}
        if inside and Ne:
            pt = torch.tensor(inside, dtype=torch.int64, device=ctx.device)
            ot = ctx.empty(len(inside), torch.int32)
            ctx.call('skb_rank_query', g.bits, g.pre256, pt, len(inside), ot)
            out = dict(zip(inside, ot.tolist()))
        return [Ne if p >= Vloc else (out[p] if Ne else 0) for p in pos]

    own0, own1, f1, l0, hl0, hu1 = g.position_ranks if g.position_ranks is not None else ranks_at(positions)
    N_k = own1 - own0
    _mark('rank_queries')
    # ---- junction MST on the junction edges of the whole volume -------------------------------------
    # every rank emits the edges whose smaller endpoint it owns (global ids), all ranks gather them
    # in rank order (= global (a, b) order, the MST's tie-break) and solve the same problem
    jcount, nej_k = _pipeline.junction_edge_counts(g, own0, own1)
    _mark('junction_edges')
    counts0 = _allgather_ints([N_k, nej_k, own0], ctx.device, group)               # count exchange 1
    Noff = np.concatenate([[0], np.cumsum(counts0[:, 0])])
    N_total = int(Noff[-1])
    if N_total >= 2**31:
        raise ValueError("more than 2^31 nodes: node ids no longer fit the reference's int32 index arrays")
    id_shift = int(Noff[rank]) - own0
    _mark('allgather_counts0')
    max_e = int(counts0[:, 1].max())
    if max_e:
        ex = _Exchange(ctx, world, max_e, [4, 4, 8, 1], [torch.int32, torch.int32, torch.int64, torch.uint8])
        ex.field(0).fill_(-1)                      # a < 0 marks the padding rows
        if nej_k:
            _pipeline.junction_edge_emit(g, jcount, ex.field(0), ex.field(1), ex.field(3), ex.field(2))
        gea, geb, gew, gek = ex.gather(group)                                     # collective 1
        # local ids of rank k -> global ids; the row order is the global (a, b) order the tie-break needs
        shifts = torch.tensor([int(Noff[k]) - int(counts0[k, 2]) for k in range(world)], dtype=torch.int64)
        ctx.call('skb_slab_edge_shift', gea, geb, world * ex.cap, ex.cap, shifts.data_ptr(), world)
    else:
        gea = geb = ctx.empty(0, torch.int32)
        gew, gek = ctx.empty(0, torch.int64), ctx.empty(0, torch.uint8)
    tree, mst_rounds = _pipeline.junction_mst(ctx, gea, geb, gew, N_total)
    # ---- CSR + node records + start numbering (one pass) ------------------------------------------
    slab = torch.zeros(13, dtype=torch.int64)
    slab[0], slab[1], slab[2], slab[3] = own0, own1, hl0, hu1
    slab[4], slab[5] = (hl0, f1) if has_lo else (0, 0)
    slab[6], slab[7] = (l0, hu1) if has_hi else (0, 0)
    slab[12] = ze0 * sz
    _pipeline.graph_csr(g, gea, geb, gek, tree, id_shift, records=True, slab=slab.data_ptr(), cc_init=False)
    _mark('mst_csr')
    indptr, indices, gdata, values, lin = g.indptr, g.indices, g.data, g.values, g.lin
    rec, startoff, _, _, is_min = g.path_state
    sel_idx = torch.tensor([own0, own1, f1, l0], dtype=torch.int64, device=ctx.device)
    rb = torch.cat([startoff.index_select(0, sel_idx), indptr.index_select(0, sel_idx[:2])]).tolist()
    so, ei = rb[:4], rb[4:6]
    S_k, SF, SL0 = so[1] - so[0], so[2] - so[0], so[3] - so[0]
    E_k = ei[1] - ei[0]
    _mark('records')
    counts = _allgather_ints([N_k, E_k, S_k, SF, SL0], ctx.device, group)        # collective 3
    Eoff = np.concatenate([[0], np.cumsum(counts[:, 1])])
    Soff = np.concatenate([[0], np.cumsum(counts[:, 2])])
    S_total = int(Soff[-1])
    start_lo, start_hi = int(Soff[rank]), int(Soff[rank + 1])
    _mark('allgather_counts1')
    slab[8], slab[9], slab[10], slab[11] = id_shift, start_lo - so[0], start_lo, start_hi
    # ---- phase 0: lists between terminals, first inside the slab ---------------------------------
    S = S_k
    su = ctx.empty(max(S, 1), torch.int32)
    se = ctx.empty(max(S, 1), torch.int32)
    buf = [ctx.empty((max(S, 1), 4), torch.int32), ctx.empty((max(S, 1), 4), torch.int32)]
    aux = [ctx.empty((max(S, 1), AUX_BYTES), torch.uint8), ctx.empty((max(S, 1), AUX_BYTES), torch.uint8)]
    stamp = torch.full((max(Ne, 1),), -1, dtype=torch.int32, device=ctx.device)
    flag = ctx.empty(4, torch.int32)
    ctx.call('skb_path_walk', indptr, indices, gdata, values, lin, rec, startoff, slab.data_ptr(), Ne, S, su, se,
             buf[0], aux[0], stamp)
    r = ctx.call('skb_path_rank', buf[0], buf[1], aux[0], aux[1], S, start_lo, start_hi, flag)
    ranked, raux = buf[r & 1], aux[r & 1]
    rank_rounds = r >> 1
    _mark('walk_rank(launch)')
    # ---- inter-slab table --------------------------------------------------------------------------
    nblk = counts[:, 3] + (counts[:, 2] - counts[:, 4])
    max_blk = int(nblk.max())
    table_rounds = 0
    cross_cycle = None
    if world > 1 and max_blk > 0:
        my_blk = int(nblk[rank])
        ex = _Exchange(ctx, world, max_blk, [16, AUX_BYTES], [torch.int32, torch.uint8])
        ctx.call('skb_slab_export', ranked, raux, SF, SL0, my_blk, ex.cap, start_lo, start_hi, ex.field(0), ex.field(1))
        T, TA = ex.gather(group)                                                  # collective 2
        n_table = world * ex.cap                                                  # padded layout: rank k at k * cap
        desc = torch.tensor([world] + [int(x) for x in Soff] + [int(x) for x in counts[:, 3]] +
                            [int(x) for x in counts[:, 4]] + [k * ex.cap for k in range(world)], dtype=torch.int64)
        T2, TA2 = torch.empty_like(T), torch.empty_like(TA)
        r2 = ctx.call('skb_slab_table_rank', T, T2, TA, TA2, n_table, desc.data_ptr(), flag)
        Tf, TAf = (T, TA) if (r2 & 1) == 0 else (T2, TA2)
        table_rounds = r2 >> 1
        ctx.call('skb_slab_apply', ranked, raux, Tf, TAf, S, desc.data_ptr(), start_lo, start_hi)
        cross_cycle = (Tf.view(-1, 4)[:, 0] >= 0).any()
    _mark('table')
    # ---- covered chain voxels, heads of the regular paths ------------------------------------------
    covered = ctx.empty(max(Ne, 1), torch.uint8)
    n_unc_t = ctx.empty(64, torch.int32)
    ctx.call('skb_slab_covered', rec, startoff, stamp, ranked, Ne, own0, own1, covered, n_unc_t)
    sel = ctx.empty(S + 1, torch.int32)
    ctx.call('skb_path_select', su, rec, ranked, S, 0, start_lo, sel, ctx.scratch(S))
    P0, n_uncovered = torch.stack([sel[S], n_unc_t.sum(dtype=torch.int32)]).tolist()
    _mark('covered_select')
    # ---- phase 1: junction-free cycles inside the slab -----------------------------------------------
    cyc = None
    Pc = 0
    if n_uncovered > 0:
        par = ctx.empty(max(Ne, 1), torch.int32)
        ctx.call('skb_cycle_roots', rec, covered, Ne, par, is_min)
        startoff_b = ctx.empty(Ne + 1, torch.int32)
        ctx.call('skb_path_start_count', rec, covered, slab.data_ptr(), Ne, startoff_b, ctx.scratch(Ne))
        sob = startoff_b.index_select(0, sel_idx[:2]).tolist()
        Sb = sob[1] - sob[0]
        if Sb:
            slab_b = slab.clone()
            slab_b[9], slab_b[10], slab_b[11] = -sob[0], 0, Sb
            sub, seb = ctx.empty(Sb, torch.int32), ctx.empty(Sb, torch.int32)
            bufb = [ctx.empty((Sb, 4), torch.int32), ctx.empty((Sb, 4), torch.int32)]
            auxb = [ctx.empty((Sb, AUX_BYTES), torch.uint8), ctx.empty((Sb, AUX_BYTES), torch.uint8)]
            ctx.call('skb_path_walk', indptr, indices, gdata, values, lin, rec, startoff_b, slab_b.data_ptr(), Ne, Sb,
                     sub, seb, bufb[0], auxb[0], None)
            rb = ctx.call('skb_path_rank', bufb[0], bufb[1], auxb[0], auxb[1], Sb, 0, Sb, flag)
            selb = ctx.empty(Sb + 1, torch.int32)
            ctx.call('skb_path_select', sub, rec, bufb[rb & 1], Sb, 1, 0, selb, ctx.scratch(Sb))
            Pc = int(selb[Sb].item())
            cyc = SimpleNamespace(S=Sb, su=sub, se=seb, ranked=bufb[rb & 1], aux=auxb[rb & 1], sel=selb)
    Pk = P0 + Pc
    _mark('cycles')
    start_node = ctx.empty(max(Pk, 1), torch.int32)
    plen = ctx.empty(Pk + 1, torch.int32)
    pmin = ctx.empty(max(Pk, 1), torch.int32)
    head_end = ctx.empty(max(Pk, 1), torch.int32)
    head_aux = ctx.empty((max(Pk, 1), AUX_BYTES), torch.uint8)
    ctx.call('skb_path_heads', su, rec, ranked, raux, sel, S, 0, 0, start_lo, id_shift, start_node, plen, pmin, None,
             head_end, head_aux)
    einfo_b = None
    if Pc:
        einfo_b = ctx.empty((cyc.S, 4), torch.int32)
        ctx.call('skb_path_heads', cyc.su, rec, cyc.ranked, cyc.aux, cyc.sel, cyc.S, 1, P0, 0, id_shift, start_node,
                 plen, pmin, einfo_b, head_end, head_aux)
    pindptr = ctx.empty(Pk + 1, torch.int32)   # local exclusive scan of the entries: regular paths, then cycles
    ctx.call('skb_scan_u32', plen, pindptr, Pk, ctx.scratch(Pk))
    ends = torch.stack([pindptr[P0], pindptr[Pk],
                        cross_cycle.int() if cross_cycle is not None else pindptr[0]]).tolist()
    L0, Lc, cc = ends[0], ends[1] - ends[0], ends[2]
    _mark('heads')
    counts2 = _allgather_ints([P0, L0, Pc, Lc, cc], ctx.device, group)            # collective 5
    if counts2[:, 4].any():
        raise NotImplementedError('a junction-free cycle crosses a slab boundary')
    _mark('allgather_counts2')
    P0off = np.concatenate([[0], np.cumsum(counts2[:, 0])])
    L0off = np.concatenate([[0], np.cumsum(counts2[:, 1])])
    Pcoff = np.concatenate([[0], np.cumsum(counts2[:, 2])])
    Lcoff = np.concatenate([[0], np.cumsum(counts2[:, 3])])
    P0_total, L0_total, Pc_total, Lc_total = int(P0off[-1]), int(L0off[-1]), int(Pcoff[-1]), int(Lcoff[-1])
    P_total, L_total = P0_total + Pc_total, L0_total + Lc_total
    if P_total == 0:
        raise ValueError('index pointer size 0 should be 1')
    # ---- path records to every rank -----------------------------------------------------------------
    maxP = int(counts2[:, 0].max())
    ex = _Exchange(ctx, world, maxP, [16, 16], [torch.int32, torch.int32])
    ctx.call('skb_slab_head_records', head_end, plen, pindptr, pmin, start_node, startoff, head_aux, slab.data_ptr(),
             Ne, P0, ex.cap, int(P0off[rank]), int(L0off[rank]), ex.field(0), ex.field(1))
    heads_all, links_all = ex.gather(group)                                       # collective 3
    maxP = ex.cap
    nrec = world * maxP
    einfo = ctx.empty((max(S_total, 1), 4), torch.int32)
    ctx.call('skb_slab_einfo', heads_all, nrec, S_total, einfo)
    # ---- path entries: every rank writes the entries of its own voxels at their global positions ----
    pidx = torch.zeros(L_total, dtype=torch.int32, device=ctx.device)
    pdata = torch.zeros(L_total, dtype=torch.float64, device=ctx.device)
    pw = torch.zeros(L_total, dtype=torch.float64, device=ctx.device)
    ctx.call('skb_path_emit', indices, gdata, rec, su, se, ranked, einfo, None, values, None, id_shift, S, pidx,
             pdata, pw)
    if Pc:
        gptr = pindptr.clone()
        gptr[P0:] += (L0_total + int(Lcoff[rank])) - L0
        ctx.call('skb_path_emit', indices, gdata, rec, cyc.su, cyc.se, cyc.ranked, einfo_b, gptr, values, None,
                 id_shift, cyc.S, pidx, pdata, pw)
    # ---- skeleton ids ------------------------------------------------------------------------------------
    kpar = ctx.empty(max(S_total, 1), torch.int32)
    kmin = ctx.empty(max(S_total, 1), torch.int32)
    node_lo, node_hi = int(Noff[rank]), int(Noff[rank + 1])
    ctx.call('skb_slab_components', links_all, nrec, S_total, kpar, kmin, node_lo, node_hi, id_shift, is_min, Ne,
             ctx.scratch(Ne))
    cr = is_min.index_select(0, sel_idx[:2]).tolist()
    _mark('emit_components')
    C_k = cr[1] - cr[0]
    counts3 = _allgather_ints([C_k], ctx.device, group)                             # collective 7
    Coff = np.concatenate([[0], np.cumsum(counts3[:, 0])])
    _mark('allgather_counts3')
    skel_all = ctx.empty(max(nrec, 1), torch.int32)
    ctx.call('skb_slab_skeleton_ids', links_all, nrec, kpar, kmin, is_min, node_lo, node_hi, id_shift,
             int(Coff[rank]), own0, skel_all)
    if world > 1:
        dist.all_reduce(skel_all, op=dist.ReduceOp.SUM, group=group)               # collective 8
    skel = ctx.empty(max(Pk, 1), torch.int32)
    if P0:
        skel[:P0] = skel_all[rank * maxP: rank * maxP + P0]
    if Pc:
        roots_local = (start_node[P0:Pk] - id_shift).long()
        skel[P0:Pk] = int(Coff[rank]) + is_min.index_select(0, roots_local) - cr[0]
    _mark('skel_allreduce(launch)')
    # ---- statistics and branch table of the paths owned here ------------------------------------------
    dist_, mean_, stdev_ = (ctx.empty(max(Pk, 1), torch.float64) for _ in range(3))
    ctx.call('skb_slab_path_stats', plen, head_aux, Pk, dist_, mean_, stdev_)
    ndim = 3
    sp = np.asarray(spacing)
    if sp.ndim == 0:
        sp = np.full(ndim, sp)
    sp_is_int = sp.dtype.kind in 'iub'
    sp_f = torch.from_numpy(np.ascontiguousarray(sp, dtype=np.float64))
    sp_i = torch.from_numpy(np.ascontiguousarray(sp.astype(np.int64) if sp_is_int else np.ones(ndim, np.int64)))
    gshape = torch.tensor([NZ, NY, NX], dtype=torch.int64)
    o32 = ctx.empty((3, max(Pk, 1)), torch.int32)
    o64 = ctx.empty((1 + 2 * ndim, max(Pk, 1)), torch.int64)
    of = ctx.empty((2 * ndim + 1, max(Pk, 1)), torch.float64)
    if Pk:
        ctx.call('skb_branch_table', Pk, ndim, 0, 1 if sp_is_int else 0, sp_f.data_ptr(), sp_i.data_ptr(), None, None,
                 indptr, skel, g.coords, values, head_aux, start_node, id_shift, gshape.data_ptr(), o32, o64, of)
    _mark('table_launch')
    return SimpleNamespace(
        rank=rank, world=world, graph=g, own=(own0, own1), id_shift=id_shift,
        n_nodes=N_k, n_edges=E_k, node_offset=int(Noff[rank]), edge_offset=int(Eoff[rank]),
        n_nodes_total=N_total, n_edges_total=int(Eoff[-1]),
        n_regular=P0, n_cycles=Pc, n_paths_total=P_total, n_entries_total=L_total,
        n_regular_total=P0_total, regular_offset=int(P0off[rank]), cycle_offset=P0_total + int(Pcoff[rank]),
        path_len=plen[:Pk], paths_indices=pidx, paths_data=pdata, paths_weights=pw,
        entry_offset_regular=int(L0off[rank]), entry_offset_cycles=L0_total + int(Lcoff[rank]),
        dist=dist_[:Pk], mean=mean_[:Pk], stdev=stdev_[:Pk], table=(o32[:, :Pk], o64[:, :Pk], of[:, :Pk]),
        spacing_is_int=sp_is_int, rank_rounds=rank_rounds, table_rounds=table_rounds, n_starts=S_k,
        n_table=int(nblk.sum()))


def owned_graph(res):
    """This rank's rows of the global graph: (indptr with global offsets, global indices, data, coordinates)."""
    g = res.graph
    a, b = res.own
    indptr = g.indptr[a:b + 1]
    e0, e1 = int(indptr[0].item()), int(indptr[-1].item())
    return (indptr - e0 + res.edge_offset, g.indices[e0:e1] + res.id_shift, g.data[e0:e1], g.coords[a:b],
            g.degrees[a:b])
