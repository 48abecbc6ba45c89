"""Host-side sequencing of the sm_100a kernels: skeleton image -> pixel graph -> junction MST
-> CSR -> paths -> per-path statistics -> branch table.

Every array lives in a torch tensor on the image's device; the kernels are reached through
the C-ABI (skan_b200/_native.py).  The host only allocates (count -> allocate -> fill) and
reads back the handful of counts that size the next allocation (N, junction edges, E, P, L).

Reference stages replaced (relative to /root/reference/src/skan):
  csr.py:1070,122-123,131  -> pack_count + scan + emit_nodes     (mask sweep, raster node ids)
  csr.py:127-157           -> raw_neighbors (+ csr_emit)         (3^d stencil, CSR)
  csr.py:980-1020          -> junction_edge_* + mst_*            (junction MST)
  csr.py:347-446           -> components + rank_arcs + path_*    (path tracing)
  csr.py:962-977, 792-816  -> path_stats                          (distance, mean, stdev)
  csr.py:909-952           -> branch_table                        (summarize columns)
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from . import _native

# torch dtype -> element code of skb_items.cuh
_DT = {
    torch.bool: 0, torch.uint8: 0, torch.int8: 1, torch.int16: 3, torch.int32: 5, torch.int64: 7,
    torch.float32: 8, torch.float64: 9,
}
for _name, _code in (('uint16', 2), ('uint32', 4), ('uint64', 6)):
    if hasattr(torch, _name):
        _DT[getattr(torch, _name)] = _code

_NP_DT = {
    np.dtype(np.bool_): 0, np.dtype(np.uint8): 0, np.dtype(np.int8): 1, np.dtype(np.uint16): 2,
    np.dtype(np.int16): 3, np.dtype(np.uint32): 4, np.dtype(np.int32): 5, np.dtype(np.uint64): 6,
    np.dtype(np.int64): 7, np.dtype(np.float32): 8, np.dtype(np.float64): 9,
}
_ESIZE = {0: 1, 1: 1, 2: 2, 3: 2, 4: 4, 5: 4, 6: 8, 7: 8, 8: 4, 9: 8}

# name -> list of (start, stop) CUDA events, filled by _Ctx.call when not None; the key '*' times
# every entry point, otherwise only the names already present are timed
EVENT_HOOK = None

# 'fused' (native runner only): sweep + scan + node emission in one pass (skb_sweep_nodes); the per-stage
# sequencing below runs the TMA sweep for it.  Measured slower than 'tma' + scan + node emission on B200
# (2048^3: 4.6 ms against 1.5 + 0.45 ms), so it is an option, not the default.
PACK_VARIANT = {'ldg': 0, 'tma': 1, 'fused': 2}
default_pack_variant = 'tma'


def neighbour_weights(ndim, spacing):
    """fp64 weight of each of the 27 neighbour codes k27 = (dz+1)*9+(dy+1)*3+(dx+1) for an image
    embedded as (nz, ny, nx).  The arithmetic is nputil.py:278-279 verbatim in NumPy, evaluated
    on the ndim real axes with spacing = ones(ndim) * spacing (csr.py:1072), so the bits equal
    the reference's edge weights."""
    spacing = np.ones(ndim, dtype=float) * spacing
    grids = np.meshgrid(*([np.arange(-1, 2)] * 3), indexing='ij')
    offs3 = np.stack([g.ravel() for g in grids], axis=-1)          # (27, 3), k27 order
    offs = offs3[:, 3 - ndim:]
    weighted = offs * spacing
    dist = np.sqrt(np.sum(weighted**2, axis=1))
    dist[np.any(offs3[:, :3 - ndim] != 0, axis=1)] = 0.0            # codes outside the real axes: unused
    return np.ascontiguousarray(dist, dtype=np.float64)


class _Ctx:
    """Allocation + call helper bound to one device and stream."""

    def __init__(self, device):
        self.lib = _native.library()
        self.device = torch.device(device)
        if self.lib.device_type == 'cuda':
            if self.device.type != 'cuda':
                raise _native.NativeError('skan_b200 runs on CUDA devices only (no CPU fallback); got ' + str(device))
            self.stream = torch.cuda.current_stream(self.device).cuda_stream
        else:  # host emulation harness injected by tests/
            if self.device.type != 'cpu':
                raise _native.NativeError('emulation library expects CPU tensors')
            self.stream = 0
        self._scratch = None

    def guard(self):
        """Make this context's device the current CUDA device for the duration of a native call: the library
        launches on the current device, which need not be the image's (Skeleton(img_on_cuda1) under cuda:0)."""
        if self.device.type == 'cuda':
            return torch.cuda.device(self.device)
        import contextlib
        return contextlib.nullcontext()

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def scratch(self, n):
        need = self.lib.scan_scratch_bytes(n)
        if self._scratch is None or self._scratch.numel() < need:
            # zeroed once: the single-pass scan tags its tile statuses with an epoch
            self._scratch = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=self.device)
        return self._scratch.data_ptr()

    def call(self, name, *args):
        conv = [a.data_ptr() if isinstance(a, torch.Tensor) else (0 if a is None else a) for a in args]
        hook = EVENT_HOOK
        if hook is not None and self.device.type == 'cuda' and (hook.get('*') is not None or name in hook):
            # CUDA events on the launching stream around this entry point (bench.py / profiling only)
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            with self.guard():
                a.record(torch.cuda.current_stream(self.device))
                r = self.lib.call(name, *conv, self.stream)
                b.record(torch.cuda.current_stream(self.device))
            hook.setdefault(name, []).append((a, b))
            return r
        with self.guard():
            return self.lib.call(name, *conv, self.stream)


def _as_device_image(image, device):
    """-> (tensor on device viewed through a supported dtype, element code, shape, numpy dtype)."""
    if isinstance(image, torch.Tensor):
        t = image
        if t.dtype not in _DT:
            raise TypeError(f'unsupported skeleton image dtype {t.dtype}')
        code = _DT[t.dtype]
        np_dtype = np.dtype(str(t.dtype).replace('torch.', '')) if t.dtype != torch.bool else np.dtype(bool)
        shape = tuple(t.shape)
        t = t.to(device, non_blocking=True).contiguous()
    else:
        arr = np.asarray(image)
        if arr.dtype not in _NP_DT:
            if arr.dtype == np.float16:
                arr = arr.astype(np.float32)
            else:
                raise TypeError(f'unsupported skeleton image dtype {arr.dtype}')
        code = _NP_DT[arr.dtype]
        np_dtype = arr.dtype
        shape = arr.shape
        arr = np.ascontiguousarray(arr)
        t = torch.from_numpy(arr.reshape(-1).view(np.uint8)).to(device, non_blocking=True)
    if t.data_ptr() % 16:
        t = t.clone()
    return t, code, shape, np_dtype


def graph_nodes(image, *, spacing=1, value_is_height=False, device, pack_variant=None, z_origin=0,
                rank_positions=None, image_stack=False):
    """Voxel and node stages of skeleton_to_csgraph (csr.py:1070,122-144,1088): mask sweep, nodes in
    raster order (ids, raveled indices, coordinates, values) and raw 3^d neighbourhood masks."""
    ctx = _Ctx(device)
    img, code, shape, np_dtype = _as_device_image(image, ctx.device)
    ndim = len(shape)
    if ndim < 1 or ndim > 3:
        raise NotImplementedError('skan_b200 supports 1-D, 2-D and 3-D skeleton images')
    V = int(np.prod(shape, dtype=np.int64))
    is_bool = np_dtype == np.dtype(bool)
    if image_stack and ndim != 3:
        raise ValueError('a stack of 2-D images is a 3-D array (image, row, column)')
    # a stack of 2-D images: 2-D weights on the last two axes, no neighbours across the first
    wtab_np = neighbour_weights(2 if image_stack else ndim, spacing)
    shape_t = torch.tensor(list(shape), dtype=torch.int64)  # host: read by the C side only
    tile = ctx.lib.tile_voxels
    ntiles = max(1, -(-V // tile))
    variant = min(1, PACK_VARIANT[pack_variant or default_pack_variant])

    g = SimpleNamespace(ctx=ctx, shape=tuple(shape), ndim=ndim, np_dtype=np_dtype, dtype_code=code,
                        spacing=spacing, value_is_height=bool(value_is_height), nvox=V)
    # --- mask sweep: 1 bit / voxel + per-tile counts, then the tile prefix ------------------
    # zero pad words on either side; the mask starts one 128-byte line into the allocation, so every 256-bit
    # block of the rank structure is one 32-byte sector
    store = ctx.empty(ntiles * 256 + 18, torch.int64)
    store[:16] = 0
    store[-2:] = 0
    bits = store[16:-2]
    tile_prefix = ctx.empty(ntiles + 1, torch.int32)
    if V > 0:
        ctx.call('skb_pack_count', img, code, V, bits, tile_prefix, ntiles, variant)
    else:
        bits.zero_()
        tile_prefix.zero_()
    ctx.call('skb_scan_u32', tile_prefix, tile_prefix, ntiles, ctx.scratch(ntiles))
    g.position_ranks = None
    if rank_positions is not None and all(p % tile == 0 and p // tile <= ntiles for p in rank_positions):
        # ranks at tile-aligned voxel positions (Z-slab plane boundaries) come with the same read-back as N
        idx = torch.tensor([ntiles] + [p // tile for p in rank_positions], dtype=torch.int64, device=ctx.device)
        vals = tile_prefix.index_select(0, idx).tolist()
        N, g.position_ranks = vals[0], vals[1:]
    else:
        N = _count(tile_prefix[ntiles], 'nodes')
    g.n_nodes = N
    g.bits = bits
    g._bits_store = store
    # --- nodes in raster order ---------------------------------------------------------------
    lin = ctx.empty(N, torch.int64)
    coords = ctx.empty((N, ndim), torch.int64)
    values = None if is_bool else ctx.empty(N, torch.float64)
    pre256 = ctx.empty(ntiles * 64, torch.int32)
    ctx.call('skb_emit_nodes', bits, tile_prefix, ntiles, ndim, shape_t.data_ptr(), img, code, int(z_origin), lin,
             coords, values, pre256)
    g.lin, g.coords, g.values, g.pre256 = lin, coords, values, pre256
    if value_is_height and values is None:
        # the reference subtracts booleans here, which NumPy refuses (csr.py:1024)
        raise TypeError('value_is_height needs a numeric (non-boolean) skeleton image')
    if value_is_height and np_dtype.kind == 'u':
        # csr.py:1024 subtracts in the image dtype: unsigned heights wrap around and the reference
        # ends up with direction-dependent weights (SURVEY.md A.8 ii).  Rejected, not reproduced.
        raise ValueError('value_is_height with an unsigned integer image wraps the height differences '
                         '(reference csr.py:1024); cast the image to a signed or floating dtype')
    g.wtab = torch.from_numpy(wtab_np).to(ctx.device)
    g.geom = (ndim, shape_t.data_ptr())
    g.height = 1 if value_is_height else 0
    # --- raw neighbourhood masks ---------------------------------------------------------------
    nb = ctx.empty(N, torch.int32)
    ctx.call('skb_raw_neighbors', bits, *g.geom, lin, N, 1 if image_stack else 0, nb)
    g.raw_nb = nb
    g._keepalive = (img, shape_t)
    return g


def junction_edge_counts(g, a0=0, a1=None):
    """Per-node offsets (scanned, N+1) and the number of junction-junction edges (csr.py:1004-1016)
    whose smaller endpoint is a node in [a0, a1)."""
    ctx, N = g.ctx, g.n_nodes
    a1 = N if a1 is None else a1
    jcount = ctx.empty(N + 1, torch.int32)
    ctx.call('skb_junction_edge_count', g.bits, g.pre256, *g.geom, g.lin, g.raw_nb, N, a0, a1, jcount)
    ctx.call('skb_scan_u32', jcount, jcount, N, ctx.scratch(N))
    return jcount, _count(jcount[N], 'junction edges')


def junction_edge_emit(g, jcount, ea, eb, ek, ew):
    """Write the edges counted by junction_edge_counts, in (a, k27) order: ea, eb int32 node ids,
    ek uint8 neighbour code, ew int64 weight bits (tensors or raw device addresses)."""
    g.ctx.call('skb_junction_edge_emit', g.bits, g.pre256, *g.geom, g.lin, g.raw_nb, jcount, g.n_nodes, g.wtab,
               g.values, g.height, g.dtype_code, ea, eb, ek, ew)


def junction_edges(g, a0=0, a1=None):
    ctx = g.ctx
    jcount, nej = junction_edge_counts(g, a0, a1)
    ea = ctx.empty(nej, torch.int32)
    eb = ctx.empty(nej, torch.int32)
    ek = ctx.empty(nej, torch.uint8)
    ew = ctx.empty(nej, torch.int64)
    if nej:
        junction_edge_emit(g, jcount, ea, eb, ek, ew)
    return ea, eb, ek, ew


def junction_mst(ctx, ea, eb, ew, n_ids):
    """Junction MST (csr.py:980-1020) over edges whose endpoints are ids < n_ids: tree flags uint8[nej]."""
    nej = int(ea.numel())
    tree = ctx.empty(nej, torch.uint8)
    if not nej:
        return tree, 0
    par = ctx.empty(n_ids, torch.int32)
    bestw = ctx.empty(n_ids, torch.int64)
    beste = ctx.empty(n_ids, torch.int32)
    era = ctx.empty(nej, torch.int32)
    erb = ctx.empty(nej, torch.int32)
    flag = ctx.empty(4, torch.int32)
    rounds = ctx.call('skb_mst_junctions', ea, eb, ew, nej, n_ids, par, bestw, beste, era, erb, tree, flag)
    return tree, rounds


def graph_csr(g, ea, eb, ek, tree, id_shift=0, records=True, slab=None, cc_init=True):
    """Final CSR (csr.py:153-157 after 1018-1019): drop the non-tree junction edges from the raw
    masks, degrees, indptr, indices, data.  With records=True the emission pass also writes what the
    path stage walks on (g.path_state: node records, start offsets, union-find init)."""
    ctx, N = g.ctx, g.n_nodes
    nej = int(ea.numel())
    if nej:
        keep = g.raw_nb.clone()
        ctx.call('skb_mst_apply', ea, eb, ek, tree, nej, id_shift, N, keep)
    else:
        keep = g.raw_nb
    deg = ctx.empty(N, torch.int32)
    indptr = ctx.empty(N + 1, torch.int32)
    ctx.call('skb_degrees_indptr', keep, N, deg, indptr, ctx.scratch(N))
    E = _count(indptr[N], 'half-edges')
    indices = ctx.empty(E, torch.int32)
    data = ctx.empty(E, torch.float64)
    rec = startoff = par = cmin = is_min = None
    if records:
        rec = ctx.empty((max(N, 1), 8), torch.int32)
        startoff = ctx.empty(N + 1, torch.int32)
        is_min = ctx.empty(N + 1, torch.int32)
        if cc_init:
            par = ctx.empty(max(N, 1), torch.int32)
            cmin = ctx.empty(max(N, 1), torch.int32)
    ctx.call('skb_csr_emit', g.bits, g.pre256, *g.geom, g.lin, keep, indptr, N, g.wtab, g.values, g.height,
             g.dtype_code, indices, data, rec, slab, 1 if slab is not None else 0, startoff, par, cmin, is_min,
             ctx.scratch(N))
    g.path_state = (rec, startoff, par, cmin, is_min) if records else None
    g.keep, g.degrees, g.indptr, g.indices, g.data, g.n_edges = keep, deg, indptr, indices, data, E
    return g


def pixel_graph(image, *, spacing=1, value_is_height=False, device, pack_variant=None):
    """skeleton_to_csgraph on the device (csr.py:1028-1089): returns a namespace of device tensors
    (coords int64 (N,d), values fp64 (N,) | None, indptr/indices int32, data fp64, degrees int32)."""
    g = graph_nodes(image, spacing=spacing, value_is_height=value_is_height, device=device,
                    pack_variant=pack_variant)
    ea, eb, ek, ew = junction_edges(g)
    g.n_junction_edges = int(ea.numel())
    tree, g.mst_rounds = junction_mst(g.ctx, ea, eb, ew, g.n_nodes)
    return graph_csr(g, ea, eb, ek, tree)


def trace_paths(ctx, indptr, indices, gdata, values, n_nodes, n_edges, lin=None, state=None):
    """_build_skeleton_path_graph on the device (csr.py:412-446) by sampled list ranking, plus the
    skeleton ids summarize needs (csr.py:909).  Raises ValueError when there is no path, like the
    reference's csr_matrix constructor does at csr.py:442."""
    N, E = n_nodes, n_edges
    p = SimpleNamespace(n_nodes=N)
    if N == 0 or E == 0:
        raise ValueError('index pointer size 0 should be 1')
    if state is not None:
        rec, startoff0, cc_par, cc_min, is_min = state
    else:
        rec = ctx.empty((N, 8), torch.int32)
        startoff0 = ctx.empty(N + 1, torch.int32)
        cc_par = ctx.empty(N, torch.int32)
        cc_min = ctx.empty(N, torch.int32)
        is_min = ctx.empty(N + 1, torch.int32)
        ctx.call('skb_path_records', indptr, indices, gdata, lin, None, N, rec, startoff0, cc_par, cc_min, is_min,
                 ctx.scratch(N))
    p.rec = rec

    def phase(ph, covered, pid_base):
        if ph == 0:
            startoff = startoff0
        else:
            startoff = ctx.empty(N + 1, torch.int32)
            ctx.call('skb_path_start_count', rec, covered, None, N, startoff, ctx.scratch(N))
        S = _count(startoff[N], 'walk starts')
        if S == 0:
            return None
        su = ctx.empty(S, torch.int32)
        se = ctx.empty(S, torch.int32)
        buf = [ctx.empty((S, 4), torch.int32), ctx.empty((S, 4), torch.int32)]
        flag = ctx.empty(4, torch.int32)
        ctx.call('skb_path_walk', indptr, indices, gdata, values, lin, rec, startoff, None, N, S, su, se, buf[0],
                 None, None)
        r = ctx.call('skb_path_rank', buf[0], buf[1], None, None, S, 0, S, flag)
        ranked = buf[r & 1]
        sel = ctx.empty(S + 1, torch.int32)
        ctx.call('skb_path_select', su, rec, ranked, S, ph, 0, sel, ctx.scratch(S))
        return SimpleNamespace(S=S, su=su, se=se, ranked=ranked, sel=sel, rounds=r >> 1, phase=ph, pid_base=pid_base)

    def heads(ph, start_node, plen, pmin):
        ph.einfo = ctx.empty((ph.S, 4), torch.int32)
        ctx.call('skb_path_heads', ph.su, rec, ph.ranked, None, ph.sel, ph.S, ph.phase, ph.pid_base, 0, 0, start_node,
                 plen, pmin, ph.einfo, None, None)

    def emit(ph, pindptr, covered, pidx, pdata, pw):
        ctx.call('skb_path_emit', indices, gdata, rec, ph.su, ph.se, ph.ranked, ph.einfo, pindptr, values, covered,
                 0, ph.S, pidx, pdata, pw)

    # ---- phase 0: paths between terminals -----------------------------------------------------
    a = phase(0, None, 0)
    n_regular = int(a.sel[a.S].item()) if a is not None else 0
    p.rank_rounds = a.rounds if a is not None else 0
    # every edge lies on exactly one path, so the regular paths account for L - P of the E/2 edges; what
    # is left are the edges (= nodes) of junction-free cycles (csr.py:369-386).  Their number is only
    # known after the lengths of the regular paths; until then size by the worst case.
    P_cap = n_regular + E // 6 + 1
    start_node = ctx.empty(P_cap, torch.int32)
    plen = ctx.empty(P_cap + 1, torch.int32)
    pmin = ctx.empty(P_cap, torch.int32)
    pindptr = ctx.empty(P_cap + 1, torch.int32)
    L_reg = 0
    if n_regular:
        heads(a, start_node, plen, pmin)
        ctx.call('skb_scan_u32', plen, pindptr, n_regular, ctx.scratch(n_regular))
        L_reg = _count(pindptr[n_regular], 'path entries')
    n_uncovered = max(0, E // 2 - (L_reg - n_regular))
    L_cap = L_reg + n_uncovered + n_uncovered // 3 + 1
    pidx = ctx.empty(L_cap, torch.int32)
    pdata = ctx.empty(L_cap, torch.float64)
    pw = ctx.empty(L_cap, torch.float64)
    covered = None
    if n_uncovered > 0:
        covered = torch.zeros(N, dtype=torch.uint8, device=ctx.device)
    if n_regular:
        emit(a, pindptr, covered, pidx, pdata, pw)
    # ---- phase 1: junction-free cycles ----------------------------------------------------------
    n_cycles = 0
    if n_uncovered > 0:
        par = ctx.empty(N, torch.int32)
        ctx.call('skb_cycle_roots', rec, covered, N, par, is_min)
        b = phase(1, covered, n_regular)
        n_cycles = int(b.sel[b.S].item())
        heads(b, start_node, plen, pmin)
        P = n_regular + n_cycles
        ctx.call('skb_scan_u32', plen, pindptr, P, ctx.scratch(P))
        emit(b, pindptr, None, pidx, pdata, pw)
    P = n_regular + n_cycles
    if P == 0:
        raise ValueError('index pointer size 0 should be 1')
    L = L_reg + n_uncovered + n_cycles
    p.n_paths, p.n_cycles, p.n_regular, p.n_entries = P, n_cycles, n_regular, L
    p.indptr, p.indices, p.data, p.weights = pindptr[:P + 1], pidx[:L], pdata[:L], pw[:L]
    # ---- skeleton ids ---------------------------------------------------------------------------
    skel_id = ctx.empty(P, torch.int32)
    ctx.call('skb_skeleton_ids', p.indptr, p.indices, pmin, N, n_regular, P, cc_par, cc_min, is_min, skel_id,
             ctx.scratch(N))
    p.skeleton_id = skel_id
    p.component_rank = is_min   # exclusive scan of "node is the minimum of its component" (N + 1 entries)
    return p


def path_stats(ctx, p, *, dist=True, mean=True, stdev=True):
    P = p.n_paths
    out = [ctx.empty(P, torch.float64) if want else None for want in (dist, mean, stdev)]
    ctx.call('skb_path_stats', p.indptr, p.data, p.weights, P, *out)
    return out


def branch_table(ctx, p, gindptr, coords, values, spacing, ndim, value_is_height):
    """Device columns of summarize (csr.py:909-952): three SoA buffers (int32, int64, fp64)."""
    P = p.n_paths
    height = 1 if value_is_height else 0
    sp = np.asarray(spacing)
    if sp.ndim == 0:
        sp = np.full(ndim, sp)
    sp_is_int = sp.dtype.kind in 'iub'
    sp_f = torch.from_numpy(np.ascontiguousarray(sp, dtype=np.float64))
    sp_i = torch.from_numpy(np.ascontiguousarray(sp.astype(np.int64) if sp_is_int else np.ones(ndim, np.int64)))
    dh = ndim + height
    o32 = ctx.empty((3, P), torch.int32)
    o64 = ctx.empty((1 + 2 * ndim, P), torch.int64)
    of = ctx.empty((2 * dh + 1, P), torch.float64)
    ctx.call('skb_branch_table', P, ndim, height, 1 if sp_is_int else 0, sp_f.data_ptr(), sp_i.data_ptr(),
             p.indptr, p.indices, gindptr, p.skeleton_id, coords, values, None, None, 0, None, o32, o64, of)
    return o32, o64, of, sp_is_int


def skeleton_and_summary_device(image, *, spacing=1, value_is_height=False, device, pack_variant=None):
    """The whole hot path with every result left on the device: pixel graph + junction MST + CSR,
    paths, per-path statistics and the branch table (what Skeleton(...) + summarize(...) compute)."""
    g = pixel_graph(image, spacing=spacing, value_is_height=value_is_height, device=device,
                    pack_variant=pack_variant)
    p = trace_paths(g.ctx, g.indptr, g.indices, g.data, g.values, g.n_nodes, g.n_edges, state=g.path_state)
    dist, mean, stdev = path_stats(g.ctx, p)
    o32, o64, of, _ = branch_table(g.ctx, p, g.indptr, g.coords, g.values, spacing, g.ndim, value_is_height)
    return SimpleNamespace(graph=g, paths=p, dist=dist, mean=mean, stdev=stdev, table=(o32, o64, of))


def algorithmic_bytes(nvox, esize, ndim, n, e, p, l, valued):
    """every input read once, every mandatory output written once."""
    return (nvox * esize + n * 8 * ndim + (n + 1) * 4 + n * 4 + e * 12 + (n * 8 if valued else 0)
            + (p + 1) * 4 + l * 12 + p * (52 + 32 * ndim))


# ---- the same path through the native runner (skb_run_image): one C call per image ---------------
DEBUG_FILL = None   # tests: fill both arenas with this byte before every native call (see tests/test_gpu_guards.py)
LAST_ARENAS = None  # tests: (work, out) tensors of the last native call when DEBUG_FILL is set
TIME_SWEEP = False  # bench.py: record CUDA events around the mask sweep kernel (counts['sweep_ns'])
USE_HINTS = True    # pass the previous image's counts (+ slack) as capacity hints to skb_run_image


HINT_SLACK = 4096   # entries added to every hint on top of 1/16 of the previous count


def _count(t, what):
    """A count read back from the last element of a scanned int32 array.  The reference's index arrays are int32
    (csr.py:206-216): a graph whose node, half-edge or entry count does not fit comes back negative here, and is refused
    instead of being used as an allocation size."""
    v = int(t.item())
    if v < 0:
        raise OverflowError(f'the number of {what} exceeds int32: the reference\'s int32 index arrays cannot hold this graph')
    return v


def _hint(count):
    return int(count * 1.0625) + HINT_SLACK
_arena_cache = {}   # (device, stream) -> dict(work=tensor, out_bytes=int)
_wtab_cache = {}


_ESIZE = {torch.int32: 4, torch.int64: 8, torch.float64: 8, torch.uint8: 1, torch.uint32: 4, torch.float32: 4}


class _ArenaViews:
    """Typed views into the output arena, created on first access."""

    def __init__(self, arena, spec):
        # spec: {name: (byte offset, count, dtype, shape or None)}, or a callable that builds it on first access (the
        # interpreter time between two images is GPU idle time: a caller that reads two columns pays for two views)
        self._arena, self._spec = arena, spec

    def __getattr__(self, name):
        spec = self.__dict__['_spec']
        if callable(spec):
            spec = self.__dict__['_spec'] = spec()
        if name not in spec:
            raise AttributeError(name)
        off, count, dtype, shape = spec[name]
        if off < 0:
            t = None
        else:
            nbytes = count * _ESIZE[dtype]
            t = self._arena[off: off + nbytes].view(dtype)
            if shape is not None:
                t = t.view(shape)
        self.__dict__[name] = t
        return t


def run_native(image, *, spacing=1, value_is_height=False, device, pack_variant=None, image_stack=False,
               graph_only=False, z_origin=0):
    """Skeleton + summarize (or only skeleton_to_csgraph) of one image through ONE native call.
    Returns the same namespace as skeleton_and_summary_device, as lazily created views of an arena."""
    ctx = _Ctx(device)
    img, code, shape, np_dtype = _as_device_image(image, ctx.device)
    ndim = len(shape)
    if ndim < 1 or ndim > 3:
        raise NotImplementedError('skan_b200 supports 1-D, 2-D and 3-D skeleton images')
    is_bool = np_dtype == np.dtype(bool)
    if value_is_height and is_bool:
        raise TypeError('value_is_height needs a numeric (non-boolean) skeleton image')
    if value_is_height and np_dtype.kind == 'u':
        raise ValueError('value_is_height with an unsigned integer image wraps the height differences '
                         '(reference csr.py:1024); cast the image to a signed or floating dtype')
    wdim = 2 if image_stack else ndim
    sp_arr = np.asarray(spacing)
    wkey = (str(ctx.device), wdim, sp_arr.dtype.str, tuple(np.atleast_1d(sp_arr).tolist()))
    if wkey not in _wtab_cache:
        _wtab_cache[wkey] = torch.from_numpy(neighbour_weights(wdim, spacing)).to(ctx.device)
    wtab = _wtab_cache[wkey]
    tdim = 3 if image_stack else ndim
    sp = sp_arr if sp_arr.ndim else np.full(wdim, sp_arr)
    sp_is_int = sp.dtype.kind in 'iub'
    if image_stack:
        sp = np.concatenate([[1 if sp_is_int else 1.0], sp])
    prm = _native.RunParams()
    prm.image, prm.dtype, prm.ndim = img.data_ptr(), code, ndim
    for i, x in enumerate(shape):
        prm.shape[i] = int(x)
    prm.z_origin, prm.image_stack = int(z_origin), int(bool(image_stack))
    prm.height = int(bool(value_is_height))
    prm.pack_variant = PACK_VARIANT[pack_variant or default_pack_variant]
    prm.stop_after_graph, prm.want_values = int(bool(graph_only)), int(not is_bool)
    prm.time_sweep = int(TIME_SWEEP)   # 0, 1 (sweep kernel) or 2 (every stage boundary)
    prm.wtab = wtab.data_ptr()
    for i in range(tdim):
        prm.spacing_f[i] = float(sp[i])
        prm.spacing_i[i] = int(sp[i]) if sp_is_int else 1
    prm.spacing_is_int, prm.table_ndim = int(sp_is_int), tdim
    prm.stream = ctx.stream
    key = (str(ctx.device), ctx.stream)
    cache = _arena_cache.setdefault(key, dict(work=None, out_bytes=1 << 22, shape=None, nodes=0, counts=None))
    # Capacity hints: what the last image of this shape and kind needed, plus slack.  They only let the runner
    # launch the kernel that fills an array while the exact count is still on its way to the host; a hint that is
    # too small costs a second launch of that kernel, never a wrong result.
    kind = (tuple(shape), code, bool(image_stack), bool(graph_only))
    last = cache['counts'] if (USE_HINTS and cache['shape'] == kind) else None
    prm.node_cap_hint = _hint(last['nodes']) if last else 0
    for i, name in enumerate(_native.RUN_HINTS):
        prm.hint[i] = _hint(last[name]) if last else 0
        if name == 'junction_nodes' and last and last[name] == 0:
            prm.hint[i] = 0   # not counted yet (the previous image went without hints): arrays indexed by node id
    V = int(np.prod(shape, dtype=np.int64))
    if cache['work'] is None:
        cache['work'] = ctx.empty(max(1 << 22, V // 4 + (1 << 22)), torch.uint8)
    res = _native.RunResult()
    for _ in range(8):
        work = cache['work']
        out = ctx.empty(cache['out_bytes'], torch.uint8)
        prm.work, prm.work_bytes = work.data_ptr(), work.numel()
        prm.out, prm.out_bytes = out.data_ptr(), out.numel()
        if DEBUG_FILL is not None:
            global LAST_ARENAS
            work.fill_(DEBUG_FILL)
            out.fill_(DEBUG_FILL)
            LAST_ARENAS = (work, out)
        with ctx.guard():
            rc = ctx.lib.call('skb_run_image', ctypes_byref(prm), ctypes_byref(res))
        if rc == 0:
            break
        if res.work_need > work.numel():
            cache['work'] = ctx.empty(int(res.work_need), torch.uint8)
        if res.out_need > out.numel():
            cache['out_bytes'] = int(res.out_need)
    else:
        raise _native.NativeError('skb_run_image kept asking for larger arenas')
    if res.out_need > 0:   # what this image needed, plus slack, sizes the next image's output arena
        cache['out_bytes'] = max(1 << 20, int(res.out_need * 1.05) + (1 << 16))
    C = {k: int(res.counts[i]) for k, i in _native.RUN_COUNT.items()}
    cache['shape'], cache['nodes'], cache['counts'] = kind, C['nodes'], C
    if prm.time_sweep == 2:
        C['stage_ns'] = {k: int(res.stage_ns[i]) for i, k in enumerate(_native.RUN_STAGES)}
    N, E, P, L = C['nodes'], C['edges'], C['paths'], C['entries']
    h = int(bool(value_is_height))
    dh = tdim + h
    off = res.off

    def spec():
        O = {k: int(off[i]) for k, i in _native.RUN_OFF.items()}
        i32, i64, f64 = torch.int32, torch.int64, torch.float64
        return {
            'lin': (O['lin'], N, i64, None), 'coords': (O['coords'], N * ndim, i64, (N, ndim)),
            'values': (O['values'], N, f64, None), 'degrees': (O['degrees'], N, i32, None),
            'indptr': (O['indptr'], N + 1, i32, None), 'indices': (O['indices'], E, i32, None),
            'data': (O['data'], E, f64, None), 'comp_rank': (O['comp_rank'], N + 1, i32, None),
            'pindptr': (O['pindptr'], P + 1, i32, None), 'pidx': (O['pidx'], L, i32, None),
            'pdata': (O['pdata'], L, f64, None), 'pw': (O['pw'], L, f64, None), 'skel': (O['skel'], P, i32, None),
            'dist': (O['dist'], P, f64, None), 'mean': (O['mean'], P, f64, None), 'stdev': (O['stdev'], P, f64, None),
            't32': (O['t32'], 3 * P, i32, (3, P)), 't64': (O['t64'], (1 + 2 * tdim) * P, i64, (1 + 2 * tdim, P)),
            'tf': (O['tf'], (2 * dh + 1) * P, f64, (2 * dh + 1, P)),
        }

    v = _ArenaViews(out, spec)
    g = _GraphView(v, ctx, tuple(shape), ndim, np_dtype, code, spacing, bool(value_is_height), V, N, E,
                   C['junction_edges'], img)
    if graph_only:
        return SimpleNamespace(graph=g, paths=None, counts=C)
    if P == 0:
        raise ValueError('index pointer size 0 should be 1')
    p = _PathView(v, N, P, C['regular'], C['cycles'], L)
    return _RunView(v, g, p, sp_is_int, C)


def ctypes_byref(x):
    import ctypes
    return ctypes.byref(x)


class _GraphView:
    """Graph part of a native run; attribute names of the namespace graph_csr() returns."""

    def __init__(self, v, ctx, shape, ndim, np_dtype, code, spacing, height, nvox, n, e, nej, keepalive):
        self._v, self.ctx, self.shape, self.ndim, self.np_dtype = v, ctx, shape, ndim, np_dtype
        self.dtype_code, self.spacing, self.value_is_height, self.nvox = code, spacing, height, nvox
        self.n_nodes, self.n_edges, self.n_junction_edges, self.mst_rounds = n, e, nej, 0
        self._keepalive = keepalive
        self.path_state = None

    lin = property(lambda self: self._v.lin)
    coords = property(lambda self: self._v.coords)
    values = property(lambda self: self._v.values)
    degrees = property(lambda self: self._v.degrees)
    indptr = property(lambda self: self._v.indptr)
    indices = property(lambda self: self._v.indices)
    data = property(lambda self: self._v.data)


class _PathView:
    def __init__(self, v, n, p, n_regular, n_cycles, l):
        self._v, self.n_nodes, self.n_paths, self.n_regular, self.n_cycles, self.n_entries = v, n, p, n_regular, n_cycles, l
        self.rank_rounds = 0

    indptr = property(lambda self: self._v.pindptr)
    indices = property(lambda self: self._v.pidx)
    data = property(lambda self: self._v.pdata)
    weights = property(lambda self: self._v.pw)
    skeleton_id = property(lambda self: self._v.skel)
    component_rank = property(lambda self: self._v.comp_rank)


class _RunView:
    def __init__(self, v, g, p, spacing_is_int, counts):
        self._v, self.graph, self.paths, self.spacing_is_int, self.counts = v, g, p, spacing_is_int, counts

    dist = property(lambda self: self._v.dist)
    mean = property(lambda self: self._v.mean)
    stdev = property(lambda self: self._v.stdev)
    table = property(lambda self: (self._v.t32, self._v.t64, self._v.tf))
