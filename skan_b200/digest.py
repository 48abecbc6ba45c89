"""Partition-independent digest of a Skeleton + summarize result.

Every integer array of the result (graph.indptr / indices, the bit patterns of graph.data, coordinates, degrees,
paths.indptr / indices, the integer columns of the branch table) is reduced to ``sum_i w(i) * value_i mod 2**64`` with
pseudo-random odd weights ``w`` of the GLOBAL position ``i``.  The sum is linear, so every rank of a Z-slab
decomposition hashes the share it holds at its global offsets and the per-rank sums add up to the digest of the
whole result: a run on 1, 2, 4 or 8 GPUs (and the CPU oracle's NumPy arrays) give the same digest exactly when the
arrays are identical.  fp64 sums that are only defined to 1e-12 (branch_distance, mean, stdev, euclidean_distance) are
reported beside the digest as rounded totals, not hashed.

Used by bench.py (`parity.digest`), tests/ and anyone who wants to compare two runs without moving the arrays.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _i64(x):
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_C1, _C2, _C3 = _i64(0x9E3779B97F4A7C15), _i64(0xBF58476D1CE4E5B9), _i64(0x94D049BB133111EB)
_CHUNK = 1 << 24


def _weights(start, n, salt, device):
    i = torch.arange(start, start + n, dtype=torch.int64, device=device)
    x = i * _C1 + _i64(salt * 0xD6E8FEB86659FD93 + 0x2545F4914F6CDD1D)
    x = (x ^ (x >> 31)) * _C2
    x = (x ^ (x >> 29)) * _C3
    return (x ^ (x >> 32)) | 1


def _as_i64(t):
    if not isinstance(t, torch.Tensor):
        t = torch.from_numpy(np.array(t))   # a copy: SciPy / pandas hand out read-only views
    if t.dtype == torch.float64:
        return t.contiguous().view(torch.int64)
    if t.dtype == torch.bool:
        return t.to(torch.int64)
    return t.to(torch.int64)


def linear_hash(t, start=0, salt=0):
    """sum_i w(start + i) * t[i] mod 2**64 of a 1-D integer (or fp64: bit pattern) array, as a python int."""
    t = _as_i64(t).reshape(-1)
    acc = 0
    for o in range(0, int(t.numel()), _CHUNK):
        c = t[o:o + _CHUNK]
        w = _weights(start + o, int(c.numel()), salt, c.device)
        acc = (acc + int((w * c).sum().item())) & _M64
    return acc


_SALT = {n: i + 1 for i, n in enumerate(
    ['indptr', 'indices', 'data', 'coords', 'degrees', 'plen', 'pidx', 'pdata', 't32', 't64'])}


def _segments_single(r):
    """(name, tensor, global start) of a single-GPU run (the namespace _pipeline.run_native returns)."""
    g, p = r.graph, r.paths
    P = p.n_paths
    plen = p.indptr[1:P + 1] - p.indptr[:P]
    segs = [('indptr', g.indptr, 0), ('indices', g.indices, 0), ('data', g.data, 0), ('coords', g.coords.reshape(-1), 0),
            ('degrees', g.degrees, 0), ('plen', plen, 0), ('pidx', p.indices, 0)]
    if g.values is not None:
        segs.append(('pdata', p.data, 0))
    t32, t64, tf = r.table
    for row in range(t32.shape[0]):
        segs.append(('t32', t32[row], row * P))
    for row in range(t64.shape[0]):
        segs.append(('t64', t64[row], row * P))
    fp = dict(branch_distance=r.dist, mean_pixel_value=r.mean, stdev_pixel_value=r.stdev, euclidean_distance=tf[-1])
    return segs, fp, dict(nodes=g.n_nodes, edges=g.n_edges, paths=P, entries=p.n_entries)


def _segments_slab(res):
    """The same of one rank's share of a Z-slab run (skan_b200.slab.skeleton_and_summary_slab)."""
    from . import slab
    indptr, indices, data, coords, degrees = slab.owned_graph(res)
    n0, e0 = res.node_offset, res.edge_offset
    last = res.rank == res.world - 1
    segs = [('indptr', indptr if last else indptr[:-1], n0), ('indices', indices, e0), ('data', data, e0),
            ('coords', coords.reshape(-1), n0 * coords.shape[1]), ('degrees', degrees, n0)]
    P, nr = res.n_paths_total, res.n_regular
    segs += [('plen', res.path_len[:nr], res.regular_offset), ('plen', res.path_len[nr:], res.cycle_offset),
             ('pidx', res.paths_indices, 0)]   # entries of other ranks' voxels are zero here: the hash is linear
    if res.graph.values is not None:
        segs.append(('pdata', res.paths_data, 0))
    t32, t64, tf = res.table
    for name, t in (('t32', t32), ('t64', t64)):
        for row in range(t.shape[0]):
            segs.append((name, t[row, :nr], row * P + res.regular_offset))
            segs.append((name, t[row, nr:], row * P + res.cycle_offset))
    fp = dict(branch_distance=res.dist, mean_pixel_value=res.mean, stdev_pixel_value=res.stdev,
              euclidean_distance=tf[-1])
    return segs, fp, dict(nodes=res.n_nodes_total, edges=res.n_edges_total, paths=P, entries=res.n_entries_total)


def _segments_oracle(sk, summary):
    """The same of the CPU oracle's (or the reference's) Skeleton + summary columns (NumPy, host)."""
    g = sk.graph
    d = sk.coordinates.shape[1]
    P = sk.n_paths
    segs = [('indptr', g.indptr, 0), ('indices', g.indices, 0), ('data', np.asarray(g.data, np.float64), 0),
            ('coords', np.ascontiguousarray(sk.coordinates, dtype=np.int64).reshape(-1), 0),
            ('degrees', np.asarray(sk.degrees), 0), ('plen', np.diff(sk.paths.indptr), 0),
            ('pidx', sk.paths.indices, 0)]
    if sk.pixel_values is not None:
        segs.append(('pdata', np.asarray(sk.paths.data, np.float64), 0))
    col = {k.replace('-', '_'): np.asarray(v) for k, v in dict(summary).items()}
    t32 = [col['skeleton_id'], col['node_id_src'], col['node_id_dst']]
    t64 = [col['branch_type']] + [col[f'image_coord_src_{i}'] for i in range(d)] + \
          [col[f'image_coord_dst_{i}'] for i in range(d)]
    for row, t in enumerate(t32):
        segs.append(('t32', t, row * P))
    for row, t in enumerate(t64):
        segs.append(('t64', t, row * P))
    fp = {k: torch.from_numpy(np.array(col[k], dtype=np.float64)) for k in
          ('branch_distance', 'mean_pixel_value', 'stdev_pixel_value', 'euclidean_distance')}
    return segs, fp, dict(nodes=g.shape[0], edges=int(g.nnz), paths=P, entries=int(sk.paths.indices.size))


def _finish(comp, fp, counts):
    h = hashlib.sha256(repr(sorted(comp.items())).encode() + repr(sorted(counts.items())).encode()).hexdigest()[:16]
    return dict(digest=h, components={k: f'{v:016x}' for k, v in sorted(comp.items())},
                fp_sums={k: float(f'{v:.10e}') for k, v in sorted(fp.items())}, counts=counts)


def _reduce(segs, fp):
    comp = {}
    for name, t, start in segs:
        if t is None or int(np.prod(t.shape)) == 0:
            comp.setdefault(name, 0)
            continue
        comp[name] = (comp.get(name, 0) + linear_hash(t, int(start), _SALT[name])) & _M64
    sums = {k: (float(torch.as_tensor(v).to(torch.float64).sum().item()) if int(np.prod(v.shape)) else 0.0)
            for k, v in fp.items()}
    return comp, sums


def digest_single(r):
    segs, fp, counts = _segments_single(r)
    comp, sums = _reduce(segs, fp)
    return _finish(comp, sums, counts)


def digest_oracle(sk, summary):
    segs, fp, counts = _segments_oracle(sk, summary)
    comp, sums = _reduce(segs, fp)
    return _finish(comp, sums, counts)


def digest_slab(res, group=None):
    """Collective over the group: every rank hashes its share, the partial sums are added (mod 2**64) everywhere."""
    import torch.distributed as dist
    segs, fp, counts = _segments_slab(res)
    comp, sums = _reduce(segs, fp)
    world = dist.get_world_size(group)
    allp = [None] * world
    dist.all_gather_object(allp, (comp, sums), group=group)
    tot, fsum = {}, {}
    for c, s in allp:
        for k, v in c.items():
            tot[k] = (tot.get(k, 0) + v) & _M64
        for k, v in s.items():
            fsum[k] = fsum.get(k, 0.0) + v
    return _finish(tot, fsum, counts)


def same(a, b, rtol=1e-9):
    """Two digests describe the same result: identical integer components and counts, fp totals within rtol (the
    per-path values are defined to 1e-12; totals over 1e5 paths of either sign are compared a little looser)."""
    if a['components'] != b['components'] or a['counts'] != b['counts']:
        return False
    for k, v in a['fp_sums'].items():
        w = b['fp_sums'][k]
        if abs(v - w) > rtol * max(abs(v), abs(w), 1e-300):
            return False
    return True
