"""Sholl analysis (reference csr.py:1211-1390) on a skan_b200.Skeleton.

The node- and edge-proportional part -- distance of every node from the centre, np.digitize of those distances,
histogram of the edges whose two ends lie in different shells -- runs in the CUDA library (skb_sholl_bins /
skb_sholl_count, skan_b200/csrc/skb_sholl.cuh) on the graph and coordinates that are already resident on the
device.  What stays on the host is what the reference also does on a few numbers: normalising the shells, and
the geodesic centre, which is a closeness-centrality computation on the *simplified* graph (one node per
junction / endpoint, one edge per branch: the P-row branch table), SURVEY.md section 8(f2).

Third-party restatement: the reference takes the geodesic centre from ``skimage.graph.central_pixel``
(scikit-image >= 0.17.1, not installed here).  `central_pixel` below restates its published algorithm (all-pairs
shortest path lengths with SciPy, infinite lengths counted as 0, the node with the smallest non-zero total wins,
first one on ties); it is pinned by the reference's own known answers (test_csr.py:265-281: centre ids 6 and 1,
centre coordinates (3, 3)).
"""
from __future__ import annotations

import ctypes
import warnings

import numpy as np
import torch
from scipy import sparse
from scipy.sparse import csgraph


def central_pixel(graph, nodes=None, shape=None, partition_size=100):
    """Node with the smallest total shortest-path length to all others (closeness centrality), and the totals."""
    n = graph.shape[0]
    if nodes is None:
        nodes = np.arange(n)
    num_splits = 1 if partition_size is None else max(2, n // partition_size)
    totals = []
    for part in np.array_split(np.arange(n), num_splits):
        if part.size == 0:
            continue
        lengths = csgraph.shortest_path(graph, directed=False, indices=part)
        totals.append(np.sum(np.nan_to_num(lengths, posinf=0), axis=1))
    total = np.concatenate(totals) if totals else np.zeros(0)
    nonzero = np.flatnonzero(total != 0)
    raw = nonzero[np.argmin(total[nonzero])]
    central = nodes[raw]
    if shape is not None:
        central = np.unravel_index(central, shape)
    return central, total


def _simplify_graph(skel):
    """Graph with the degree-2 nodes removed: one edge per branch (csr.py:1211-1262).  Returns the symmetric
    CSR matrix of branch distances and the original node id of every simplified node."""
    from .csr import summarize
    if np.sum(skel.degrees > 2) == 0:  # no junctions: nothing to reduce
        return skel.graph, np.arange(skel.graph.shape[0])
    summary = summarize(skel, separator='_')
    src = np.asarray(summary['node_id_src'])
    dst = np.asarray(summary['node_id_dst'])
    distance = np.asarray(summary['branch_distance'])
    nodes = np.unique(np.append(src, dst))
    n = len(nodes)
    src_relab, dst_relab = np.searchsorted(nodes, src), np.searchsorted(nodes, dst)
    directed = sparse.coo_matrix((distance, (src_relab, dst_relab)), shape=(n, n)).tocsr()
    return directed + directed.T, nodes


def _fast_graph_center_idx(skel):
    """Index of the geodesic centre of the skeleton, found on the simplified graph (csr.py:1265-1283)."""
    simple, reduced_nodes = _simplify_graph(skel)
    idx, _ = central_pixel(simple)
    return reduced_nodes[idx]


def _max_distance(skel, center):
    """max_i || coordinates_i * spacing - center ||, on the device (csr.py:1314-1316)."""
    ctx = skel._ctx
    n = int(skel._coords_dev.shape[0])
    out = torch.zeros(1, dtype=torch.int64, device=ctx.device)
    _call_bins(skel, center, None, 0, 0, None, None, out)
    # np.float64 like the reference's np.max: a Python float would make `end_radius + float32 eps` a float32 (NEP 50)
    return out.cpu().numpy().view(np.float64)[0] if n else np.float64(0.0)


def _geom_args(skel, center):
    ndim = int(skel._coords_dev.shape[1])
    spacing = np.asarray(skel.spacing)
    is_int = spacing.dtype.kind in 'iub'
    sp_f = (ctypes.c_double * 3)(*([float(x) for x in spacing] + [1.0] * (3 - ndim)))
    sp_i = (ctypes.c_int64 * 3)(*([int(x) if is_int else 1 for x in spacing] + [1] * (3 - ndim)))
    c = (ctypes.c_double * 3)(*([float(x) for x in center] + [0.0] * (3 - ndim)))
    return ndim, int(is_int), sp_f, sp_i, c


def _call_bins(skel, center, radii_dev, n_radii, decreasing, bins, dist, max_bits):
    ndim, is_int, sp_f, sp_i, c = _geom_args(skel, center)
    n = int(skel._coords_dev.shape[0])
    skel._ctx.call('skb_sholl_bins', skel._coords_dev, n, ndim, is_int, sp_f, sp_i, c, radii_dev, n_radii,
                   decreasing, bins, dist, max_bits)


def _normalize_shells(shells, *, center, skel, spacing):
    """Shell radii from any of the forms sholl_analysis accepts (csr.py:1286-1329); the largest node distance
    it needs for an int / None `shells` is reduced on the device."""
    if isinstance(shells, (list, tuple, np.ndarray)):
        shell_radii = np.asarray(shells)
    else:
        start_radius = 0
        end_radius = _max_distance(skel, center)
        if shells is None:
            stepsize = np.linalg.norm(spacing)
        else:
            stepsize = (end_radius - start_radius) / shells
        epsilon = np.finfo(np.float32).eps
        shell_radii = np.arange(start_radius, end_radius + epsilon, stepsize)
    if (sp := np.linalg.norm(spacing)) > (sh := np.min(np.diff(shell_radii))):
        warnings.warn(
            'This implementation of Sholl analysis may not be accurate if '
            'the spacing between shells is smaller than the (diagonal) '
            f'voxel spacing. The given voxel spacing is {sp}, and the '
            f'smallest shell spacing is {sh}.',
            stacklevel=3,
        )
    return shell_radii


def sholl_analysis(skeleton, center=None, shells=None):
    """Sholl analysis of a Skeleton (csr.py:1332-1390): returns (center, shell_radii, intersection_counts)."""
    if center is None:
        center_idx = _fast_graph_center_idx(skeleton)
        center = skeleton.coordinates[center_idx] * skeleton.spacing
    else:
        center = np.asarray(center)
    shell_radii = _normalize_shells(shells, center=center, skel=skeleton, spacing=skeleton.spacing)
    radii = np.ascontiguousarray(shell_radii, dtype=np.float64)
    diffs = np.diff(radii)
    if np.all(diffs >= 0):
        decreasing = 0
    elif np.all(diffs <= 0):
        decreasing = 1
    else:
        raise ValueError('bins must be monotonically increasing or decreasing')  # np.digitize's error
    ctx = skeleton._ctx
    n = int(skeleton._coords_dev.shape[0])
    r = int(radii.size)
    radii_dev = torch.from_numpy(radii).to(ctx.device)
    bins = torch.empty(max(n, 1), dtype=torch.int32, device=ctx.device)
    _call_bins(skeleton, center, radii_dev, r, decreasing, bins, None, None)
    hist = torch.zeros(r + 1, dtype=torch.int64, device=ctx.device)
    indptr, indices, _ = skeleton._graph_dev
    ctx.call('skb_sholl_count', indptr, indices, bins, n, hist)
    # the graph is undirected: every crossing was counted from both ends (csr.py:1383-1385)
    counts = hist.cpu().numpy()
    top = int(np.max(np.flatnonzero(counts), initial=-1)) + 1
    intersection_counts = counts[:max(r, top)] // 2
    return center, shell_radii, intersection_counts
