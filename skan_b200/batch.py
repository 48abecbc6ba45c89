"""Batches of independent 2-D skeleton images in one pass (BASELINE.json configs[2]).

The reference processes a list of images one by one (pipe.py:140-148: a thread pool around
``Skeleton`` + ``summarize`` per file).  On a B200 a 512 x 512 image is far too small to fill the
device and a per-image pipeline would be pure launch overhead, so a stack ``(B, H, W)`` is pushed
through the same kernels as ONE image whose leading axis carries no neighbours
(``skb_raw_neighbors(..., planes_independent=1)``).  Raster order makes every image's nodes,
paths and components contiguous id ranges, so per-image results are slices plus an offset.
Independent images shard across GPUs without any collective: give each rank its part of the stack.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _pipeline
from .csr import _default_device, _to_host


def stack_device(images, *, spacing=1, device=None, pack_variant=None):
    """All device results for a stack of images: namespace(graph, paths, dist, mean, stdev, table,
    node_offsets) -- node_offsets[i] = id of the first node of image i (host list, B + 1 entries)."""
    if device is None:
        device = images.device if isinstance(images, torch.Tensor) and images.device.type == 'cuda' else _default_device()
    shape = tuple(images.shape)
    if len(shape) != 3:
        raise ValueError('expected a stack of 2-D images with shape (B, H, W)')
    B, H, W = shape
    r = _pipeline.run_native(images, spacing=spacing, device=device, pack_variant=pack_variant, image_stack=True)
    g, p = r.graph, r.paths
    # id of the first node of every image: nodes are in raster order, so count the nodes before each image
    lin = g.lin
    bounds = torch.arange(B + 1, dtype=torch.int64, device=lin.device) * (H * W)
    node_off = torch.searchsorted(lin, bounds).tolist()
    from types import SimpleNamespace
    return SimpleNamespace(graph=g, paths=p, dist=r.dist, mean=r.mean, stdev=r.stdev, table=r.table,
                           spacing_is_int=r.spacing_is_int, node_offsets=node_off, n_images=B, counts=r.counts)


def summarize_images(images, *, spacing=1, separator='_', device=None):
    """``summarize(Skeleton(img, spacing=spacing))`` for every image of a stack, as one DataFrame with
    an extra leading column ``image_id``; node ids and skeleton ids are per image, rows are in the
    reference's per-image order.  Images without any path contribute no rows (the reference raises
    for such an image, csr.py:442)."""
    import pandas as pd
    r = stack_device(images, spacing=spacing, device=device)
    p = r.paths
    o32, o64, of = (_to_host(t) for t in r.table)
    node_off = np.asarray(r.node_offsets, dtype=np.int64)
    src = o32[1].astype(np.int64)
    image_id = np.searchsorted(node_off, src, side='right') - 1
    is_cycle = (np.arange(p.n_paths) >= p.n_regular).astype(np.int64)
    order = np.argsort(image_id * 2 + is_cycle, kind='stable')
    comp_rank = _to_host(p.component_rank)
    comp_off = comp_rank[node_off[:-1]].astype(np.int64)
    real = of.view(np.int64) if r.spacing_is_int else of
    cols = {
        'image_id': image_id,
        'skeleton_id': (o32[0] - comp_off[image_id]).astype(np.int32),
        'node_id_src': (src - node_off[image_id]).astype(np.int32),
        'node_id_dst': (o32[2] - node_off[image_id]).astype(np.int32),
        'branch_distance': _to_host(r.dist),
        'branch_type': o64[0],
        'mean_pixel_value': _to_host(r.mean),
        'stdev_pixel_value': _to_host(r.stdev),
        'image_coord_src_0': o64[2], 'image_coord_src_1': o64[3],
        'image_coord_dst_0': o64[5], 'image_coord_dst_1': o64[6],
        'coord_src_0': real[1], 'coord_src_1': real[2],
        'coord_dst_0': real[4], 'coord_dst_1': real[5],
        'euclidean_distance': of[6],
    }
    df = pd.DataFrame({k: np.asarray(v)[order] for k, v in cols.items()})
    df.rename(columns=lambda s: s.replace('_', separator) if s != 'image_id' else s, inplace=True)
    return df
