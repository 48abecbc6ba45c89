"""Drop-in mirror of ``skan.image_stats`` (reference src/skan/image_stats.py), SURVEY.md section 8(f3).

``mesh_sizes`` labels the 4-connected background regions of a 2-D skeleton image and returns the areas of those
that do not touch the border (image_stats.py:8-46); ``image_summary`` puts them next to the junction count of
``skeleton_to_csgraph`` and the number of connected skeletons (image_stats.py:49-90).  The pixel-proportional
work runs in the CUDA library: the mask sweep of the hot path, then a run-based connected-component labelling
of the zero bits of the packed mask (skan_b200/csrc/skb_mesh.cuh) -- no label image is ever built.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _pipeline
from .csr import _device_of, skeleton_to_csgraph


def _packed_mask(skeleton, device):
    """Image -> (ctx, bits tensor with pad words, nvox, number of foreground pixels) through the mask sweep."""
    ctx = _pipeline._Ctx(device)
    img, code, shape, _ = _pipeline._as_device_image(skeleton, ctx.device)
    nvox = int(np.prod(shape, dtype=np.int64))
    tile = ctx.lib.tile_voxels
    ntiles = (nvox + tile - 1) // tile
    store = torch.zeros(ntiles * (tile // 64) + 18, dtype=torch.int64, device=ctx.device)
    bits = store[16:-2]
    counts = torch.empty(ntiles + 1, dtype=torch.int32, device=ctx.device)
    variant = min(1, _pipeline.PACK_VARIANT[_pipeline.default_pack_variant])   # LDG or TMA sweep (the fused one emits nodes too)
    if nvox:
        ctx.call('skb_pack_count', img, code, nvox, bits, counts, ntiles, variant)
    else:
        counts.zero_()
    return ctx, store, bits, nvox, shape, counts[:ntiles]


def mesh_sizes(skeleton):
    """Areas in pixels of the spaces between skeleton branches that do not touch the image border
    (image_stats.py:8-46).  2-D images only, like the reference ("this only makes sense for 2D images").

    Reference quirk, reproduced: ``np.bincount(labeled.flat)`` also counts label 0, i.e. the skeleton pixels
    themselves, and that entry is only zeroed when a skeleton pixel lies on the image border
    (image_stats.py:35-44).  For a skeleton that stays clear of the border the first element of the result is
    therefore the number of skeleton pixels.
    """
    shape = tuple(skeleton.shape)
    if len(shape) != 2:
        raise ValueError('mesh_sizes only makes sense for 2-D images (reference image_stats.py:11)')
    ctx, store, bits, nvox, shape, tile_counts = _packed_mask(skeleton, _device_of(skeleton))
    ny, nx = int(shape[0]), int(shape[1])
    if nvox == 0:
        return np.zeros(0, dtype=np.int64)
    nwords = (nvox + 63) // 64
    i32, i64, u8 = torch.int32, torch.int64, torch.uint8
    rs = torch.empty(nwords, dtype=i64, device=ctx.device)
    rsp = torch.empty(nwords + 1, dtype=i32, device=ctx.device)
    ctx.call('skb_mesh_runs', bits, ny, nx, rs, rsp, ctx.scratch(nwords + 1))
    n_runs = int(rsp[nwords].item())
    n_fg = int(tile_counts.sum(dtype=i64).item())
    flag = torch.zeros(4, dtype=i32, device=ctx.device)
    ctx.call('skb_mesh_border_foreground', bits, ny, nx, flag)
    sizes = np.zeros(0, dtype=np.int64)
    if n_runs:
        par = torch.empty(n_runs, dtype=i32, device=ctx.device)
        length = torch.empty(n_runs, dtype=i32, device=ctx.device)
        border = torch.empty(n_runs, dtype=u8, device=ctx.device)
        size = torch.zeros(n_runs, dtype=i64, device=ctx.device)
        root_border = torch.zeros(n_runs, dtype=i32, device=ctx.device)
        keepoff = torch.empty(n_runs + 1, dtype=i32, device=ctx.device)
        ctx.call('skb_mesh_link', bits, ny, nx, rs, rsp, n_runs, par, length, border, size, root_border, keepoff,
                 ctx.scratch(n_runs + 1))
        n_keep = int(keepoff[n_runs].item())
        out = torch.empty(max(n_keep, 1), dtype=i64, device=ctx.device)
        ctx.call('skb_mesh_emit', par, root_border, keepoff, size, n_runs, out)
        sizes = out[:n_keep].cpu().numpy()
    fg_on_border = bool(flag[0].item())
    if n_fg and not fg_on_border:
        sizes = np.concatenate([np.array([n_fg], dtype=np.int64), sizes])
    return sizes


def _count_skeletons(skeleton, graph):
    """Number of connected components of the pixel graph, isolated pixels included: what
    ``ndi.label(skeleton, np.ones((3,) * ndim))[1]`` counts (image_stats.py:86-87; the junction MST removes edges
    inside clusters only, never a connection)."""
    ctx = _pipeline._Ctx(_device_of(skeleton))
    n = graph.shape[0]
    if n == 0:
        return 0
    indptr = torch.from_numpy(np.ascontiguousarray(graph.indptr, dtype=np.int32)).to(ctx.device)
    indices = torch.from_numpy(np.ascontiguousarray(graph.indices, dtype=np.int32)).to(ctx.device)
    par = torch.empty(n, dtype=torch.int32, device=ctx.device)
    count = torch.zeros(1, dtype=torch.int64, device=ctx.device)
    ctx.call('skb_cc_count', indptr, indices, n, par, count)
    return int(count.item())


def image_summary(skeleton, *, spacing=1):
    """Summary statistics of a skeleton image as a one-row pandas.DataFrame (image_stats.py:49-90): scale, number of
    junctions (degree > 2 after the junction clean-up), image area in real units, junction density, mean / median /
    standard deviation of the mesh areas and the number of disjoint skeletons -- same column names, order and dtypes."""
    import pandas as pd
    graph, _ = skeleton_to_csgraph(skeleton, spacing=spacing)
    shape = tuple(skeleton.shape)
    n_junctions = np.sum(np.diff(graph.indptr) > 2)
    area = np.prod(shape) * (spacing**len(shape) if np.isscalar(spacing) else np.prod(spacing))
    meshes = mesh_sizes(skeleton)
    columns = [
        ('number of junctions', n_junctions),
        ('area', area),
        ('junctions per unit area', None),          # derived below from the two columns, like the reference
        ('average mesh area', np.mean(meshes)),
        ('median mesh area', np.median(meshes)),
        ('mesh area standard deviation', np.std(meshes)),
        ('number of disjoint skeletons', _count_skeletons(skeleton, graph)),
    ]
    stats = pd.DataFrame()
    stats['scale'] = [spacing]
    for name, value in columns:
        stats[name] = stats['number of junctions'] / stats['area'] if value is None else value
    return stats
