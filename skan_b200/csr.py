"""Drop-in mirror of ``skan.csr`` for the skeleton-to-graph hot path, computed on a B200.

Same names, signatures, attribute dtypes and error behaviour as the reference
(/root/reference/src/skan/csr.py): ``Skeleton`` (csr.py:565-858), ``summarize`` (csr.py:861-959),
``skeleton_to_csgraph`` (csr.py:1028-1089) and ``PathGraph`` (csr.py:449-551).  All numeric
work runs in the sm_100a kernels behind the C-ABI (see _pipeline.py); this module only wraps
device results in the SciPy / pandas containers the reference returns.  Host copies are made
lazily, the first time an attribute is read.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
from scipy import sparse

from . import _native, _pipeline


def _default_device():
    lib = _native.library()
    if lib.device_type != 'cuda':
        return torch.device('cpu')  # host emulation harness installed by tests/
    if not torch.cuda.is_available():
        raise _native.NativeError('skan_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _device_of(image):
    if isinstance(image, torch.Tensor) and image.device.type == 'cuda':
        return image.device
    return _default_device()


def _to_host(t):
    return t.cpu().numpy()


def _weighted_abs_diff(values0, values1, distances):
    """Default edge function of a complete image graph: abs(values0 - values1) * distances (csr.py:24-48)."""
    return np.abs(values0 - values1) * distances


def pixel_graph(image, *, mask=None, edge_function=None, connectivity=1, spacing=None):
    """Adjacency graph of the pixels of an image (csr.py:51-158): the raw 3^d stencil graph of the mask, before
    any junction clean-up.  Returns ``(graph, nodes)``: a csr_matrix whose entries are the distances between
    neighbouring mask pixels (or ``edge_function(values, neighbour values, distances)``) and the raveled indices
    of the mask pixels.

    Mask sweep, node ids, neighbourhood stencil and count-scan-emit CSR run on the device (the stage entry points
    of the hot path, with the neighbour directions restricted to ``connectivity``); a user-supplied
    ``edge_function`` is a Python callable and is evaluated on the host on the per-edge value arrays, as in the
    reference.
    """
    image = _to_host(image) if isinstance(image, torch.Tensor) else np.asarray(image)
    if image.dtype == bool and mask is None:
        mask = image
    if mask is None and edge_function is None:
        mask = np.ones_like(image, dtype=bool)
        edge_function = _weighted_abs_diff
    if mask is None:
        raise TypeError('pixel_graph needs a mask when an edge_function is given for a non-boolean image '
                        '(the reference fails in np.pad(None), csr.py:122)')
    mask = np.ascontiguousarray(mask, dtype=bool)
    ndim = mask.ndim
    if ndim < 1 or ndim > 3:
        raise NotImplementedError('skan_b200 supports 1-D, 2-D and 3-D images')
    g = _pipeline.graph_nodes(mask, spacing=1 if spacing is None else spacing, device=_default_device())
    if connectivity < ndim:
        # scipy.ndimage.generate_binary_structure(ndim, connectivity): at most `connectivity` non-zero offsets
        offs = np.stack([a.ravel() for a in np.meshgrid(*([np.arange(-1, 2)] * 3), indexing='ij')], axis=-1)
        allowed = int(sum(1 << k for k in range(27) if 0 < np.count_nonzero(offs[k]) <= max(int(connectivity), 0)))
        g.ctx.call('skb_mask_neighbors', g.raw_nb, g.n_nodes, allowed)
    empty32 = g.ctx.empty(0, torch.int32)
    _pipeline.graph_csr(g, empty32, empty32, g.ctx.empty(0, torch.uint8), g.ctx.empty(0, torch.uint8), records=False)
    n = g.n_nodes
    nodes = _to_host(g.lin)
    indptr, indices, data = _to_host(g.indptr), _to_host(g.indices), _to_host(g.data)
    if edge_function is not None:
        image_r = image.reshape(-1)
        rows = np.repeat(nodes, np.diff(indptr))
        data = edge_function(image_r[rows], image_r[nodes[indices]], data)
    return sparse.csr_matrix((data, indices, indptr), shape=(n, n)), nodes


def make_degree_image(skeleton_image):
    """Image of the degree of connectivity of every skeleton pixel under the full 3^d neighbourhood
    (csr.py:1173-1208): mask sweep + neighbourhood stencil on the device, scattered back into an int64 image."""
    g = _pipeline.graph_nodes(skeleton_image, device=_device_of(skeleton_image))
    out = torch.zeros(max(g.nvox, 1), dtype=torch.int64, device=g.ctx.device)
    g.ctx.call('skb_degree_image', g.raw_nb, g.lin, g.n_nodes, out)
    return _to_host(out[:g.nvox]).reshape(g.shape)


def submatrix(M, idxs):
    """The outer-index product submatrix ``M[idxs, :][:, idxs]`` (csr.py:1145-1170)."""
    return M[idxs, :][:, idxs]


def distance_with_height(source_values, neighbor_values, distances):
    """Edge function of value_is_height: hypot(height difference, distance) (csr.py:1023-1025)."""
    return np.hypot(source_values - neighbor_values, distances)


def csr_to_nbgraph(csr, node_props=None):
    """Host view of a CSR matrix with the reference's NBGraph fields (csr.py:206-216)."""
    from ._nbgraph import csr_to_nbgraph as _impl
    return _impl(csr, node_props)


def skeleton_to_csgraph(skel, *, spacing=1, value_is_height=False):
    """Convert a skeleton image of thin lines to a graph of neighbor pixels (csr.py:1028-1089).

    Returns ``(graph, pixel_coordinates)``: a ``scipy.sparse.csr_matrix`` of shape (N, N) with
    fp64 distances (int32 index arrays, sorted columns) and a tuple of ndim int64 coordinate
    arrays, node ids being 0-based raster ranks of the nonzero pixels.
    """
    g = _pipeline.run_native(skel, spacing=spacing, value_is_height=value_is_height, device=_device_of(skel),
                             graph_only=True).graph
    n = g.n_nodes
    graph = sparse.csr_matrix((_to_host(g.data), _to_host(g.indices), _to_host(g.indptr)), shape=(n, n))
    coords = _to_host(g.coords)
    return graph, tuple(np.ascontiguousarray(coords[:, i]) for i in range(g.ndim))


def _pixel_values_host(values_dev, np_dtype):
    """csr.py:554-562: None for bool images, float64 for integer images, the image dtype for floats."""
    if values_dev is None:
        return None
    v = _to_host(values_dev)
    if np_dtype.kind == 'f' and np_dtype != np.float64:
        return v.astype(np_dtype)
    return v


class Skeleton:
    """Object to group together all the properties of a skeleton (csr.py:565-858).

    Parameters and attributes follow the reference: ``graph`` (csr_matrix (N, N), fp64 data,
    int32 indices), ``coordinates`` (N, ndim) int64, ``paths`` (csr_matrix (P, L)), ``n_paths``,
    ``distances``, ``degrees`` (int32, post junction-MST), ``pixel_values``, ``spacing``,
    ``skeleton_shape``, ``skeleton_dtype``, ``skeleton_image``, ``source_image``.
    ``skeleton_image`` may be a NumPy array, a (pinned) CPU torch tensor or a CUDA tensor.
    """

    def __init__(self, skeleton_image, *, spacing=1, source_image=None, keep_images=True,
                 value_is_height=False):
        # the whole path (graph, paths, statistics, branch table) in one native call
        run = _pipeline.run_native(skeleton_image, spacing=spacing, value_is_height=value_is_height,
                                   device=_device_of(skeleton_image))
        g = run.graph
        self._run = run
        self._g = g
        self._ctx = g.ctx
        self._values_dev = g.values
        self._coords_dev = g.coords
        self._graph_dev = (g.indptr, g.indices, g.data)
        self._n_nodes = g.n_nodes
        self._p = run.paths
        self.n_paths = self._p.n_paths
        self._host = {}
        self._stats = None
        self._value_is_height = bool(value_is_height)
        self.skeleton_image = None
        self.skeleton_shape = tuple(g.shape)
        self.skeleton_dtype = g.np_dtype
        self.source_image = None
        self.spacing = (np.asarray(spacing) if not np.isscalar(spacing) else np.full(g.ndim, spacing))
        if keep_images:
            self.keep_images = keep_images
            self.skeleton_image = skeleton_image
            self.source_image = source_image

    @classmethod
    def from_path_graph(cls, pg):
        """A Skeleton from a PathGraph (csr.py:670-688): same accessors and ``summarize`` on an arbitrary graph.
        The reference builds a dummy 4^ndim image first and overwrites its attributes; here the object is
        filled from the PathGraph directly (device handles included), with the same resulting attributes."""
        d = pg._dev
        obj = object.__new__(cls)
        coords = np.asarray(pg.node_coordinates)
        obj._run = None
        obj._g = None
        obj._ctx = d['ctx']
        obj._values_dev = d['values']
        obj._coords_dev = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int64)).to(obj._ctx.device)
        obj._graph_dev = (d['indptr'], d['indices'], d['data'])
        obj._n_nodes = int(pg.adj.shape[0])
        obj._p = d['p']
        obj.n_paths = pg.n_paths
        obj._host = {'graph': pg.adj, 'coordinates': coords, 'pixel_values': pg.node_values, 'paths': pg.paths,
                     'degrees': pg.degrees, 'distances': pg.distances}
        obj._stats = None
        obj._value_is_height = False
        obj.skeleton_image = None
        obj.source_image = None
        obj.skeleton_shape = np.max(coords, axis=0) + 1
        obj.skeleton_dtype = bool if pg.node_values is None else np.asarray(pg.node_values).dtype
        obj.spacing = pg.spacing
        return obj

    # ---- lazily materialised host views of the device results ---------------------------
    def _cached(self, key, make):
        if key not in self._host:
            self._host[key] = make()
        return self._host[key]

    @property
    def graph(self):
        def make():
            indptr, indices, data = self._graph_dev
            n = self._n_nodes
            return sparse.csr_matrix((_to_host(data), _to_host(indices), _to_host(indptr)), shape=(n, n))
        return self._cached('graph', make)

    @graph.setter
    def graph(self, value):
        self._host['graph'] = value

    @property
    def coordinates(self):
        return self._cached('coordinates', lambda: _to_host(self._coords_dev))

    @coordinates.setter
    def coordinates(self, value):
        self._host['coordinates'] = value

    @property
    def degrees(self):
        return self._cached('degrees', lambda: np.diff(self.graph.indptr))

    @degrees.setter
    def degrees(self, value):
        self._host['degrees'] = value

    @property
    def pixel_values(self):
        return self._cached('pixel_values', lambda: _pixel_values_host(self._values_dev, np.dtype(self.skeleton_dtype)))

    @pixel_values.setter
    def pixel_values(self, value):
        self._host['pixel_values'] = value

    @property
    def paths(self):
        def make():
            p = self._p
            return sparse.csr_matrix((_to_host(p.data), _to_host(p.indices), _to_host(p.indptr)),
                                     shape=(p.n_paths, p.n_entries))
        return self._cached('paths', make)

    @paths.setter
    def paths(self, value):
        self._host['paths'] = value

    @property
    def nbgraph(self):
        """numba view of ``graph`` (csr.py:650), built on demand for code that wants it."""
        def make():
            from ._nbgraph import csr_to_nbgraph
            return csr_to_nbgraph(self.graph, self.pixel_values)
        return self._cached('nbgraph', make)

    def _device_stats(self):
        if self._stats is None:
            if self._run is not None:
                self._stats = (self._run.dist, self._run.mean, self._run.stdev)
            else:   # built from a PathGraph: no native run behind it
                self._stats = tuple(_pipeline.path_stats(self._ctx, self._p))
        return self._stats

    @property
    def distances(self):
        return self.path_lengths()

    # ---- reference accessors ---------------------------------------------------------------
    def path(self, index):
        """Pixel indices of path number `index`, endpoints included (csr.py:690-714)."""
        start, stop = self.paths.indptr[index:index + 2]
        return self.paths.indices[start:stop]

    def path_coordinates(self, index):
        """Image coordinates of the pixels in the path (csr.py:716-730)."""
        return self.coordinates[self.path(index)]

    def path_with_data(self, index):
        """Pixel indices and pixel values on a path (csr.py:732-748)."""
        start, stop = self.paths.indptr[index:index + 2]
        return self.paths.indices[start:stop], self.paths.data[start:stop]

    def path_lengths(self):
        """Length of each path (csr.py:750-764; sum of edge weights left to right, csr.py:962-977)."""
        return self._cached('distances', lambda: _to_host(self._device_stats()[0]))

    def paths_list(self):
        """All the paths, endpoints included (csr.py:766-774)."""
        return [list(self.path(i)) for i in range(self.n_paths)]

    def path_means(self):
        """Mean pixel value along each path (csr.py:792-802)."""
        return self._cached('means', lambda: _to_host(self._device_stats()[1]))

    def path_stdev(self):
        """Standard deviation of pixel values along each path (csr.py:804-816)."""
        return self._cached('stdev', lambda: _to_host(self._device_stats()[2]))

    def path_label_image(self):
        """Image of the skeleton's shape holding path id + 1 at every path pixel (csr.py:776-790)."""
        ctx, p = self._ctx, self._p
        shape = tuple(int(x) for x in self.skeleton_shape)
        v = int(np.prod(shape, dtype=np.int64))
        out = torch.zeros(v, dtype=torch.int64, device=ctx.device)
        if self._g is not None:
            lin = self._g.lin
        else:   # built from a PathGraph: raveled indices from the node coordinates
            lin = torch.from_numpy(np.ravel_multi_index(tuple(self.coordinates.T), shape).astype(np.int64)).to(ctx.device)
        ctx.call('skb_path_label_image', p.indptr, p.indices, lin, p.n_paths, out)
        return _to_host(out).reshape(shape)

    def prune_paths(self, indices):
        """Prune paths from the skeleton (csr.py:818-854): the non-junction pixels of the given paths are wiped from a copy
        of the skeleton image, the image is re-skeletonized and a new Skeleton is built from it.  The re-thinning is the
        reference's own call of ``skimage.morphology.skeletonize`` on the host (pre-processing, outside the hot path:
        SURVEY.md section 8 (f5)); where scikit-image is not installed this raises NotImplementedError after the
        reference's index check."""
        if not np.all(np.array(indices) < self.n_paths):
            raise ValueError(
                f'The path index {np.max(indices)} does not exist in this '
                f'skeleton. (The highest path index is {self.n_paths}.)\n'
                'If you obtained the index from a summary table, you '
                'probably need to resummarize the skeleton.')
        try:
            from skimage import morphology
        except ImportError:
            raise NotImplementedError('prune_paths needs skimage.morphology.skeletonize (scikit-image is not installed; '
                                      'the thinning is out of this package\'s scope)') from None
        if self.skeleton_image is None:
            raise ValueError('prune_paths needs the skeleton image (Skeleton(..., keep_images=True))')
        image = self.skeleton_image
        if isinstance(image, torch.Tensor):   # the reference's images are NumPy arrays; ours may live on a device
            image = image.detach().cpu().numpy()
        image_cp = np.copy(image)
        degrees = self.degrees
        coordinates = self.coordinates
        for i in indices:
            pixel_ids_to_wipe = self.path(i)
            junctions = degrees[pixel_ids_to_wipe] > 2
            pixel_ids_to_wipe = pixel_ids_to_wipe[~junctions]
            coords_to_wipe = coordinates[pixel_ids_to_wipe]
            coords_idxs = tuple(np.round(coords_to_wipe).astype(int).T)
            image_cp[coords_idxs] = 0
        new_skeleton = morphology.skeletonize(image_cp.astype(bool)) * image_cp
        return Skeleton(new_skeleton, spacing=self.spacing, source_image=self.source_image,
                        keep_images=self.keep_images)

    def __array__(self, dtype=None):
        """Array representation of the skeleton path labels (csr.py:856-858)."""
        return self.path_label_image()


class PathGraph:
    """Generalization of a skeleton (csr.py:449-551): the same path pipeline from an arbitrary
    symmetric CSR adjacency (``from_graph``) or from an image (``from_image``)."""

    def __init__(self, adj, node_coordinates, node_values, paths, spacing=1, _dev=None):
        self.adj = adj
        self.node_coordinates = node_coordinates
        self.node_values = node_values
        self.paths = paths
        self.spacing = spacing
        self._dev = _dev
        self._distances = None

    @classmethod
    def from_graph(cls, *, adj, node_coordinates, node_values=None, spacing=1):
        ctx = _pipeline._Ctx(_default_device())
        adj = sparse.csr_matrix(adj) if not sparse.issparse(adj) or adj.format != 'csr' else adj
        n = adj.shape[0]
        indptr = torch.from_numpy(np.ascontiguousarray(adj.indptr, dtype=np.int32)).to(ctx.device)
        indices = torch.from_numpy(np.ascontiguousarray(adj.indices, dtype=np.int32)).to(ctx.device)
        data = torch.from_numpy(np.ascontiguousarray(adj.data, dtype=np.float64)).to(ctx.device)
        values = None
        if node_values is not None:
            values = torch.from_numpy(np.ascontiguousarray(node_values, dtype=np.float64)).to(ctx.device)
        p = _pipeline.trace_paths(ctx, indptr, indices, data, values, n, int(adj.indices.size))
        paths = sparse.csr_matrix((_to_host(p.data), _to_host(p.indices), _to_host(p.indptr)),
                                  shape=(p.n_paths, p.n_entries))
        dev = dict(ctx=ctx, indptr=indptr, indices=indices, data=data, values=values, p=p)
        return cls(adj, node_coordinates, node_values, paths, spacing, _dev=dev)

    @classmethod
    def from_image(cls, skeleton_image, *, spacing=1, value_is_height=False):
        sk = Skeleton(skeleton_image, spacing=spacing, value_is_height=value_is_height, keep_images=False)
        dev = dict(ctx=sk._ctx, indptr=sk._graph_dev[0], indices=sk._graph_dev[1], data=sk._graph_dev[2],
                   values=sk._values_dev, p=sk._p)
        return cls(sk.graph, sk.coordinates, sk.pixel_values, sk.paths, spacing, _dev=dev)

    @property
    def distances(self):
        if self._distances is None:
            d = self._dev
            dist, _, _ = _pipeline.path_stats(d['ctx'], d['p'], mean=False, stdev=False)
            self._distances = _to_host(dist)
        return self._distances

    @property
    def n_paths(self):
        return self.paths.shape[0]

    @property
    def degrees(self):
        return np.diff(self.adj.indptr)


def summarize(skel, *, value_is_height=False, find_main_branch=False, separator=None):
    """Compute statistics for every skeleton and branch in ``skel`` (csr.py:861-959).

    Returns a pandas.DataFrame with the reference's columns, order and dtypes: skeleton_id,
    node_id_src/dst (int32), branch_distance, branch_type (int64), mean/stdev_pixel_value,
    image_coord_src/dst_i (int64), coord_src/dst_i, euclidean_distance.
    """
    import pandas as pd
    if isinstance(skel, PathGraph):
        skel = Skeleton.from_path_graph(skel)
    if separator is None:
        warnings.warn(
            "separator in column name will change to _ in version 0.13; "
            "to silence this warning, use `separator='-'` to maintain "
            "current behavior and use `separator='_'` to switch to the "
            "new default behavior.",
            np.exceptions.VisibleDeprecationWarning,
            stacklevel=2,
        )
        separator = '-'
    ctx, p = skel._ctx, skel._p
    indptr, indices, data = skel._graph_dev
    ndim = int(skel._coords_dev.shape[1])
    spacing = skel.spacing
    if value_is_height and skel._values_dev is None:
        raise TypeError("'NoneType' object is not subscriptable")  # csr.py:933 on a boolean skeleton
    run = getattr(skel, '_run', None)
    if run is not None and bool(value_is_height) == skel._value_is_height:
        dist, mean, stdev = run.dist, run.mean, run.stdev      # computed by the native run already
        (o32, o64, of), sp_is_int = run.table, run.spacing_is_int
    else:
        dist, mean, stdev = _pipeline.path_stats(ctx, p)
        o32, o64, of, sp_is_int = _pipeline.branch_table(ctx, p, indptr, skel._coords_dev, skel._values_dev,
                                                          spacing, ndim, value_is_height)
    o32, o64, of = _to_host(o32), _to_host(o64), _to_host(of)
    h = int(bool(value_is_height))
    dh = ndim + h
    summary = {}
    summary['skeleton_id'] = o32[0]
    summary['node_id_src'] = o32[1]
    summary['node_id_dst'] = o32[2]
    summary['branch_distance'] = _to_host(dist)
    summary['branch_type'] = o64[0]
    summary['mean_pixel_value'] = _to_host(mean)
    summary['stdev_pixel_value'] = _to_host(stdev)
    for i in range(ndim):
        summary[f'image_coord_src_{i}'] = o64[1 + i]
    for i in range(ndim):
        summary[f'image_coord_dst_{i}'] = o64[1 + ndim + i]
    real = of.view(np.int64) if sp_is_int else of
    vdtype = None
    if h:
        pv = skel.pixel_values
        vdtype = np.asarray(pv).dtype if pv is not None else np.float64
    for i in range(ndim):
        summary[f'coord_src_{i}'] = real[i]
    if h:
        summary[f'coord_src_{ndim}'] = of[ndim].astype(vdtype, copy=False)
    for i in range(ndim):
        summary[f'coord_dst_{i}'] = real[dh + i]
    if h:
        summary[f'coord_dst_{ndim}'] = of[dh + ndim].astype(vdtype, copy=False)
    summary['euclidean_distance'] = of[2 * dh]
    df = pd.DataFrame(summary)
    if find_main_branch:
        # host-side post-processing of the P-row table, like the reference (csr.py:955-957)
        from .summary_utils import find_main_branches
        df['main'] = find_main_branches(df)
    df.rename(columns=lambda s: s.replace('_', separator), inplace=True)
    return df


# Sholl analysis (csr.py:1211-1390): device kernels for the node / edge part, see _sholl.py
from ._sholl import _fast_graph_center_idx, _simplify_graph, sholl_analysis  # noqa: E402,F401


def skeleton_to_nx(skeleton, summary=None):
    """Convert a Skeleton to a networkx MultiGraph (csr.py:1393-1435): one node per junction / endpoint, one edge
    per path carrying its pixel coordinates ('path'), node ids ('indices') and pixel values ('values').  A host
    container built from the P-row branch table, like in the reference."""
    import networkx as nx
    if summary is None:
        summary = summarize(skeleton, separator='_')
    csum = summary.rename(columns=lambda s: s.replace('-', '_'))
    g = nx.MultiGraph(shape=skeleton.skeleton_shape, dtype=skeleton.skeleton_dtype)
    for row in csum.itertuples(name='Edge'):
        index = row.Index
        indices, values = skeleton.path_with_data(index)
        g.add_edge(row.node_id_src, row.node_id_dst,
                   **{'path': skeleton.path_coordinates(index), 'indices': indices, 'values': values})
    return g


def nx_to_skeleton(g):
    """Convert a networkx graph with 'shape' / 'dtype' graph attributes and 'path' / 'values' edge attributes back
    to a Skeleton (csr.py:1438-1480): the image is regenerated from the paths and skeletonized to a graph again."""
    image = np.zeros(g.graph['shape'], dtype=g.graph['dtype'])
    all_coords = np.concatenate([path for _, _, path in g.edges.data('path')], axis=0)
    all_values = np.concatenate([values for _, _, values in g.edges.data('values')], axis=0)
    image[tuple(all_coords.T)] = all_values
    return Skeleton(image)
