"""CPU oracle for the skeleton-to-graph hot path of jni/skan.

TEST INFRASTRUCTURE ONLY.  Nothing in ``skan_b200/`` may import this module: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker (or as the thing timed on the host cores), never as
the product path.

This is an independent restatement, in NumPy/SciPy (+ numba for the two serial loops the
reference also JIT-compiles), of what the reference computes on this path.  Every function
cites the reference lines (relative to /root/reference) it restates.  Parity is PINNED:
``tests/test_oracle_golden.py`` checks this module against golden vectors produced by the
unmodified reference (``tests/golden/make_golden.py`` imports /root/reference/src/skan in the
build container and stores its outputs in ``tests/golden/*.npz``), including the reference's
own fixtures (src/skan/_testdata.py:6-75) and the known answers of its tests
(src/skan/test/test_csr.py, test_skeleton_class.py).

Third-party arithmetic the reference reaches on this path and that is not in its tree:
``scipy.sparse`` (coo->csr canonicalisation, sparse +/-), ``scipy.sparse.csgraph``
(``minimum_spanning_tree``, ``connected_components``; SciPy 1.18.1 here, reference pins
``scipy>=1.7``) and ``skimage.util.map_array`` (pure integer lookup; scikit-image>=0.17.1,
absent here).  ``mst_junctions`` below calls the same SciPy routines the reference calls;
``mst_junctions_rule`` restates their *result* as an explicit Kruskal under the strict total
order (weight, min id, max id) -- that rule is what the CUDA path implements and the test
suite checks both against the reference's golden outputs.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.sparse import csgraph

try:  # the reference JIT-compiles its serial loops with numba (csr.py:347,390,962,970)
    import numba
    _jit = numba.njit(cache=False)
    HAVE_NUMBA = True
except Exception:  # pragma: no cover - numba is in the image; keep a slow fallback
    HAVE_NUMBA = False

    def _jit(f):
        return f


# --------------------------------------------------------------------------------------
# neighbourhood table  (src/skan/nputil.py:188-295, called at src/skan/csr.py:124)
# --------------------------------------------------------------------------------------
def neighbour_table(shape, spacing):
    """All 3^d-1 neighbour offsets of the full neighbourhood (connectivity = ndim,
    csr.py:1083), as coordinate offsets (K, d), raveled offsets for an array of ``shape``
    (K,), and fp64 distances (K,) = sqrt(sum((offset*spacing)**2)) exactly as
    nputil.py:278-279 writes it.  Rows are ordered by ascending raveled offset for the
    *unclipped* C-order lattice (dz, dy, dx lexicographic); the reference sorts by distance
    (nputil.py:280-281) but that order never reaches a result because coo->csr
    canonicalises (csr.py:157)."""
    d = len(shape)
    spacing = np.ones(d, dtype=float) * spacing          # csr.py:1072
    grids = np.meshgrid(*([np.arange(-1, 2)] * d), indexing='ij')
    offs = np.stack([g.ravel() for g in grids], axis=-1)  # lexicographic
    offs = offs[np.any(offs != 0, axis=1)]
    strides = np.cumprod((tuple(shape[1:]) + (1,))[::-1])[::-1]  # nputil.py:270-271
    raveled = (offs * strides).sum(axis=1)
    weighted = offs * spacing
    dist = np.sqrt(np.sum(weighted**2, axis=1))           # nputil.py:278-279
    return offs, raveled, dist


# --------------------------------------------------------------------------------------
# pixel graph  (src/skan/csr.py:51-158) + skeleton_to_csgraph front half (csr.py:1068-1085)
# --------------------------------------------------------------------------------------
def raw_pixel_graph(image, spacing=1, value_is_height=False):
    """Raster-ordered node ids (csr.py:131-132), full-neighbourhood adjacency (csr.py:122-144),
    edge data = table distance, or hypot(height difference, distance) in height mode
    (csr.py:148-151, 1023-1025), canonical CSR (csr.py:153-157).  Returns (csr, nodes)
    where nodes are the raveled indices of the foreground voxels."""
    image = np.asarray(image)
    mask = image.astype(bool)                             # csr.py:1070
    shape = mask.shape
    d = mask.ndim
    nodes = np.flatnonzero(mask)                          # csr.py:131
    n = nodes.size
    offs, _, dist = neighbour_table(shape, spacing)
    coords = np.stack(np.unravel_index(nodes, shape), axis=-1) if n else np.zeros((0, d), np.int64)
    rows, cols, data = [], [], []
    flat_mask = mask.reshape(-1)
    flat_img = image.reshape(-1)
    shape_arr = np.asarray(shape)
    for k in range(offs.shape[0]):
        # instead of padding (csr.py:122) test bounds explicitly: same set of edges
        nb = coords + offs[k]
        inside = np.all((nb >= 0) & (nb < shape_arr), axis=1)
        src = np.flatnonzero(inside)
        nb_r = np.ravel_multi_index(tuple(nb[src].T), shape) if src.size else np.zeros(0, np.int64)
        hit = flat_mask[nb_r]
        src = src[hit]
        nb_r = nb_r[hit]
        dst = np.searchsorted(nodes, nb_r)                # the map_array lookup, csr.py:135,142
        rows.append(src)
        cols.append(dst)
        if value_is_height:
            # difference taken in the image dtype (csr.py:1024), then np.hypot (csr.py:1025)
            hd = flat_img[nodes[src]] - flat_img[nb_r]
            data.append(np.hypot(hd, dist[k]))
        else:
            data.append(np.full(src.size, dist[k]))
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    data = np.concatenate(data) if data else np.zeros(0, float)
    graph = sparse.coo_matrix((data, (rows, cols)), shape=(n, n)).tocsr()   # csr.py:153-157
    return graph, nodes


# --------------------------------------------------------------------------------------
# junction MST  (src/skan/csr.py:980-1020)
# --------------------------------------------------------------------------------------
def mst_junctions(graph):
    """Same SciPy calls as the reference: zero every entry that touches a node of raw degree
    < 3 (csr.py:1004-1016), minimum_spanning_tree of what is left (csr.py:1017), then remove
    the non-tree junction-junction entries from the graph (csr.py:1018-1019)."""
    graph = graph.tocsr()
    deg = np.diff(graph.indptr)                           # pattern row count == csr.py:1005 for a symmetric pattern
    coo = graph.tocoo()
    jj = (deg[coo.row] >= 3) & (deg[coo.col] >= 3)
    sub = sparse.csr_matrix((coo.data[jj], (coo.row[jj], coo.col[jj])), shape=graph.shape)
    sub.eliminate_zeros()                                 # csr.py:1016
    mst = csgraph.minimum_spanning_tree(sub)
    non_tree = sub - (mst + mst.T)
    final = graph - non_tree
    return final.tocsr()


def mst_junctions_rule(graph):
    """Explicit statement of the result of csr.py:980-1020: Kruskal over the
    undirected junction-junction edges under the strict total order (weight, min id, max id),
    fp64 '<' on the stored weights; edges among junctions that are not in that forest are
    deleted in both directions, all data values unchanged.  Valid for symmetric weights."""
    graph = graph.tocsr()
    indptr, indices, data = graph.indptr, graph.indices, graph.data
    n = graph.shape[0]
    deg = np.diff(indptr)
    rows = np.repeat(np.arange(n), deg)
    jj = (deg[rows] >= 3) & (deg[indices] >= 3)
    und = jj & (rows < indices)
    a, b, w = rows[und], indices[und], data[und]
    order = np.lexsort((b, a, w))
    keep = _kruskal(n, a[order].astype(np.int64), b[order].astype(np.int64))
    tree = set(zip(a[order][keep].tolist(), b[order][keep].tolist()))
    drop = np.zeros(indices.size, bool)
    for e in np.flatnonzero(jj):
        u, v = int(rows[e]), int(indices[e])
        if (min(u, v), max(u, v)) not in tree:
            drop[e] = True
    keep_e = ~drop
    new_deg = np.bincount(rows[keep_e], minlength=n)
    new_indptr = np.concatenate([[0], np.cumsum(new_deg)]).astype(indptr.dtype)
    return sparse.csr_matrix((data[keep_e], indices[keep_e], new_indptr), shape=graph.shape)


@_jit
def _kruskal(n, a, b):
    parent = np.arange(n)
    keep = np.zeros(a.size, np.bool_)
    for i in range(a.size):
        x = a[i]
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        y = b[i]
        while parent[y] != y:
            parent[y] = parent[parent[y]]
            y = parent[y]
        if x != y:
            parent[max(x, y)] = min(x, y)
            keep[i] = True
    return keep


def skeleton_to_csgraph(image, spacing=1, value_is_height=False):
    """csr.py:1028-1089: raw pixel graph -> junction MST -> (graph, coordinate tuple)."""
    image = np.asarray(image)
    raw, nodes = raw_pixel_graph(image, spacing, value_is_height)
    graph = mst_junctions(raw)
    coords = np.unravel_index(nodes, image.shape)          # csr.py:1088
    return graph, coords


def extract_values(image, coords):
    """csr.py:554-562: None for bool, ints -> float64, floats keep their dtype."""
    image = np.asarray(image)
    if image.dtype == np.bool_:
        return None
    vals = image[coords]
    if np.issubdtype(image.dtype, np.integer):
        return vals.astype(np.float64)
    return vals


# --------------------------------------------------------------------------------------
# path tracing  (src/skan/csr.py:347-446)
# --------------------------------------------------------------------------------------
@_jit
def _trace(indptr, indices, values, out_ptr, out_idx, out_val):
    """Serial walk restating csr.py:347-409.  Phase 1: every node whose degree is 1 or > 2,
    ascending; every neighbour ascending; follow an unvisited edge through degree-2 nodes
    until a node of degree != 2 or the start node.  Phase 2: what is still unvisited lies
    on junction-free cycles; start each at its lowest node toward its first neighbour."""
    n = indptr.size - 1
    seen = np.zeros(indices.size, np.bool_)
    npaths = 0
    fill = 0
    for phase in range(2):
        for a in range(n):
            dega = indptr[a + 1] - indptr[a]
            if phase == 0:
                if not (dega == 1 or dega > 2):
                    continue
                lo, hi = indptr[a], indptr[a + 1]
            else:
                if dega == 0:
                    continue
                lo, hi = indptr[a], indptr[a] + 1          # first neighbour only, csr.py:371
            for e in range(lo, hi):
                if seen[e]:
                    continue
                first = fill
                out_idx[fill] = a
                out_val[fill] = values[a]
                fill += 1
                prev = a
                ecur = e
                while True:
                    cur = indices[ecur]
                    seen[ecur] = True
                    # mark the opposite direction (csr.py:399-400)
                    for r in range(indptr[cur], indptr[cur + 1]):
                        if indices[r] == prev:
                            seen[r] = True
                            break
                    out_idx[fill] = cur
                    out_val[fill] = values[cur]
                    fill += 1
                    if indptr[cur + 1] - indptr[cur] != 2 or cur == a:
                        break
                    r0 = indptr[cur]
                    enext = r0 if indices[r0] != prev else r0 + 1
                    if seen[enext]:
                        break
                    prev = cur
                    ecur = enext
                npaths += 1
                out_ptr[npaths] = fill
                _ = first
    return npaths, fill


def build_paths(graph, values=None):
    """csr.py:412-446: returns (indptr int32 (P+1,), indices int32 (L,), data fp64 (L,))."""
    indptr = graph.indptr.astype(np.int64)
    indices = graph.indices.astype(np.int64)
    n = graph.shape[0]
    vals = np.ones(n, float) if values is None else np.asarray(values, dtype=np.float64)  # csr.py:207-215
    deg = np.diff(indptr)
    term_deg = deg[deg != 2]
    cap_paths = int(term_deg.sum()) + indices.size // 4 + 1          # csr.py:413-425
    cap_pts = int(indices.size + np.maximum(0, term_deg - 1).sum() + indices.size // 4 + 2)
    out_ptr = np.zeros(cap_paths + 1, np.int64)
    out_idx = np.zeros(cap_pts, np.int64)
    out_val = np.zeros(cap_pts, float)
    p, l = _trace(indptr, indices, vals, out_ptr, out_idx, out_val)
    return out_ptr[:p + 1].astype(np.int32), out_idx[:l].astype(np.int32), out_val[:l]


@_jit
def _distances(indptr, indices, data, pptr, pidx, out):
    """csr.py:962-977: left-to-right fp64 sum of graph.edge(u, v) along each path."""
    for p in range(out.size):
        acc = 0.0
        for j in range(pptr[p], pptr[p + 1] - 1):
            u = pidx[j]
            v = pidx[j + 1]
            w = 0.0
            for r in range(indptr[u], indptr[u + 1]):
                if indices[r] == v:
                    w = data[r]
                    break
            acc += w
        out[p] = acc


# --------------------------------------------------------------------------------------
# Skeleton + summarize  (src/skan/csr.py:565-858, 861-959)
# --------------------------------------------------------------------------------------
class OracleSkeleton:
    """Attribute contract of reference Skeleton (csr.py:634-668), host NumPy only."""

    def __init__(self, image, *, spacing=1, value_is_height=False):
        image = np.asarray(image)
        graph, coords = skeleton_to_csgraph(image, spacing, value_is_height)
        self._finish(graph, np.transpose(coords), extract_values(image, coords))
        self.skeleton_shape = image.shape
        self.skeleton_dtype = image.dtype
        self.spacing = (np.asarray(spacing) if not np.isscalar(spacing)
                        else np.full(image.ndim, spacing))                 # csr.py:661-664

    @classmethod
    def from_graph(cls, adj, node_coordinates, node_values=None, spacing=1):
        """PathGraph.from_graph (csr.py:477-499) + Skeleton.from_path_graph (csr.py:670-688)."""
        self = cls.__new__(cls)
        self._finish(adj.tocsr(), np.asarray(node_coordinates), node_values)
        self.skeleton_shape = tuple(np.max(self.coordinates, axis=0) + 1)
        self.skeleton_dtype = bool if node_values is None else np.asarray(node_values).dtype
        self.spacing = spacing
        return self

    def _finish(self, graph, coordinates, values):
        self.graph = graph
        self.coordinates = coordinates
        self.pixel_values = values
        pptr, pidx, pdata = build_paths(graph, values)
        if pptr.size - 1 == 0:
            # the reference's csr_matrix constructor raises for an empty index pointer (csr.py:442)
            raise ValueError('index pointer size 0 should be 1')
        self.paths = sparse.csr_matrix((pdata, pidx, pptr), shape=(pptr.size - 1, pidx.size))
        self.n_paths = pptr.size - 1
        self.degrees = np.diff(graph.indptr)                              # csr.py:660

    def path(self, i):
        s, e = self.paths.indptr[i:i + 2]
        return self.paths.indices[s:e]

    def paths_list(self):
        return [list(self.path(i)) for i in range(self.n_paths)]

    def path_lengths(self):
        out = np.empty(self.n_paths, float)
        _distances(self.graph.indptr.astype(np.int64), self.graph.indices.astype(np.int64),
                   self.graph.data.astype(np.float64), self.paths.indptr.astype(np.int64),
                   self.paths.indices.astype(np.int64), out)
        return out

    def path_means(self):                                                   # csr.py:792-802
        sums = np.add.reduceat(self.paths.data, self.paths.indptr[:-1])
        return sums / np.diff(self.paths.indptr)

    def path_stdev(self):                                                   # csr.py:804-816
        data = self.paths.data
        sumsq = np.add.reduceat(data * data, self.paths.indptr[:-1])
        n = np.diff(self.paths.indptr)
        m = self.path_means()
        return np.sqrt(np.clip(sumsq / n - m * m, 0, None))

    def path_label_image(self):                                             # csr.py:776-790
        out = np.zeros(self.skeleton_shape, dtype=int)
        for i in range(self.n_paths):
            c = self.coordinates[self.path(i)]
            out[tuple(np.round(c).astype(int).T)] = i + 1
        return out


def summarize(skel, *, value_is_height=False, separator='_'):
    """csr.py:861-959 -> ordered dict of NumPy columns (wrap in pandas.DataFrame to compare
    with the reference)."""
    out = {}
    d = skel.coordinates.shape[1]
    _, labels = csgraph.connected_components(skel.graph, directed=False)   # csr.py:909
    pptr, pidx = skel.paths.indptr, skel.paths.indices
    src = pidx[pptr[:-1]]
    dst = pidx[pptr[1:] - 1]
    out['skeleton_id'] = labels[src]
    out['node_id_src'] = src
    out['node_id_dst'] = dst
    out['branch_distance'] = skel.path_lengths()
    ds, dd = skel.degrees[src], skel.degrees[dst]
    kind = np.full(ds.shape, 2)
    kind[(ds == 1) | (dd == 1)] = 1
    kind[(ds == 1) & (dd == 1)] = 0
    kind[src == dst] = 3
    out['branch_type'] = kind
    out['mean_pixel_value'] = skel.path_means()
    out['stdev_pixel_value'] = skel.path_stdev()
    for i in range(d):
        out[f'image_coord_src_{i}'] = skel.coordinates[src, i]
    for i in range(d):
        out[f'image_coord_dst_{i}'] = skel.coordinates[dst, i]
    rs = skel.coordinates[src] * skel.spacing
    for i in range(d):
        out[f'coord_src_{i}'] = rs[:, i]
    if value_is_height:
        vs = skel.pixel_values[src]
        out[f'coord_src_{d}'] = vs
        rs = np.concatenate([rs, vs[:, None]], axis=1)
    rd = skel.coordinates[dst] * skel.spacing
    for i in range(d):
        out[f'coord_dst_{i}'] = rd[:, i]
    if value_is_height:
        vd = skel.pixel_values[dst]
        out[f'coord_dst_{d}'] = vd
        rd = np.concatenate([rd, vd[:, None]], axis=1)
    out['euclidean_distance'] = np.sqrt((rd - rs)**2 @ np.ones(d + int(value_is_height)))
    return {k.replace('_', separator): v for k, v in out.items()}


def skeleton_and_summary(image, *, spacing=1, value_is_height=False):
    """The whole timed job of the CPU baseline: Skeleton(...) + summarize(...)."""
    sk = OracleSkeleton(image, spacing=spacing, value_is_height=value_is_height)
    return sk, summarize(sk, value_is_height=value_is_height)
