"""Minimal stand-in for the reference's numba ``NBGraph`` view (csr.py:161-216): the same
fields and the two methods callers use (``edge``, ``neighbors``), over host arrays.  The hot
path itself never touches it -- it exists so that code reading ``Skeleton.nbgraph`` keeps working."""
import numpy as np


class NBGraph:
    def __init__(self, indptr, indices, data, shape, node_props):
        self.indptr = indptr
        self.indices = indices
        self.data = data
        self.shape = shape
        self.node_properties = node_props

    def edge(self, i, j):
        lo, hi = self.indptr[i], self.indptr[i + 1]
        for k in range(lo, hi):
            if self.indices[k] == j:
                return self.data[k]
        return 0.0

    def neighbors(self, row):
        return self.indices[self.indptr[row]:self.indptr[row + 1]]

    @property
    def has_node_props(self):
        return self.node_properties.strides != (0,)


def csr_to_nbgraph(csr, node_props=None):
    if node_props is None:
        node_props = np.broadcast_to(1., csr.shape[0])
        node_props.flags.writeable = True
    return NBGraph(csr.indptr.astype(np.int32, copy=False), csr.indices.astype(np.int32, copy=False),
                   csr.data, np.array(csr.shape, dtype=np.int32), node_props.astype(np.float64))
