"""Golden vectors for the public pixel_graph (reference csr.py:51-158): the UNMODIFIED reference on boolean masks at
every connectivity, complete image graphs with the default edge function, and a mask with a custom edge function.

    python tests/golden/make_pixel_graph.py      ->  tests/golden/pixel_graph.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)


def edge_fn(v0, v1, d):
    return (v0 + 2.0 * v1) / d


def cases():
    from skan_b200.synth import walks
    rng = np.random.default_rng(6)
    blob2 = rng.random((17, 23)) < 0.45
    blob3 = rng.random((7, 9, 11)) < 0.35
    for c in (1, 2):
        yield f'blob2_c{c}', blob2, None, None, c, None
        yield f'blob2_c{c}_sp', blob2, None, None, c, (2.0, 0.5)
    for c in (1, 2, 3):
        yield f'blob3_c{c}', blob3, None, None, c, (0.5, 0.1, 0.1)
    yield 'line1d', rng.random(40) < 0.6, None, None, 1, None
    yield 'walks2', walks((60, 70), 10, 40, seed=2, rings=2), None, None, 2, None
    vals = rng.integers(0, 50, (6, 7)).astype(np.int64)
    yield 'complete_int', vals, None, None, 1, None                  # default edge function on the full lattice
    yield 'complete_float_c2', rng.random((5, 6)), None, None, 2, (1.0, 3.0)
    yield 'masked_custom', rng.random((12, 14)), rng.random((12, 14)) < 0.5, 'custom', 2, (1.5, 1.0)
    yield 'masked_plain', rng.random((12, 14)).astype(np.float32), rng.random((12, 14)) < 0.5, None, 1, None


def main():
    import _refshim
    csr = _refshim.import_reference()
    store = {}
    for name, image, mask, fn, conn, spacing in cases():
        g, nodes = csr.pixel_graph(image, mask=mask, edge_function=edge_fn if fn else None, connectivity=conn,
                                   spacing=spacing)
        g.sort_indices()
        store[name + '/image'] = image
        if mask is not None:
            store[name + '/mask'] = mask
        store[name + '/custom'] = np.array(bool(fn))
        store[name + '/connectivity'] = np.array(conn)
        if spacing is not None:
            store[name + '/spacing'] = np.asarray(spacing)
        store[name + '/indptr'] = g.indptr
        store[name + '/indices'] = g.indices
        store[name + '/data'] = g.data
        store[name + '/nodes'] = nodes
        print(name, g.shape, g.nnz, g.data.dtype, g.indices.dtype, nodes.dtype)
    np.savez_compressed(os.path.join(HERE, 'pixel_graph.npz'), **store)


if __name__ == '__main__':
    main()
