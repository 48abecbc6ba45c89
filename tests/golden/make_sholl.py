"""Golden vectors for sholl_analysis (reference csr.py:1332-1390): the UNMODIFIED reference applied to its own
fixtures and to seeded skeletons, in the build container:

    python tests/golden/make_sholl.py      ->  tests/golden/sholl.npz

The reference takes the geodesic centre from skimage.graph.central_pixel, which is not installed here; for the
cases with center=None the shim's stub is replaced by the restatement in skan_b200/_sholl.py (the reference's own
known answers for it are asserted in tests/test_sholl.py).  Everything else -- _simplify_graph, the shell
normalisation, distance_matrix, digitize, bincount -- is the reference's code and SciPy / NumPy.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)


def cases(fix):
    from skan_b200.synth import walks
    # (name, image, spacing, center, shells)
    yield 'skeleton0_default', fix['skeleton0'], 1, None, None
    yield 'skeleton0_ref_test', fix['skeleton0'], 1, None, np.arange(0, 5, 1.5)          # test_csr.py:276-280
    yield 'skeleton0_spacing', fix['skeleton0'], (1, 5), [3, 15], np.arange(1, 20, 6)    # test_csr.py:291-294
    yield 'skeleton0_spacing17', fix['skeleton0'], (1, 5), [3, 15], np.arange(17)        # test_csr.py:286-290
    yield 'skeleton4_diag', fix['skeleton4'], 1, [1, 1], np.arange(0.09, 5, 1.45)        # test_csr.py:302-305
    yield 'skeleton3d_int', fix['skeleton3d'], 1, None, 4
    yield 'skeleton2_float', fix['skeleton2'], (1.5, 0.7), None, None
    w2 = walks((160, 200), 30, 120, seed=3, isolated=5, rings=3)
    yield 'walks2d_none', w2, (2.0, 0.5), None, None
    yield 'walks2d_int', w2, 1, [80, 100], 12
    yield 'walks2d_offnode', w2, (0.25, 0.75), [17.3, 80.9], np.linspace(0.0, 140.0, 29)
    yield 'walks2d_decreasing', w2, 1, [80, 100], np.arange(120, 0, -7.5)
    w3 = walks((40, 48, 56), 24, 90, seed=5, isolated=3, rings=2)
    yield 'walks3d_aniso', w3, (0.5, 0.1, 0.1), None, None
    yield 'walks3d_int', w3, (2, 1, 1), [40, 24, 28], 9
    yield 'walks3d_list', w3, (0.5, 0.1, 0.1), [10.0, 2.4, 2.8], [0.5, 1.0, 2.0, 4.0, 8.0, 16.0]


def main():
    import _refshim
    csr = _refshim.import_reference()
    from skan_b200._sholl import central_pixel
    csr.central_pixel = central_pixel            # the absent skimage function (see the module docstring)
    fix = _refshim.reference_fixtures()
    store = {}
    for name, img, spacing, center, shells in cases(fix):
        img = np.asarray(img)
        sk = csr.Skeleton(img, spacing=spacing)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter('always')
            c, r, counts = csr.sholl_analysis(sk, center=center, shells=shells)
        warned = any('Sholl' in str(x.message) for x in w)
        store[name + '/image'] = np.packbits(img.astype(bool).reshape(-1))
        store[name + '/shape'] = np.array(img.shape)
        store[name + '/spacing'] = np.atleast_1d(np.asarray(spacing))
        store[name + '/spacing_scalar'] = np.array(np.isscalar(spacing))
        store[name + '/center_arg'] = np.array([]) if center is None else np.asarray(center)
        store[name + '/center_none'] = np.array(center is None)
        if shells is None:
            store[name + '/shells_kind'] = np.array(0)
        elif np.isscalar(shells):
            store[name + '/shells_kind'] = np.array(1)
            store[name + '/shells_arg'] = np.array(shells)
        else:
            store[name + '/shells_kind'] = np.array(2)
            store[name + '/shells_arg'] = np.asarray(shells)
        store[name + '/center'] = np.asarray(c)
        store[name + '/radii'] = np.asarray(r)
        store[name + '/counts'] = np.asarray(counts)
        store[name + '/warned'] = np.array(warned)
        store[name + '/center_idx'] = np.array(csr._fast_graph_center_idx(sk))
        print(name, np.asarray(c), len(r), counts.tolist()[:12], 'warned' if warned else '')
    np.savez_compressed(os.path.join(HERE, 'sholl.npz'), **store)


if __name__ == '__main__':
    main()
