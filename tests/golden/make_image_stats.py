"""Golden vectors for image_stats.mesh_sizes / image_summary: the UNMODIFIED reference (src/skan/image_stats.py,
SciPy's ndimage.label) on its doctest images, its fixtures and seeded skeletons, in the build container:

    python tests/golden/make_image_stats.py      ->  tests/golden/image_stats.npz
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)


def cases(fix):
    from skan_b200.synth import walks
    doc = np.array([[0, 0, 1, 0, 0],
                    [0, 0, 1, 1, 1],
                    [0, 0, 1, 0, 0],
                    [0, 1, 0, 1, 0]])                      # image_stats.py:25-28
    yield 'doctest', doc, 1
    yield 'doctest_padded', np.pad(doc, 1, mode='constant', constant_values=1), 1     # image_stats.py:32-34: [7 2 3 1]
    yield 'skeleton0', np.asarray(fix['skeleton0']), 1
    yield 'skeleton1', np.asarray(fix['skeleton1']), 2.5
    yield 'skeleton2', np.asarray(fix['skeleton2']), (1.5, 0.7)
    yield 'tinycycle', np.asarray(fix['tinycycle']), 1
    yield 'empty', np.zeros((7, 9), bool), 1
    yield 'full', np.ones((5, 6), bool), 1
    yield 'frame', np.pad(np.zeros((6, 70), bool), 1, mode='constant', constant_values=True), 1
    grid = np.zeros((67, 131), bool)
    grid[::6, :] = True
    grid[:, ::10] = True
    yield 'grid', grid, (0.5, 2.0)                          # closed meshes of two sizes, skeleton on the border
    inner = np.zeros((80, 200), bool)
    inner[10:70:12, 20:180] = True
    inner[10:59, 20:180:16] = True
    yield 'grid_inside', inner, 1                           # skeleton clear of the border: sizes[0] = pixel count
    yield 'walks_a', walks((160, 200), 60, 150, seed=3, isolated=5, rings=6), 1
    yield 'walks_b', walks((257, 129), 90, 200, seed=8, isolated=9, rings=4), (2.0, 0.5)
    yield 'walks_wide', walks((33, 1000), 40, 300, seed=12, isolated=3, rings=2), 1
    rng = np.random.default_rng(4)
    yield 'noise', rng.random((90, 77)) < 0.45, 1           # many tiny spaces, ragged row length
    yield 'valued', (walks((120, 120), 50, 100, seed=2, rings=3) * rng.random((120, 120))).astype(np.float32), 1


def main():
    import _refshim
    _refshim.import_reference()
    stats = importlib.import_module('skan.image_stats')
    fix = _refshim.reference_fixtures()
    store = {}
    for name, img, spacing in cases(fix):
        sizes = stats.mesh_sizes(img)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')          # mean of an empty array
            df = stats.image_summary(img, spacing=spacing)
        store[name + '/image'] = img
        store[name + '/spacing'] = np.atleast_1d(np.asarray(spacing))
        store[name + '/spacing_scalar'] = np.array(np.isscalar(spacing))
        store[name + '/sizes'] = sizes
        cols = [c for c in df.columns if c != 'scale']
        store[name + '/columns'] = np.array(list(df.columns))
        for c in cols:
            store[name + '/col/' + c] = np.asarray(df[c])
        print(name, sizes[:10], {c: df[c].iloc[0] for c in cols if 'number' in c})
    np.savez_compressed(os.path.join(HERE, 'image_stats.npz'), **store)


if __name__ == '__main__':
    main()
