"""Golden vectors for summarize(find_main_branch=True): the UNMODIFIED reference's find_main_branches
(summary_utils.py:32-62, networkx) applied to branch tables of seeded skeletons.  Run in the build container:

    python tests/golden/make_main_branch.py      ->  tests/golden/main_branch.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)


def cases():
    from skan_b200.synth import walks
    yield 'walks2d', walks((96, 128), 12, 80, seed=3, isolated=4, rings=2), 1
    yield 'walks2d_b', walks((160, 160), 30, 60, seed=11, isolated=0, rings=3), (2.0, 0.5)
    yield 'walks3d', walks((40, 48, 56), 16, 70, seed=5, isolated=3, rings=2), (0.5, 0.1, 0.1)
    blob = np.random.default_rng(2).random((40, 40)) < 0.3
    yield 'dense2d', blob, 1


def main():
    import _refshim
    csr = _refshim.import_reference()
    import warnings
    store = {}
    for name, img, spacing in cases():
        sk = csr.Skeleton(img, spacing=spacing)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            df = csr.summarize(sk, find_main_branch=True, separator='_')
        store[name + '/image'] = np.packbits(img.reshape(-1))
        store[name + '/shape'] = np.array(img.shape)
        store[name + '/spacing'] = np.atleast_1d(np.asarray(spacing, dtype=np.float64))
        store[name + '/main'] = df['main'].to_numpy()
        print(name, len(df), int(df['main'].sum()))
    np.savez_compressed(os.path.join(HERE, 'main_branch.npz'), **store)


if __name__ == '__main__':
    main()
