"""Comparison of skan_b200 results with the CPU oracle on the same input (bit-exact for
integer/index arrays, rtol 1e-12 for fp64 sums: BASELINE.json north_star)."""
import numpy as np

from golden_util import assert_close, assert_same_int
from oracle import skan_oracle as orc

import skan_b200


def compare_with_oracle(img, spacing=1, value_is_height=False, image_for_device=None, exact=True):
    ref = orc.OracleSkeleton(img, spacing=spacing, value_is_height=value_is_height)
    rs = orc.summarize(ref, value_is_height=value_is_height)
    sk = skan_b200.Skeleton(img if image_for_device is None else image_for_device, spacing=spacing,
                            value_is_height=value_is_height)
    df = skan_b200.summarize(sk, value_is_height=value_is_height, separator='_')
    assert_same_int(sk.graph.indptr, ref.graph.indptr, 'graph.indptr')
    assert_same_int(sk.graph.indices, ref.graph.indices, 'graph.indices')
    if exact:
        assert np.array_equal(sk.graph.data, ref.graph.data), 'graph.data'
    else:
        assert_close(sk.graph.data, ref.graph.data, 'graph.data')
    assert_same_int(sk.coordinates, ref.coordinates, 'coordinates')
    assert_same_int(sk.degrees, ref.degrees, 'degrees')
    assert_same_int(sk.paths.indptr, ref.paths.indptr, 'paths.indptr')
    assert_same_int(sk.paths.indices, ref.paths.indices, 'paths.indices')
    assert_close(sk.paths.data, ref.paths.data, 'paths.data')
    assert tuple(sk.paths.shape) == tuple(ref.paths.shape)
    assert list(df.columns) == list(rs.keys())
    m = np.abs(rs['mean_pixel_value'])
    for c in df.columns:
        got, want = df[c].to_numpy(), np.asarray(rs[c])
        if want.dtype.kind in 'iub':
            assert_same_int(got, want, c)
        elif c.startswith('stdev'):
            assert np.all(np.abs(got**2 - want**2) <= 1e-12 * np.maximum(m * m, 1e-300) + 1e-12 * want**2), c
        else:
            assert_close(got, want, c)
    return sk, df, ref


def long_chain_image(kind, n=257):
    """Images whose chains are much longer than the 1/32 splitter spacing of the sampled list
    ranking: a serpentine line, a big ring (junction-free cycle with many splitters), nested
    rings, a ring with a tail (one junction), and a comb."""
    img = np.zeros((n, n), bool)
    if kind == 'serpentine':
        for r in range(1, n - 1, 2):
            img[r, 1:n - 1] = True
            col = n - 2 if (r // 2) % 2 == 0 else 1
            if r + 1 < n - 1:
                img[r + 1, col] = True
    elif kind in ('ring', 'nested_rings'):
        # diamond rings |y-c| + |x-c| == R: every pixel has exactly two (diagonal) neighbours
        c = n // 2
        yy, xx = np.mgrid[0:n, 0:n]
        dist = np.abs(yy - c) + np.abs(xx - c)
        radii = [c - 2] if kind == 'ring' else list(range(3, c - 1, 3))
        for R in radii:
            img |= dist == R
    elif kind == 'ring_with_tail':
        img[10, 10:n - 10] = True
        img[n - 11, 10:n - 10] = True
        img[10:n - 10, 10] = True
        img[10:n - 10, n - 11] = True
        img[n // 2, 0:10] = True
    elif kind == 'comb':
        img[n // 2, :] = True
        for c in range(3, n - 3, 4):
            img[5:n // 2, c] = True
    else:
        raise ValueError(kind)
    return img


LONG_CHAIN_KINDS = ['serpentine', 'ring', 'nested_rings', 'ring_with_tail', 'comb']
