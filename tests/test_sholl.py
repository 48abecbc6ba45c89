"""sholl_analysis (reference csr.py:1332-1390) against golden vectors produced by the unmodified reference
(tests/golden/make_sholl.py), through the host emulation harness and on the GPU; plus the reference's own known
answers for the geodesic centre (test_csr.py:265-305), which pin the restated skimage.graph.central_pixel."""
import os
import warnings

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'sholl.npz'))
NAMES = sorted({k.split('/')[0] for k in GOLD.files})


def load(name):
    g = lambda k: GOLD[name + '/' + k]
    shape = tuple(int(x) for x in g('shape'))
    img = np.unpackbits(g('image'))[:int(np.prod(shape))].astype(bool).reshape(shape)
    sp = g('spacing')
    spacing = sp[0].item() if bool(g('spacing_scalar')) else tuple(x.item() for x in sp)
    center = None if bool(g('center_none')) else g('center_arg')
    kind = int(g('shells_kind'))
    shells = None if kind == 0 else (g('shells_arg').item() if kind == 1 else g('shells_arg'))
    return img, spacing, center, shells


def check(name, to_device):
    import skan_b200
    from skan_b200.csr import sholl_analysis, _fast_graph_center_idx
    img, spacing, center, shells = load(name)
    sk = skan_b200.Skeleton(to_device(img), spacing=spacing)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        c, r, counts = sholl_analysis(sk, center=center, shells=shells)
    assert any('Sholl' in str(x.message) for x in w) == bool(GOLD[name + '/warned'])
    want_c, want_r, want_counts = GOLD[name + '/center'], GOLD[name + '/radii'], GOLD[name + '/counts']
    assert np.asarray(c).dtype == want_c.dtype and np.array_equal(c, want_c)
    assert np.asarray(r).dtype == want_r.dtype and np.array_equal(r, want_r)      # bit-exact: same host arithmetic,
    assert np.array_equal(counts, want_counts)                                     # device max distance included
    assert counts.dtype == want_counts.dtype
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        assert _fast_graph_center_idx(sk) == int(GOLD[name + '/center_idx'])


@pytest.mark.parametrize('name', NAMES)
def test_sholl_emulated(name):
    from emul_util import emulated_device
    with emulated_device():
        check(name, lambda a: a)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_sholl_gpu(name):
    check(name, lambda a: torch.from_numpy(a).cuda())


def test_reference_known_answers_emulated():
    # test_csr.py:265-305 of the reference, verbatim expectations
    import skan_b200
    from skan_b200 import csr
    from emul_util import emulated_device
    from golden_util import case
    skeleton0 = case('skeleton0')['image']
    skeleton4 = case('skeleton4')['image']
    with emulated_device():
        s = skan_b200.Skeleton(skeleton0)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            assert csr._fast_graph_center_idx(s) == 6
            assert csr._fast_graph_center_idx(skan_b200.Skeleton(skeleton4)) == 1
            c, r, counts = csr.sholl_analysis(s, shells=np.arange(0, 5, 1.5))
        np.testing.assert_equal(c, [3, 3])
        np.testing.assert_equal(counts, [0, 3, 3, 0])
        s = skan_b200.Skeleton(skeleton0, spacing=(1, 5))
        with pytest.warns(UserWarning):
            c, r, counts = csr.sholl_analysis(s, center=[3, 15], shells=np.arange(17))
            for i in range(4):
                assert np.isin(i, counts)
        c, r, counts = csr.sholl_analysis(s, center=[3, 15], shells=np.arange(1, 20, 6))
        np.testing.assert_equal(counts, [3, 2, 2, 0])
        s = skan_b200.Skeleton(skeleton4)
        c, r, counts = csr.sholl_analysis(s, center=[1, 1], shells=np.arange(0.09, 5, 1.45))
        np.testing.assert_equal(counts, [3, 2, 1, 0])


@pytest.mark.gpu
def test_sholl_large_gpu_vs_numpy():
    # a volume-scale skeleton: the device histogram against the reference's NumPy formula on the host copies
    import skan_b200
    from skan_b200.csr import sholl_analysis
    from skan_b200.synth import walks_device
    img = walks_device((256, 256, 256), 512, 300, 9, torch.device('cuda'), isolated=50, rings=8)
    sk = skan_b200.Skeleton(img, spacing=(0.5, 0.1, 0.1))
    center = np.array([60.0, 12.5, 13.0])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        c, r, counts = sholl_analysis(sk, center=center, shells=40)
    scaled = sk.coordinates * sk.spacing
    edges = sk.graph.tocoo()
    d = np.sqrt(np.sum(np.abs(center - scaled) ** 2.0, axis=-1))
    assert r[-1] <= d.max() + 1e-6 and np.isclose(r[1] - r[0], d.max() / 40)
    b = np.digitize(d, r)
    b0, b1 = b[edges.row], b[edges.col]
    cross = b0 != b1
    want = np.bincount(np.minimum(b0[cross], b1[cross]), minlength=len(r)) // 2
    assert np.array_equal(counts, want)


@pytest.mark.parametrize('seed', range(8))
def test_sholl_random_vs_numpy_emulated(seed):
    # random skeletons, centres on and off the lattice, integer / float / anisotropic spacing: the device histogram
    # against the reference's formula (csr.py:1372-1388) evaluated with SciPy / NumPy on the host copies
    from scipy.spatial import distance_matrix
    import skan_b200
    from emul_util import emulated_device
    from skan_b200.csr import sholl_analysis
    from skan_b200.synth import walks
    rng = np.random.default_rng(50 + seed)
    shape = [(60, 70), (20, 24, 28)][seed % 2]
    spacing = [1, (2, 3), (0.5, 1.25), (0.5, 0.1, 0.1), 2, (1.5, 1.0, 0.25)][seed % 6]
    if np.ndim(spacing) and len(spacing) != len(shape):
        spacing = tuple(spacing)[:len(shape)] if len(spacing) > len(shape) else tuple(spacing) + (1.0,)
    img = walks(shape, 12, 50, seed=seed, isolated=2, rings=2)
    with emulated_device():
        sk = skan_b200.Skeleton(img, spacing=spacing)
        scaled = sk.coordinates * sk.spacing
        center = scaled[rng.integers(0, len(scaled))] + (rng.random(len(shape)) if seed % 2 else 0)
        shells = [None, 7, np.linspace(0.0, 40.0, 9)][seed % 3]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            c, r, counts = sholl_analysis(sk, center=center, shells=shells)
            edges = sk.graph.tocoo()
            d0 = distance_matrix(scaled[edges.row], [c]).ravel()
            d1 = distance_matrix(scaled[edges.col], [c]).ravel()
    b0, b1 = np.digitize(d0, r), np.digitize(d1, r)
    cross = b0 != b1
    want = np.bincount(np.minimum(b0[cross], b1[cross]), minlength=len(r)) // 2
    assert np.array_equal(counts, want)
