"""image_stats.mesh_sizes / image_summary (reference src/skan/image_stats.py) against golden vectors of the unmodified
reference (tests/golden/make_image_stats.py), through the host emulation harness and on the GPU.  Integer results
(mesh sizes, junction and skeleton counts) bit-exact; the fp64 statistics of the size list are computed by the same
NumPy calls on the same integers, so they are compared exactly too."""
import os
import warnings

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'image_stats.npz'))
NAMES = sorted({k.split('/')[0] for k in GOLD.files})


def load(name):
    img = GOLD[name + '/image']
    sp = GOLD[name + '/spacing']
    spacing = sp[0].item() if bool(GOLD[name + '/spacing_scalar']) else tuple(x.item() for x in sp)
    return img, spacing


def check(name, to_device):
    from skan_b200 import image_stats
    img, spacing = load(name)
    sizes = image_stats.mesh_sizes(to_device(img))
    want = GOLD[name + '/sizes']
    assert sizes.dtype == want.dtype and np.array_equal(sizes, want)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')     # np.mean of an empty list, like the reference
        df = image_stats.image_summary(to_device(img), spacing=spacing)
    assert list(df.columns) == [str(c) for c in GOLD[name + '/columns']]
    for c in df.columns:
        if c == 'scale':
            continue
        got, ref = np.asarray(df[c]), GOLD[name + '/col/' + c]
        assert got.dtype == ref.dtype, (c, got.dtype, ref.dtype)
        assert np.array_equal(got, ref, equal_nan=True), (c, got, ref)


@pytest.mark.parametrize('name', NAMES)
def test_image_stats_emulated(name):
    from emul_util import emulated_device
    with emulated_device():
        check(name, lambda a: a)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_image_stats_gpu(name):
    check(name, lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda())


def test_mesh_sizes_rejects_other_dimensions_emulated():
    from emul_util import emulated_device
    from skan_b200 import image_stats
    with emulated_device():
        with pytest.raises(ValueError):
            image_stats.mesh_sizes(np.zeros((4, 4, 4), bool))


@pytest.mark.gpu
def test_mesh_sizes_large_gpu_vs_scipy():
    # a 4096^2 skeleton: run-based labelling on the device against scipy.ndimage.label on the host copy
    from scipy import ndimage as ndi
    from skan_b200 import image_stats
    from skan_b200.synth import walks
    img = walks((4096, 4096), 3000, 400, seed=31, isolated=200, rings=64)
    sizes = image_stats.mesh_sizes(torch.from_numpy(img).cuda())
    labeled = ndi.label(~img)[0]
    touching = np.unique(np.concatenate((labeled[0], labeled[-1], labeled[:, 0], labeled[:, -1])))
    want = np.bincount(labeled.flat)
    want[touching] = 0
    want = want[want != 0]
    assert np.array_equal(sizes, want)


@pytest.mark.parametrize('seed', range(12))
def test_mesh_sizes_random_vs_scipy_emulated(seed):
    # random masks of every density and ragged widths (rows that straddle 64-bit words in every phase) against
    # scipy.ndimage.label, with the reference's border / label-0 handling (image_stats.py:31-45)
    from scipy import ndimage as ndi
    from emul_util import emulated_device
    from skan_b200 import image_stats
    rng = np.random.default_rng(100 + seed)
    ny, nx = int(rng.integers(1, 90)), int(rng.integers(1, 200))
    img = rng.random((ny, nx)) < rng.choice([0.05, 0.2, 0.4, 0.6, 0.9])
    if seed % 3 == 0:
        img[1:-1, 1:-1] |= False
        img[0, :] = img[-1, :] = False      # skeleton clear of the top and bottom rows
    with emulated_device():
        got = image_stats.mesh_sizes(img)
    labeled = ndi.label(~img)[0]
    touching = np.unique(np.concatenate((labeled[0], labeled[-1], labeled[:, 0], labeled[:, -1])))
    want = np.bincount(labeled.flat)
    want[touching] = 0
    want = want[want != 0]
    assert np.array_equal(got, want), (ny, nx)
