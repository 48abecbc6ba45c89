"""The public pixel_graph (reference csr.py:51-158) against golden vectors of the unmodified reference
(tests/golden/make_pixel_graph.py): masks at every connectivity, complete image graphs with the default edge
function, a custom edge function.  Structure bit-exact; distances are the same table entries, so the data are
compared exactly as well."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'pixel_graph.npz'))
NAMES = sorted({k.split('/')[0] for k in GOLD.files})


def edge_fn(v0, v1, d):
    return (v0 + 2.0 * v1) / d


def check(name):
    from skan_b200 import csr
    g = lambda k: GOLD[name + '/' + k]
    has = lambda k: (name + '/' + k) in GOLD.files
    spacing = tuple(float(x) for x in g('spacing')) if has('spacing') else None
    graph, nodes = csr.pixel_graph(g('image'), mask=g('mask') if has('mask') else None,
                                   edge_function=edge_fn if bool(g('custom')) else None,
                                   connectivity=int(g('connectivity')), spacing=spacing)
    assert graph.indptr.dtype == g('indptr').dtype and np.array_equal(graph.indptr, g('indptr'))
    assert graph.indices.dtype == g('indices').dtype and np.array_equal(graph.indices, g('indices'))
    assert graph.data.dtype == g('data').dtype and np.array_equal(graph.data, g('data'))
    assert nodes.dtype == g('nodes').dtype and np.array_equal(nodes, g('nodes'))


@pytest.mark.parametrize('name', NAMES)
def test_pixel_graph_emulated(name):
    from emul_util import emulated_device
    with emulated_device():
        check(name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_pixel_graph_gpu(name):
    check(name)


def _degree_cases():
    from skan_b200.synth import walks
    rng = np.random.default_rng(8)
    yield walks((50, 60), 10, 40, seed=1, isolated=3, rings=2)
    yield rng.random((9, 8, 7)) < 0.4
    yield (rng.random((21, 19)) < 0.5) * rng.random((21, 19))        # valued image: nonzero pixels are the skeleton
    yield rng.random(30) < 0.5


def _check_degree_image(to_device):
    # the reference's formula (csr.py:1197-1205) with SciPy, on the host copy
    from scipy import ndimage as ndi
    from skan_b200 import csr
    for img in _degree_cases():
        b = img.astype(bool)
        kernel = np.ones((3,) * b.ndim)
        kernel[(1,) * b.ndim] = 0
        want = ndi.convolve(b.astype(int), kernel, mode='constant') * b
        got = csr.make_degree_image(to_device(img))
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want)


def test_make_degree_image_emulated():
    from emul_util import emulated_device
    with emulated_device():
        _check_degree_image(lambda a: a)


@pytest.mark.gpu
def test_make_degree_image_gpu():
    import torch
    _check_degree_image(lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda())


def test_small_utilities():
    from scipy import sparse
    from skan_b200 import csr
    M = sparse.csr_matrix(np.arange(16).reshape((4, 4)))
    assert csr.submatrix(M, [0, 2]).toarray().tolist() == [[0, 2], [8, 10]]        # csr.py:1163-1167
    assert np.array_equal(csr.distance_with_height(np.array([3.0]), np.array([0.0]), np.array([4.0])), [5.0])
    nbg = csr.csr_to_nbgraph(M.astype(float))
    assert nbg.edge(1, 2) == 6.0 and nbg.neighbors(0).tolist() == [1, 2, 3]
    with pytest.raises(AttributeError):                                              # test_csr.py:590-600
        csr.csr_to_nbgraph(np.zeros((3, 3)))


@pytest.mark.gpu
def test_skeleton_from_path_graph_gpu():
    import warnings
    import skan_b200
    from skan_b200.synth import walks
    # csr.py:670-688: a Skeleton built from a PathGraph has the same paths, statistics and label image
    rng = np.random.default_rng(2)
    img = walks((64, 72), 16, 50, seed=4, isolated=2, rings=2)
    valued = (img * (rng.random(img.shape) + 0.5)).astype(np.float64)
    sk = skan_b200.Skeleton(valued, spacing=(1.5, 0.5))
    pg = skan_b200.PathGraph.from_graph(adj=sk.graph, node_coordinates=sk.coordinates, node_values=sk.pixel_values,
                                        spacing=(1.5, 0.5))
    sk2 = skan_b200.Skeleton.from_path_graph(pg)
    assert sk2.n_paths == sk.n_paths and sk2.paths_list() == sk.paths_list()
    assert np.array_equal(sk2.path_lengths(), sk.path_lengths())
    assert np.array_equal(sk2.path_means(), sk.path_means())
    assert np.array_equal(sk2.path_stdev(), sk.path_stdev())
    assert np.array_equal(sk2.degrees, sk.degrees) and np.array_equal(sk2.coordinates, sk.coordinates)
    assert tuple(sk2.skeleton_shape) == tuple(np.max(sk.coordinates, axis=0) + 1)
    lab, lab2 = sk.path_label_image(), sk2.path_label_image()
    assert np.array_equal(lab[:lab2.shape[0], :lab2.shape[1]], lab2)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        a, b = skan_b200.summarize(sk, separator='_'), skan_b200.summarize(sk2, separator='_')
    for col in a.columns:
        assert np.array_equal(a[col].to_numpy(), b[col].to_numpy()), col


def _both_ways_body(to_device):
    """skb_run_image can write every chain from both ends ("emit_both_ways"); the result must be the same bit for bit
    as with the one-ended walk, for regular paths, self-loops, junction-free cycles and valued images."""
    import numpy as np
    from parity_util import LONG_CHAIN_KINDS, compare_with_oracle, long_chain_image
    from skan_b200 import _native
    from skan_b200.synth import walks
    lib = _native.library()
    rng = np.random.default_rng(12)
    images = [(long_chain_image(k, 193), 1) for k in LONG_CHAIN_KINDS]
    images.append((walks((120, 150), 40, 120, seed=5, isolated=6, rings=5), (2.0, 0.5)))
    images.append((walks((30, 36, 40), 24, 80, seed=6, isolated=4, rings=3), (0.5, 0.1, 0.1)))
    images.append(((walks((90, 90), 30, 90, seed=7, rings=3) * (rng.random((90, 90)) + 0.5)).astype(np.float32), 1))
    try:
        for mode in (0, 1):
            lib.call('skb_set_option', b'emit_both_ways', mode)
            for img, spacing in images:
                compare_with_oracle(img, spacing, image_for_device=to_device(img))
    finally:
        lib.call('skb_set_option', b'emit_both_ways', 1)


def test_emit_both_ways_emulated():
    from emul_util import emulated_device
    with emulated_device():
        _both_ways_body(lambda a: a)


@pytest.mark.gpu
def test_emit_both_ways_gpu():
    import torch
    _both_ways_body(lambda a: torch.from_numpy(a).cuda())


def _splitter_bits_body(to_device):
    """One chain voxel in 2^bits splits the chains into walks ("splitter_bits"); any value gives the same result."""
    from parity_util import compare_with_oracle, long_chain_image
    from skan_b200 import _native
    from skan_b200.synth import walks
    lib = _native.library()
    images = [(long_chain_image(k, 161), 1) for k in ('serpentine', 'nested_rings', 'ring_with_tail')]
    images.append((walks((40, 44, 48), 30, 90, seed=3, isolated=4, rings=3), (0.5, 0.1, 0.1)))
    try:
        for bits in (1, 3, 8, 12):
            lib.call('skb_set_option', b'splitter_bits', bits)
            for both in (0, 1):
                lib.call('skb_set_option', b'emit_both_ways', both)
                for img, spacing in images:
                    compare_with_oracle(img, spacing, image_for_device=to_device(img))
    finally:
        lib.call('skb_set_option', b'splitter_bits', 0)
        lib.call('skb_set_option', b'emit_both_ways', 1)


def test_splitter_bits_emulated():
    from emul_util import emulated_device
    with emulated_device():
        _splitter_bits_body(lambda a: a)


@pytest.mark.gpu
def test_splitter_bits_gpu():
    import torch
    _splitter_bits_body(lambda a: torch.from_numpy(a).cuda())
