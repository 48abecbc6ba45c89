"""Host-side logic and the per-item device functions, run through the host emulation harness
(tests/emul: the same skb_items.cuh / skb_api.inl compiled with g++, kernels as serial loops)
against the golden vectors of the unmodified reference and against the oracle.  This is what
can be checked without a GPU; the CUDA path itself is checked by tests/test_gpu_parity.py."""
import warnings

import numpy as np
import pytest

from emul_util import emulated_device
from golden_util import case, case_names, check_skeleton_against
from oracle import skan_oracle as orc
from parity_util import LONG_CHAIN_KINDS, compare_with_oracle, long_chain_image
from skan_b200.synth import walks

import skan_b200


@pytest.fixture(scope='module', autouse=True)
def _emulated():
    with emulated_device():
        yield


@pytest.mark.parametrize('name', case_names())
def test_golden_through_emulation(name):
    rec = case(name)
    if 'error' in rec:
        with pytest.raises(ValueError):
            skan_b200.Skeleton(rec['image'], spacing=rec['spacing_arg'], value_is_height=rec['height'])
        return
    sk = skan_b200.Skeleton(rec['image'], spacing=rec['spacing_arg'], value_is_height=rec['height'])
    df = skan_b200.summarize(sk, value_is_height=rec['height'], separator='_')
    summ = {c: df[c].to_numpy() for c in df.columns}
    check_skeleton_against(rec, sk, summ, exact_weights=True)
    for c, t in zip(rec['summary_columns'], rec['summary_dtypes']):
        assert str(df[str(c)].dtype) == str(t), (c, t)
    assert np.array_equal(sk.path_label_image(), rec['path_label_image'])
    if 'pixel_values' in rec:
        assert sk.pixel_values.dtype == rec['pixel_values'].dtype
        assert np.array_equal(sk.pixel_values, rec['pixel_values'])
    else:
        assert sk.pixel_values is None


@pytest.mark.parametrize('seed', range(6))
def test_random_dense_blobs_vs_oracle(seed):
    # dense random images: large junction clusters with many equal-weight ties
    rng = np.random.default_rng(1000 + seed)
    shape = [(23, 31), (9, 10, 11), (40, 17), (6, 6, 30), (64, 64), (12, 12, 12)][seed]
    spacing = [1, (0.5, 0.1, 0.1), (2.0, 1.0), 1, (1, 1), (1.5, 1.5, 0.7)][seed]
    img = rng.random(shape) < [0.5, 0.3, 0.45, 0.35, 0.4, 0.25][seed]
    compare_with_oracle(img, spacing)


@pytest.mark.parametrize('kind', LONG_CHAIN_KINDS)
def test_long_chains_and_big_cycles_vs_oracle(kind):
    # chains and junction-free cycles far longer than the splitter spacing of the list ranking
    img = long_chain_image(kind, 129)
    sk, df, ref = compare_with_oracle(img, (1.0, 0.5))
    if kind in ('ring', 'nested_rings'):
        assert np.all(df['branch_type'].to_numpy() == 3)
    rng = np.random.default_rng(1)
    compare_with_oracle((img * (rng.random(img.shape) + 0.5)).astype(np.float64), 1)


def test_values_and_dtypes_vs_oracle():
    rng = np.random.default_rng(5)
    w = walks((70, 90), 16, 80, seed=3, isolated=4, rings=2)
    for dt in (np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64, np.float32,
               np.float64):
        vals = rng.integers(1, 100, w.shape) if np.issubdtype(dt, np.integer) else rng.random(w.shape) + 0.5
        compare_with_oracle((w * vals).astype(dt), (1.0, 1.5))


@pytest.mark.parametrize('dt', [np.float32, np.float64, np.int8, np.int16, np.int32, np.int64])
def test_height_mode_vs_oracle(dt):
    rng = np.random.default_rng(6)
    w = walks((90, 90), 24, 64, seed=4, isolated=5, rings=3)
    vals = rng.integers(1, 100, w.shape) if np.issubdtype(dt, np.integer) else rng.random(w.shape) * 4 + 1
    compare_with_oracle((w * vals).astype(dt), (1.0, 2.0), value_is_height=True)
    with pytest.raises(ValueError):
        skan_b200.Skeleton((w * 3).astype(np.uint16), value_is_height=True)


def test_negative_zero_and_nan_are_handled_like_astype_bool():
    img = np.zeros((6, 8), np.float32)
    img[2, 1:7] = 1.0
    img[2, 3] = np.nan      # nonzero -> foreground
    img[4, 4] = -0.0        # zero -> background
    sk = skan_b200.Skeleton(img)
    assert sk.graph.shape[0] == 6 and sk.n_paths == 1


def test_pathgraph_equivalence():
    # test_csr.py:387-417: Skeleton == PathGraph.from_image == PathGraph.from_graph
    img = walks((80, 80), 20, 60, seed=9, isolated=3, rings=2)
    sk = skan_b200.Skeleton(img, spacing=2.5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        a = skan_b200.summarize(sk, separator='_')
        b = skan_b200.summarize(skan_b200.PathGraph.from_image(img, spacing=2.5), separator='_')
        c = skan_b200.summarize(skan_b200.PathGraph.from_graph(adj=sk.graph, node_coordinates=sk.coordinates,
                                                               spacing=2.5), separator='_')
    for col in a.columns:
        assert np.array_equal(a[col].to_numpy(), b[col].to_numpy()), col
        assert np.array_equal(a[col].to_numpy(), c[col].to_numpy()), col


def test_skeleton_from_path_graph():
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


def test_separator_warning_and_accessors():
    img = case('skeleton1')['image']
    sk = skan_b200.Skeleton(img)
    with pytest.warns(np.exceptions.VisibleDeprecationWarning):
        df = skan_b200.summarize(sk)
    assert 'branch-distance' in df.columns                      # test_csr.py:374-378
    assert [list(map(int, p)) for p in sk.paths_list()] == [
        [7, 5, 0, 1, 2, 3, 4, 6, 10, 9, 12], [7, 8, 12], [7, 11, 13], [12, 14, 15, 16]]
    assert sk.path_coordinates(3).tolist() == [[3, 3], [4, 4], [4, 5], [4, 6]]   # test_skeleton_class.py:41-44
    idx, data = sk.path_with_data(1)
    assert idx.tolist() == [7, 8, 12] and data.tolist() == [1.0, 1.0, 1.0]
    assert tuple(sk.paths.shape) == (4, 21)
    assert sk.degrees.dtype == np.int32 and sk.coordinates.dtype == np.int64
    assert sk.graph.indices.dtype == np.int32 and sk.graph.data.dtype == np.float64
    with pytest.raises(ValueError):
        sk.prune_paths([99])


def test_skeleton_to_csgraph_contract():
    g, coords = skan_b200.skeleton_to_csgraph(case('tinycycle')['image'])     # test_csr.py:53-63
    assert g.indptr.tolist() == [0, 2, 4, 6, 8]
    assert g.indices.tolist() == [1, 2, 0, 3, 0, 3, 1, 2]
    assert np.all(g.data == np.sqrt(2))
    assert np.ravel_multi_index(coords, (3, 3)).tolist() == [1, 3, 5, 7]


def test_prune_paths_like_the_reference(monkeypatch):
    """Skeleton.prune_paths (csr.py:818-854): wipe the non-junction pixels of the paths, re-skeletonize on the host with
    scikit-image, build a new Skeleton.  scikit-image is not in this image: without it the call raises
    NotImplementedError after the reference's index check; with a stand-in whose skeletonize is the identity (true for
    an image that is still one pixel wide after the wipe) the result is the skeleton of the wiped image."""
    import sys
    import types
    img = case('tinycycle')['image'].copy().astype(bool)
    img = np.pad(img, 3)
    img[3, 6:9] = True                                    # a spur hanging off the ring
    with emulated_device():
        sk = skan_b200.Skeleton(img)
        with pytest.raises(ValueError):
            sk.prune_paths([sk.n_paths])
        monkeypatch.setitem(sys.modules, 'skimage', None)  # import skimage -> ImportError
        with pytest.raises(NotImplementedError):
            sk.prune_paths([0])
        fake = types.ModuleType('skimage')
        fake.morphology = types.ModuleType('skimage.morphology')
        fake.morphology.skeletonize = lambda a: np.asarray(a, bool)
        monkeypatch.setitem(sys.modules, 'skimage', fake)
        monkeypatch.setitem(sys.modules, 'skimage.morphology', fake.morphology)
        spur = [i for i in range(sk.n_paths) if np.any(sk.degrees[sk.path(i)] == 1)]
        assert len(spur) == 1
        pruned = sk.prune_paths(spur)
        wiped = img.copy()
        ids = sk.path(spur[0])
        keep = sk.degrees[ids] > 2
        wiped[tuple(sk.coordinates[ids[~keep]].T)] = False
        ref = orc.OracleSkeleton(wiped)
        assert np.array_equal(pruned.graph.indices, ref.graph.indices)
        assert np.array_equal(pruned.paths.indices, ref.paths.indices)
        assert pruned.n_paths == ref.n_paths
