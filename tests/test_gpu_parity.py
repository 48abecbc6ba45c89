"""Parity of the CUDA path (through the C-ABI, on a real B200) with the golden vectors of the
unmodified reference and with the CPU oracle on seeded inputs.  Integer / index results must be
bit-exact, fp64 sums within 1e-12 relative (BASELINE.json north_star)."""
import warnings

import numpy as np
import pytest
import torch

from golden_util import case, case_names, check_skeleton_against
from parity_util import LONG_CHAIN_KINDS, compare_with_oracle, long_chain_image
from skan_b200.synth import walks, walks_device

import skan_b200
from skan_b200 import _pipeline
from oracle import skan_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['ldg', 'tma', 'fused'])
def pack_variant(request):
    old = _pipeline.default_pack_variant
    _pipeline.default_pack_variant = request.param
    yield request.param
    _pipeline.default_pack_variant = old


def test_cuda_library_is_the_one_loaded():
    from skan_b200 import _native
    lib = _native.library()
    assert lib.device_type == 'cuda' and lib.path.endswith('libskan_b200.so')


@pytest.mark.parametrize('name', case_names())
def test_golden(name, pack_variant):
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


@pytest.mark.parametrize('shape,nw,steps,spacing', [
    ((1024, 1024), 256, 256, 1),
    ((777, 1031), 128, 256, (1.5, 0.7)),          # ragged: nothing divides the tile or word size
    ((96, 128, 160), 96, 128, (0.5, 0.1, 0.1)),
    ((61, 67, 71), 64, 96, 1),
    ((100000,), 0, 0, 2.0),                        # 1-D image
])
def test_walks_vs_oracle(shape, nw, steps, spacing, pack_variant):
    if len(shape) == 1:
        rng = np.random.default_rng(3)
        img = rng.random(shape) < 0.6
    else:
        img = walks(shape, nw, steps, seed=sum(shape), isolated=20, rings=8)
    compare_with_oracle(img, spacing)


@pytest.mark.parametrize('kind', LONG_CHAIN_KINDS)
def test_long_chains_and_big_cycles_vs_oracle(kind):
    # chains and junction-free cycles far longer than the splitter spacing of the list ranking
    img = long_chain_image(kind, 513)
    compare_with_oracle(img, (1.0, 0.5))
    rng = np.random.default_rng(1)
    compare_with_oracle((img * (rng.random(img.shape) + 0.5)).astype(np.float32), 1)


@pytest.mark.parametrize('seed', range(4))
def test_dense_blobs_vs_oracle(seed, pack_variant):
    # adversarial junction clusters: one giant cluster, many equal-weight ties
    rng = np.random.default_rng(2000 + seed)
    shape = [(200, 300), (40, 40, 40), (512, 64), (24, 48, 96)][seed]
    spacing = [1, (0.5, 0.1, 0.1), (2.0, 1.0), 1][seed]
    img = rng.random(shape) < [0.5, 0.3, 0.42, 0.2][seed]
    compare_with_oracle(img, spacing)


@pytest.mark.parametrize('seed', range(4))
def test_dense_junction_ids_vs_oracle(seed):
    """The Boruvka arrays indexed by dense ids of the junction voxels (the runner's choice above 12 M nodes; forced here):
    the same image three times -- without hints, with hints but the junction voxels not counted yet, then with dense ids --
    and once more after a much sparser image, whose count of junction voxels is too small a capacity (the launch does
    nothing, the runner goes again with node-indexed arrays).  Every run is compared with the oracle."""
    lib = skan_b200._native.library()
    rng = np.random.default_rng(2100 + seed)
    shape = [(200, 300), (40, 40, 40), (512, 64), (24, 48, 96)][seed]
    spacing = [1, (0.5, 0.1, 0.1), (2.0, 1.0), 1][seed]
    img = rng.random(shape) < [0.5, 0.3, 0.42, 0.2][seed]
    sparse = img & (rng.random(shape) < 0.5)
    from skan_b200 import _pipeline
    slack, _pipeline.HINT_SLACK = _pipeline.HINT_SLACK, 1
    lib.call('skb_set_option', b'dense_mst_min_nodes', 0)
    try:
        _pipeline._arena_cache.clear()
        for k in range(3):
            compare_with_oracle(img, spacing)
        (cache,) = [c for c in _pipeline._arena_cache.values() if c.get('counts')][-1:]
        assert cache['counts']['junction_nodes'] > 0 and cache['counts']['retaken'] == 0
        for k in range(3):
            compare_with_oracle(sparse, spacing)
        compare_with_oracle(img, spacing)       # hints of the sparse image: too small, also for the junction voxels
        (cache,) = [c for c in _pipeline._arena_cache.values() if c.get('counts')][-1:]
        assert cache['counts']['retaken'] >= 1
    finally:
        lib.call('skb_set_option', b'dense_mst_min_nodes', 12000000)
        _pipeline.HINT_SLACK = slack


@pytest.mark.parametrize('dt', [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64,
                                np.float32, np.float64])
def test_valued_images_vs_oracle(dt, pack_variant):
    rng = np.random.default_rng(5)
    w = walks((300, 421), 64, 200, seed=3, isolated=10, rings=5)
    vals = rng.integers(1, 100, w.shape) if np.issubdtype(dt, np.integer) else rng.random(w.shape) + 0.5
    compare_with_oracle((w * vals).astype(dt), (1.0, 1.5))


def test_height_mode_rejects_unsigned_images():
    # the reference wraps unsigned height differences; we refuse instead
    with pytest.raises(ValueError):
        skan_b200.Skeleton(np.eye(8, dtype=np.uint8) * 7, value_is_height=True)


@pytest.mark.parametrize('dt', [np.float32, np.float64, np.int8, np.int16, np.int32, np.int64])
def test_height_mode_vs_oracle(dt):
    rng = np.random.default_rng(6)
    w = walks((200, 200), 48, 128, seed=4, isolated=5, rings=3)
    vals = rng.integers(1, 100, w.shape) if np.issubdtype(dt, np.integer) else rng.random(w.shape) * 4 + 1
    compare_with_oracle((w * vals).astype(dt), (1.0, 2.0), value_is_height=True)


def test_float_zero_semantics(pack_variant):
    img = np.zeros((6, 40), np.float32)
    img[2, 1:37] = 1.0
    img[2, 3] = np.nan
    img[4, 4] = -0.0
    sk = skan_b200.Skeleton(img)
    assert sk.graph.shape[0] == 36 and sk.n_paths == 1


def test_inputs_cpu_tensor_cuda_tensor_numpy_agree(pack_variant):
    img = walks((333, 257), 40, 150, seed=8, isolated=5, rings=2)
    a = skan_b200.Skeleton(img)
    b = skan_b200.Skeleton(torch.from_numpy(img).pin_memory())
    c = skan_b200.Skeleton(torch.from_numpy(img).cuda())
    big = torch.zeros(img.size + 3, dtype=torch.bool, device='cuda')
    big[3:] = torch.from_numpy(img).cuda().reshape(-1)
    d = skan_b200.Skeleton(big[3:].reshape(img.shape))      # misaligned device pointer
    for o in (b, c, d):
        assert np.array_equal(a.graph.indices, o.graph.indices)
        assert np.array_equal(a.paths.indices, o.paths.indices)
        assert np.array_equal(a.coordinates, o.coordinates)


def test_zero_path_inputs_raise():
    for img in (np.zeros((4, 4), bool), np.eye(1, dtype=bool), np.array([[1, 0, 0, 1]], bool)):
        with pytest.raises(ValueError):
            skan_b200.Skeleton(img)


def test_pathgraph_equivalence():
    img = walks((256, 256), 64, 128, seed=9, isolated=3, rings=2)
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


def _structural_checks(sk, df, n_expected):
    """Size-independent properties of a correct result (used where the oracle is too slow)."""
    g = sk.graph
    n = g.shape[0]
    assert n == n_expected
    assert (g != g.T).nnz == 0                                  # symmetric
    assert g.has_sorted_indices
    deg = np.diff(g.indptr)
    assert np.array_equal(deg, sk.degrees)
    p = sk.paths
    src = p.indices[p.indptr[:-1]]
    dst = p.indices[p.indptr[1:] - 1]
    cyc = src == dst
    assert np.all((deg[src] != 2) | cyc) and np.all((deg[dst] != 2) | cyc)
    # interior path nodes have degree 2 and appear on exactly one path
    interior = np.ones(p.indices.size, bool)
    interior[p.indptr[:-1]] = False
    interior[p.indptr[1:] - 1] = False
    assert np.all(deg[p.indices[interior]] == 2)
    counts = np.bincount(p.indices[interior], minlength=n)
    assert counts.max() <= 1
    # every edge of the graph lies on exactly one path: sum over paths of (len - 1) == E / 2
    assert int(np.sum(np.diff(p.indptr) - 1)) == g.nnz // 2
    # consecutive path nodes are adjacent; branch_distance is the sum of those weights
    a, b = p.indices[:-1], p.indices[1:]
    inside = np.ones(a.size, bool)
    inside[p.indptr[1:-1] - 1] = False
    w = np.asarray(g[a[inside], b[inside]]).ravel()
    assert np.all(w > 0)
    seg = np.repeat(np.arange(p.shape[0]), np.diff(p.indptr) - 1)
    np.testing.assert_allclose(np.bincount(seg, weights=w, minlength=p.shape[0]), df['branch_distance'].to_numpy(),
                               rtol=1e-12)
    # skeleton ids agree with SciPy's component labels of the final graph
    from scipy.sparse import csgraph
    _, lab = csgraph.connected_components(g, directed=False)
    assert np.array_equal(lab[src], df['skeleton_id'].to_numpy())


@pytest.mark.slow
def test_config2_size_properties():
    """BASELINE configs[1]: 16384 x 16384 float32-valued skeleton; structural properties at full size."""
    shape = (16384, 16384)
    mask = walks_device(shape, 8192, 512, seed=2, device='cuda', isolated=1000, rings=64)
    gen = torch.Generator(device='cuda').manual_seed(2)
    img = (torch.rand(shape, device='cuda', generator=gen) * 0.9 + 0.1) * mask
    n_expected = int(mask.sum().item())
    del mask
    sk = skan_b200.Skeleton(img, keep_images=False)
    df = skan_b200.summarize(sk, separator='_')
    _structural_checks(sk, df, n_expected)
    pv = sk.pixel_values
    assert pv.dtype == np.float32 and np.all(pv > 0)
    means = df['mean_pixel_value'].to_numpy()
    assert np.all((means >= 0.1 - 1e-6) & (means <= 1.0 + 1e-6))


@pytest.mark.slow
def test_config4_size_properties():
    """BASELINE configs[3]: 1024^3 boolean volume, spacing (0.5, 0.1, 0.1)."""
    shape = (1024, 1024, 1024)
    img = walks_device(shape, 2048, 512, seed=4, device='cuda', isolated=1000, rings=64)
    n_expected = int(img.sum().item())
    sk = skan_b200.Skeleton(img, spacing=(0.5, 0.1, 0.1), keep_images=False)
    df = skan_b200.summarize(sk, separator='_')
    _structural_checks(sk, df, n_expected)


@pytest.mark.slow
def test_more_than_2_pow_32_voxels():
    """Raveled indices beyond 2^32: a small skeleton stamped at the far corner of a 4.4 Gvox volume
    must give the graph of the same skeleton in a small volume, translated."""
    small = walks((40, 48, 56), 16, 64, seed=21, isolated=3, rings=2)
    shape = (2100, 1500, 1400)
    off = (shape[0] - 41, shape[1] - 50, shape[2] - 57)
    big = torch.zeros(shape, dtype=torch.bool, device='cuda')
    big[off[0]:off[0] + 40, off[1]:off[1] + 48, off[2]:off[2] + 56] = torch.from_numpy(small).cuda()
    a = skan_b200.Skeleton(small, spacing=(0.5, 0.1, 0.1))
    b = skan_b200.Skeleton(big, spacing=(0.5, 0.1, 0.1), keep_images=False)
    assert np.array_equal(a.graph.indptr, b.graph.indptr)
    assert np.array_equal(a.graph.indices, b.graph.indices)
    assert np.array_equal(a.graph.data, b.graph.data)
    assert np.array_equal(a.coordinates + np.array(off), b.coordinates)
    assert np.array_equal(a.paths.indices, b.paths.indices)
    assert np.array_equal(a.path_lengths(), b.path_lengths())


@pytest.mark.parametrize('shape,valued', [((700, 900), False), ((80, 96, 112), False), ((512, 512), True)])
def test_native_runner_equals_stagewise_sequencing(shape, valued):
    """skb_run_image (one native call) and the per-stage Python sequencing run the same kernels in the
    same order: every output must be bit-identical."""
    img = walks(shape, 96, 200, seed=17, isolated=9, rings=5)
    if valued:
        img = (img * (np.random.default_rng(1).random(shape) + 0.5)).astype(np.float32)
    t = torch.from_numpy(img).cuda()
    a = _pipeline.skeleton_and_summary_device(t, spacing=1.5, device=t.device)
    b = _pipeline.run_native(t, spacing=1.5, device=t.device)
    pairs = [(a.graph.indptr, b.graph.indptr), (a.graph.indices, b.graph.indices), (a.graph.data, b.graph.data),
             (a.graph.coords, b.graph.coords), (a.paths.indptr, b.paths.indptr), (a.paths.indices, b.paths.indices),
             (a.paths.data, b.paths.data), (a.paths.skeleton_id, b.paths.skeleton_id), (a.dist, b.dist),
             (a.mean, b.mean), (a.stdev, b.stdev), (a.table[0], b.table[0]), (a.table[1], b.table[1]),
             (a.table[2], b.table[2])]
    for i, (x, y) in enumerate(pairs):
        assert torch.equal(x.view(torch.uint8) if x.dtype.is_floating_point else x,
                           y.view(torch.uint8) if y.dtype.is_floating_point else y), i


def test_image_on_a_device_that_is_not_current():
    """Skeleton(image on cuda:k) while another device is current: the library must run on the image's device (per-device
    SM count, shared-memory opt-in, events and mailbox)."""
    n = torch.cuda.device_count()
    img = walks((40, 48, 56), 24, 96, seed=7, isolated=5, rings=3)
    ref = orc.OracleSkeleton(img, spacing=(0.5, 0.1, 0.1))
    prev = torch.cuda.current_device()
    try:
        for k in range(min(n, 3) - 1, -1, -1):           # the highest device first, device 0 current
            torch.cuda.set_device(0 if k else n - 1)
            sk = skan_b200.Skeleton(torch.from_numpy(img).to(f'cuda:{k}'), spacing=(0.5, 0.1, 0.1))
            assert np.array_equal(sk.graph.indices, ref.graph.indices)
            assert np.array_equal(sk.paths.indices, ref.paths.indices)
            df = skan_b200.summarize(sk, separator='_')
            assert len(df) == ref.n_paths
    finally:
        torch.cuda.set_device(prev)


@pytest.mark.parametrize('small', [0, 1])
def test_junction_free_cycles_small_and_general_path(small):
    """Up to 4096 uncovered chain voxels are handled by one thread block in shared memory (skb_small_cycles_kernel),
    more by the general list-ranking machinery; both must reproduce the reference's rings, order and entries."""
    from skan_b200 import _native
    lib = _native.library()
    rng = np.random.default_rng(3)
    try:
        lib.call('skb_set_option', b'small_cycles', small)
        for kind, n in (('ring', 65), ('ring', 257), ('nested_rings', 97), ('nested_rings', 385), ('ring_with_tail', 129)):
            compare_with_oracle(long_chain_image(kind, n), 1)
        img = walks((300, 400), 60, 200, seed=4, isolated=10, rings=120)
        compare_with_oracle(img, (2.0, 0.5))
        compare_with_oracle((img * (rng.random(img.shape) + 0.5)).astype(np.float32), 1)
        only_rings = np.zeros((64, 64), bool)          # no regular path at all: rings and isolated pixels
        for c in range(6, 60, 12):
            only_rings[c - 1, c] = only_rings[c, c - 1] = only_rings[c, c + 1] = only_rings[c + 1, c] = True
        only_rings[2, 60] = True
        compare_with_oracle(only_rings, 1)
        vol = walks((48, 64, 80), 30, 100, seed=8, isolated=5, rings=40)
        compare_with_oracle(vol, (0.5, 0.1, 0.1))
    finally:
        lib.call('skb_set_option', b'small_cycles', 1)


def test_one_very_long_path():
    """A single chain of ~10^6 voxels (a serpentine).  Default: statistics of a path longer than 32768 entries are summed by
    one thread block in a fixed tree order, the label image by 32 lanes per path; integer results stay bit-exact, the fp64
    sums agree with the reference's left-to-right / pairwise sums to the rounding error bound of those sums themselves
    (n * 2^-53 for the left-to-right sum of n terms: 1.2e-10 here; measured 1.05e-11).  With
    skb_set_option('long_path_exact', 1) every path is summed by one thread in the reference's order: bit-exact."""
    import time
    n = 1449
    img = np.zeros((n, n), bool)      # a serpentine whose turns go through one diagonal pixel: no junctions, ONE chain
    for k, r in enumerate(range(1, n - 1, 2)):
        img[r, 2:n - 2] = True
        if r + 2 < n - 1:
            img[r + 1, n - 2 if k % 2 == 0 else 1] = True
    rng = np.random.default_rng(8)
    valued = (img * (rng.random(img.shape) + 0.5)).astype(np.float64)
    ref = orc.OracleSkeleton(valued, spacing=(1.0, 0.7))
    rs = orc.summarize(ref)
    assert ref.n_paths == 1 and ref.paths.indices.size > 1_000_000
    bound = ref.paths.indices.size * 2.0 ** -53
    dev = torch.from_numpy(valued).cuda()
    skan_b200.Skeleton(dev, spacing=(1.0, 0.7))                      # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sk = skan_b200.Skeleton(dev, spacing=(1.0, 0.7))
    df = skan_b200.summarize(sk, separator='_')
    dt = time.perf_counter() - t0
    assert np.array_equal(sk.paths.indices, ref.paths.indices)
    for k in ('branch_distance', 'mean_pixel_value', 'euclidean_distance'):
        np.testing.assert_allclose(df[k].to_numpy(), rs[k], rtol=bound, err_msg=k)
    np.testing.assert_allclose(df['stdev_pixel_value'].to_numpy() ** 2, rs['stdev_pixel_value'] ** 2, rtol=1e-9)
    assert np.array_equal(sk.path_label_image(), ref.path_label_image())
    assert dt < 0.5, f'{dt:.3f} s for one long path'
    lib = skan_b200._native.library()
    lib.call('skb_set_option', b'long_path_exact', 1)
    try:
        sk = skan_b200.Skeleton(dev, spacing=(1.0, 0.7))
        df = skan_b200.summarize(sk, separator='_')
        for k in ('branch_distance', 'mean_pixel_value', 'stdev_pixel_value', 'euclidean_distance'):
            assert np.array_equal(df[k].to_numpy(), rs[k]), k
    finally:
        lib.call('skb_set_option', b'long_path_exact', 0)
