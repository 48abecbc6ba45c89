"""The fused sweep (skb_sweep_nodes: mask + scan + nodes in one pass) against the three-kernel path
(skb_pack_count -> skb_scan_u32 -> skb_emit_nodes) and against NumPy, called through the C-ABI."""
import numpy as np
import pytest
import torch

from skan_b200 import _native, _pipeline

pytestmark = pytest.mark.gpu

DTYPES = [np.bool_, np.uint8, np.int16, np.float32, np.float64, np.int64]


def fused(img, node_cap, z_origin=0):
    dev = torch.device('cuda', 0)
    ctx = _pipeline._Ctx(dev)
    t, code, shape, np_dtype = _pipeline._as_device_image(img, dev)
    V = int(np.prod(shape))
    tile = ctx.lib.tile_voxels
    ntiles = max(1, -(-V // tile))
    store = torch.zeros(ntiles * 256 + 4, dtype=torch.int64, device=dev)
    bits = store[2:-2]
    pre256 = torch.full((ntiles * 64,), -1, dtype=torch.int32, device=dev)
    ndim = len(shape)
    lin = torch.full((node_cap,), -1, dtype=torch.int64, device=dev)
    coords = torch.full((node_cap, ndim), -1, dtype=torch.int64, device=dev)
    values = None if np_dtype == np.dtype(bool) else torch.full((node_cap,), -1.0, dtype=torch.float64, device=dev)
    scratch = torch.empty(ctx.lib.call('skb_sweep_scratch_bytes', V), dtype=torch.uint8, device=dev)
    n = torch.zeros(4, dtype=torch.int32, device=dev)
    shape_t = torch.tensor(list(shape), dtype=torch.int64)
    ctx.call('skb_sweep_nodes', t, code, ndim, shape_t.data_ptr(), z_origin, bits, ntiles, pre256, node_cap, lin,
             coords, values, scratch, n)
    torch.cuda.synchronize()
    return dict(bits=bits.cpu().numpy().view(np.uint64), pre256=pre256.cpu().numpy(), lin=lin.cpu().numpy(),
                coords=coords.cpu().numpy(), values=None if values is None else values.cpu().numpy(),
                n=int(n[0].item()), ntiles=ntiles)


def expected(img):
    flat = np.asarray(img).reshape(-1)
    mask = flat.astype(bool)
    lin = np.flatnonzero(mask)
    coords = np.stack(np.unravel_index(lin, img.shape), axis=1)
    V = flat.size
    ntiles = max(1, -(-V // 16384))
    padded = np.zeros(ntiles * 16384, bool)
    padded[:V] = mask
    bits = np.packbits(padded, bitorder='little').view(np.uint64)
    blocks = padded.reshape(-1, 256).sum(1)
    pre256 = np.concatenate([[0], np.cumsum(blocks)[:-1]])
    return lin, coords, bits, pre256, flat[lin].astype(np.float64)


@pytest.mark.parametrize('shape', [(5,), (3, 3), (97, 131), (300, 700), (17, 33, 65), (64, 64, 64), (70, 100, 130)])
@pytest.mark.parametrize('dt', DTYPES)
def test_fused_sweep_matches_numpy(shape, dt):
    rng = np.random.default_rng(hash((shape, np.dtype(dt).num)) % (1 << 31))
    mask = rng.random(shape) < (0.3 if np.prod(shape) < 100 else 0.02)
    img = mask if dt is np.bool_ else (mask * rng.integers(1, 100, shape)).astype(dt)
    lin, coords, bits, pre256, vals = expected(img)
    r = fused(img, node_cap=len(lin) + 7)
    assert r['n'] == len(lin)
    assert np.array_equal(r['bits'], bits)
    assert np.array_equal(r['pre256'], pre256)
    assert np.array_equal(r['lin'][:len(lin)], lin)
    assert np.array_equal(r['coords'][:len(lin)], coords)
    assert (r['lin'][len(lin):] == -1).all()
    if r['values'] is not None:
        assert np.array_equal(r['values'][:len(lin)], vals)


def test_fused_sweep_dense_and_empty():
    for img in (np.ones((130, 517), np.uint8), np.zeros((130, 517), bool), np.ones((1,), bool)):
        lin, coords, bits, pre256, _ = expected(img)
        r = fused(img, node_cap=len(lin) + 1)
        assert r['n'] == len(lin)
        assert np.array_equal(r['bits'], bits) and np.array_equal(r['pre256'], pre256)
        assert np.array_equal(r['lin'][:len(lin)], lin) and np.array_equal(r['coords'][:len(lin)], coords)


def test_fused_sweep_capacity_is_respected():
    rng = np.random.default_rng(3)
    img = rng.random((200, 300, 40)) < 0.01
    lin, coords, bits, pre256, _ = expected(img)
    cap = len(lin) // 3
    r = fused(img, node_cap=cap, z_origin=11)
    assert r['n'] == len(lin) and cap < len(lin)
    assert np.array_equal(r['lin'], lin[:cap])
    c = coords[:cap].copy()
    c[:, 0] += 11
    assert np.array_equal(r['coords'], c)
    assert np.array_equal(r['bits'], bits) and np.array_equal(r['pre256'], pre256)


def test_native_runner_recovers_from_a_small_capacity_guess():
    """skb_run_image with a node capacity hint below N falls back to the two-pass emission."""
    import skan_b200
    from skan_b200.synth import walks
    from oracle import skan_oracle as orc
    img = walks((96, 128, 80), 64, 128, seed=9, isolated=20, rings=4)
    dev = torch.device('cuda', 0)
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    _pipeline.run_native(torch.from_numpy(img).to(dev), spacing=1, device=dev)          # fills the cache
    _pipeline._arena_cache[key]['nodes'] = 1                                              # a hint far below N
    _pipeline._arena_cache[key]['shape'] = img.shape
    r = _pipeline.run_native(torch.from_numpy(img).to(dev), spacing=1, device=dev)
    ref = orc.OracleSkeleton(img, spacing=1)
    assert np.array_equal(r.graph.indptr.cpu().numpy(), ref.graph.indptr)
    assert np.array_equal(r.graph.indices.cpu().numpy(), ref.graph.indices)
    assert np.array_equal(r.graph.coords.cpu().numpy(), ref.coordinates)
    assert np.array_equal(r.paths.indices.cpu().numpy(), ref.paths.indices)


def test_fused_sweep_large_volume_counts():
    """> 2**32 voxels would not fit the test box's time budget here; 1024 x 1024 x 600 checks many chunks per CTA,
    a ragged last chunk, and the look-back chain across ~10k chunks."""
    from skan_b200.synth import walks_device
    dev = torch.device('cuda', 0)
    shape = (601, 1023, 1025)
    img = walks_device(shape, 512, 256, 12, dev, isolated=100, rings=8)
    lin = torch.nonzero(img.reshape(-1)).reshape(-1)
    r = fused(img, node_cap=int(lin.numel()) + 3)
    assert r['n'] == lin.numel()
    assert np.array_equal(r['lin'][:r['n']], lin.cpu().numpy())
    c = np.stack(np.unravel_index(lin.cpu().numpy(), shape), axis=1)
    assert np.array_equal(r['coords'][:r['n']], c)
