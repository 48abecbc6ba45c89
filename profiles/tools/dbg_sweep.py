import sys, numpy as np, torch
sys.path.insert(0, '.')  # run from the repo root
import __graft_entry__ as g; g.build()
from skan_b200 import _pipeline
dev = torch.device('cuda', 0)
SITES = {1: 'converter free wait', 2: 'converter free wait (end)', 3: 'converter full wait', 4: 'scanner ready wait', 5: 'look-back'}
def run(shape, dens=0.02):
    rng = np.random.default_rng(1)
    img = rng.random(shape) < dens
    ctx = _pipeline._Ctx(dev)
    t, code, shp, npdt = _pipeline._as_device_image(img, dev)
    V = int(np.prod(shape)); ntiles = max(1, -(-V // 16384)); nchunks = -(-V // 65536)
    store = torch.zeros(ntiles * 256 + 4, dtype=torch.int64, device=dev); bits = store[2:-2]
    pre = torch.zeros(ntiles * 64, dtype=torch.int32, device=dev)
    cap = int(img.sum()) + 5
    lin = torch.zeros(cap, dtype=torch.int64, device=dev); co = torch.zeros((cap, len(shape)), dtype=torch.int64, device=dev)
    sb = ctx.lib.call('skb_sweep_scratch_bytes', V)
    scratch = torch.zeros(sb, dtype=torch.uint8, device=dev)
    n = torch.zeros(4, dtype=torch.int32, device=dev)
    st = torch.tensor(list(shape), dtype=torch.int64)
    import time; t0 = time.time()
    ctx.call('skb_sweep_nodes', t, code, len(shape), st.data_ptr(), 0, bits, ntiles, pre, cap, lin, co, None, scratch, n)
    torch.cuda.synchronize()
    dt = time.time() - t0
    nn = n.cpu().numpy(); w = scratch.cpu().numpy().view(np.uint64)
    ok = nn[0] == img.sum() and np.array_equal(lin[:nn[0]].cpu().numpy(), np.flatnonzero(img.reshape(-1)))
    print(shape, 'chunks', nchunks, 'N', nn[0], 'expected', int(img.sum()), 'timeouts', nn[1], 'ok', ok, f'{dt:.3f}s', flush=True)
    if nn[1]:
        d = w[1 + nchunks:]
        for k in range(min(int(d[0]), 15)):
            e = d[1 + 4 * k: 5 + 4 * k]
            print('   ', SITES.get(int(e[0] >> 32)), 'block', int(e[0] & 0xffffffff), [int(x) for x in e[1:3]], hex(int(e[3])), flush=True)
        st_ = w[1:1 + nchunks]
        print('    statuses', [(int(x >> 32), int(x & 0xffffffff)) for x in st_[:40]], flush=True)
for shape in [(64, 64, 64), (70, 100, 130), (64, 64, 80), (64, 64, 128), (256, 256, 256), (1024, 1024, 64)]:
    run(shape)
