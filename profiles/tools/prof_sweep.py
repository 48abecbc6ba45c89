"""In-kernel cycle counters of the fused sweep (library built with -DSKB_SWEEP_PROFILE=<cta>)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')  # run from the repo root
from skan_b200 import _native, _pipeline
lib = _native.Library(sys.argv[1], device_type='cuda')
_native._install_library_for_tests(lib)
dev = torch.device('cuda', 0)
NAMES = ['conv full_wait', 'conv free_wait', 'scan ready_wait', 'scan lookback', 'lookback polls', 'lookback rounds',
         'scan pass1', 'scan pass2', 'conv chunks', 'conv bar', 'conv issue', 'scan chunks']
def run(shape, dens):
    V = int(np.prod(shape))
    img = (torch.rand(shape, device=dev) < dens)
    ctx = _pipeline._Ctx(dev)
    ntiles = -(-V // 16384); nchunks = -(-V // 65536)
    store = torch.zeros(ntiles * 256 + 4, dtype=torch.int64, device=dev); bits = store[2:-2]
    pre = torch.zeros(ntiles * 64, dtype=torch.int32, device=dev)
    cap = int(img.sum().item()) + 5
    lin = torch.zeros(cap, dtype=torch.int64, device=dev); co = torch.zeros((cap, len(shape)), dtype=torch.int64, device=dev)
    scratch = torch.zeros(ctx.lib.call('skb_sweep_scratch_bytes', V), dtype=torch.uint8, device=dev)
    n = torch.zeros(4, dtype=torch.int32, device=dev)
    st = torch.tensor(list(shape), dtype=torch.int64)
    for it in range(3):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.call('skb_sweep_nodes', img, 0, len(shape), st.data_ptr(), 0, bits, ntiles, pre, cap, lin, co, None, scratch, n)
        b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    w = scratch.cpu().numpy().view(np.uint64)[1 + nchunks + 64:]
    print(shape, 'chunks', nchunks, 'N', int(n[0]), 'timeouts', int(n[1]), f'{ms:.3f} ms', f'{V / ms / 1e6:.1f} GB/s', flush=True)
    for i, nm in enumerate(NAMES):
        print(f'   {nm:18s} {int(w[i]):>14d}' + (f'  = {int(w[i]) / 1.9e3:10.1f} us' if 'chunks' not in nm and 'polls' not in nm and 'rounds' not in nm else ''), flush=True)
for shape, dens in [((256, 1024, 1024), 1e-3), ((1024, 1024, 1024), 1e-3)]:
    run(shape, dens)
