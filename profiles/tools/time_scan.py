import sys, torch
sys.path.insert(0, '.')  # run from the repo root
from skan_b200 import _native, _pipeline
dev = torch.device('cuda', 0)
ctx = _pipeline._Ctx(dev)
for n in (174154, 1000000, 7513720, 30000000):
    x = torch.randint(0, 3, (n + 1,), dtype=torch.int32, device=dev)
    ref = torch.cumsum(x[:n].long(), 0)
    out = torch.empty(n + 1, dtype=torch.int32, device=dev)
    sc = ctx.scratch(n)
    ctx.call('skb_scan_u32', x, out, n, sc)
    torch.cuda.synchronize()
    assert int(out[0]) == 0 and torch.equal(out[1:].long(), ref), n
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ctx.call('skb_scan_u32', x, out, n, sc)
    b.record(); torch.cuda.synchronize()
    print(f'scan n={n}: {a.elapsed_time(b) / 20 * 1e3:.1f} us', flush=True)
