import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from guard_util import guarded
from skan_b200 import _native, _pipeline
from test_gpu_guards import _cases
dev = torch.device('cuda', 0)
_pipeline.default_pack_variant = 'fused'
lib = _native.library()
for name, img, spacing, kw in _cases():
    for fill in (0x00, 0xA5):
        with guarded(fill) as g:
            for rep in range(2):
                r = _pipeline.run_native(torch.from_numpy(img).to(dev), spacing=spacing, device=dev, **kw)
                torch.cuda.synchronize()
                for which, arena in enumerate(_pipeline.LAST_ARENAS):
                    offs = np.zeros(512, dtype=np.int64)
                    n = lib._dll.skb_debug_guard_offsets(which, offs.ctypes.data, 512)
                    for gi, o in enumerate(offs[:n]):
                        gap = arena[int(o): int(o) + 256].cpu().numpy()
                        bad = np.flatnonzero(gap != fill)
                        if bad.size:
                            print(name, 'fill', fill, 'rep', rep, 'arena', which, 'gap#', gi, 'off', int(o), 'bad bytes', bad[:8], gap[bad[:16]], 'N', r.graph.n_nodes, 'offs', offs[:min(n,6)])
print('done')
