"""TEST INFRASTRUCTURE: build tests/emul/skb_emul.cpp with g++ and route skan_b200 through it,
so that the host-side logic and the per-item device functions can be exercised on a machine
without a GPU.  Never used by the package itself (skan_b200/_native.py only loads the CUDA
library); GPU tests (-m gpu) never call this."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'emul', 'skb_emul.cpp')
OUT = os.path.join(HERE, 'emul', '_build', 'libskb_emul.so')
DEPS = [SRC] + [os.path.join(ROOT, 'skan_b200', 'csrc', f) for f in ('skb_items.cuh', 'skb_sholl.cuh', 'skb_mesh.cuh', 'skb_api.inl', 'skb_platform.h', 'skb_run.inl', 'skb_slab.inl')]


def build_emulation_library():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-w', '-shared', '-fPIC',
                               '-I', os.path.join(ROOT, 'skan_b200', 'csrc'), '-o', OUT, SRC])
    return OUT


class emulated_device:
    """Context manager: skan_b200 runs on the host emulation harness inside the block."""

    def __enter__(self):
        from skan_b200 import _native
        self._native = _native
        lib = _native.Library(build_emulation_library(), device_type='cpu')
        self._old = _native._install_library_for_tests(lib)
        return lib

    def __exit__(self, *exc):
        self._native._install_library_for_tests(self._old)
        return False
