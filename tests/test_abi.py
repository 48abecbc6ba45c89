"""The C-ABI library loads and exports every symbol include/skan_b200.h declares (no compute:
this runs without a GPU), and the python binding table agrees with the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'skan_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(skb_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def built_library():
    import __graft_entry__ as g
    g.build()
    return g.LIB


def test_header_and_binding_table_agree():
    from skan_b200 import _native
    assert declared_symbols() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol(built_library):
    dll = ctypes.CDLL(built_library)
    for sym in declared_symbols():
        assert hasattr(dll, sym), sym
    dll.skb_abi_version.restype = ctypes.c_int
    assert dll.skb_abi_version() == 1
    dll.skb_tile_voxels.restype = ctypes.c_int64
    assert dll.skb_tile_voxels() == 16384


def test_library_is_sm100a_and_uses_tma(built_library):
    """The cubin in the library targets sm_100a and the TMA variant of the sweep is a bulk copy."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '-lelf', built_library], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    sass = subprocess.run([cuobjdump, '-sass', built_library], capture_output=True, text=True).stdout
    functions = {part.split()[0]: part for part in sass.split('Function : ')[1:]}
    for prefix in ('_Z19skb_pack_tma_kernelILi1E', '_Z22skb_sweep_nodes_kernelILi1E'):
        names = [n for n in functions if n.startswith(prefix)]
        assert names, f'{prefix}* not found in the library'
        assert all('UBLKCP' in functions[n] for n in names), f'TMA bulk copy (cp.async.bulk) not found in {names}'


def test_package_has_no_cpu_fallback():
    """Without a CUDA device the public API must raise instead of computing somewhere else."""
    import numpy as np
    import torch
    import skan_b200
    from skan_b200 import _native
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    with pytest.raises(_native.NativeError):
        skan_b200.Skeleton(np.eye(5, dtype=bool))


def test_numa_placement_never_raises():
    """skan_b200/_numa.py is an optimisation of the end-to-end leg: without NVML / a GPU it reports that nothing was bound."""
    import os
    from skan_b200._numa import bind_thread_to_gpu
    before = os.sched_getaffinity(0)
    info = bind_thread_to_gpu(0)
    assert set(info) >= {'bound', 'cpus', 'numa_node'}
    if not info['bound']:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)
