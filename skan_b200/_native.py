"""ctypes binding of the C-ABI library ``libskan_b200.so`` (declared in include/skan_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) into
``skan_b200/lib/``.  There is no CPU implementation behind this module: if the library is
missing, or no CUDA device is visible, using the package raises.  torch tensors are the
only array interop -- every pointer argument is ``tensor.data_ptr()``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libskan_b200.so')

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int

# name -> (restype, argtypes); mirrors include/skan_b200.h
_SIGNATURES = {
    'skb_abi_version': (_int, []),
    'skb_last_error': (ctypes.c_char_p, []),
    'skb_scan_scratch_bytes': (_i64, [_i64]),
    'skb_tile_voxels': (_i64, []),
    'skb_kernel_launches': (_i64, []),
    'skb_debug_guard_offsets': (_i64, [_int, _vp, _i64]),
    'skb_set_option': (_int, [ctypes.c_char_p, _i64]),
    'skb_pack_count': (_int, [_vp, _int, _i64, _vp, _vp, _i64, _int, _vp]),
    'skb_emit_nodes': (_int, [_vp, _vp, _i64, _int, _vp, _vp, _int, _i64, _vp, _vp, _vp, _vp, _vp]),
    'skb_emit_nodes_cap': (_int, [_vp, _vp, _i64, _int, _vp, _vp, _int, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    'skb_sweep_scratch_bytes': (_i64, [_i64]),
    'skb_sweep_nodes': (_int, [_vp, _int, _int, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'skb_tile_prefix': (_int, [_vp, _vp, _i64, _vp, _vp]),
    'skb_raw_neighbors': (_int, [_vp, _int, _vp, _vp, _i64, _int, _vp, _vp]),
    'skb_mask_neighbors': (_int, [_vp, _i64, ctypes.c_uint32, _vp]),
    'skb_degree_image': (_int, [_vp, _vp, _i64, _vp, _vp]),
    'skb_junction_edge_count': (_int, [_vp, _vp, _int, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    'skb_junction_edge_emit': (_int, [_vp, _vp, _int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _int, _int,
                                      _vp, _vp, _vp, _vp, _vp]),
    'skb_mst_junctions': (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'skb_mst_apply': (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    'skb_degrees_indptr': (_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    'skb_csr_emit': (_int, [_vp, _vp, _int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _int,
                            _vp, _vp, _vp, _vp, _vp, _vp]),
    'skb_path_records': (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'skb_path_start_count': (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'skb_path_walk': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'skb_path_rank': (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    'skb_path_select': (_int, [_vp, _vp, _vp, _i64, _int, _i64, _vp, _vp, _vp]),
    'skb_path_heads': (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                              _vp]),
    'skb_path_emit': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    'skb_cycle_roots': (_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    'skb_skeleton_ids': (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'skb_path_stats': (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    'skb_branch_table': (_int, [_i64, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                _vp, _vp, _vp, _vp, _vp]),
    'skb_rank_query': (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    'skb_slab_covered': (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    'skb_slab_export': (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    'skb_slab_table_rank': (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'skb_slab_apply': (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp]),
    'skb_slab_head_records': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp,
                                     _vp, _vp]),
    'skb_copy_segments': (_int, [_vp, _vp]),
    'skb_slab_edge_shift': (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    'skb_slab_einfo': (_int, [_vp, _i64, _i64, _vp, _vp]),
    'skb_slab_components': (_int, [_vp, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp]),
    'skb_slab_skeleton_ids': (_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    'skb_slab_path_stats': (_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    'skb_path_label_image': (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    'skb_scan_u32': (_int, [_vp, _vp, _i64, _vp, _vp]),
    'skb_sholl_bins': (_int, [_vp, _i64, _int, _int, _vp, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _vp]),
    'skb_sholl_count': (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    'skb_mesh_runs': (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    'skb_mesh_link': (_int, [_vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'skb_mesh_emit': (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    'skb_mesh_border_foreground': (_int, [_vp, _i64, _i64, _vp, _vp]),
    'skb_cc_count': (_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
}



class RunParams(ctypes.Structure):
    """struct skb_run_params (include/skan_b200.h)"""
    _fields_ = [('image', _vp), ('dtype', ctypes.c_int32), ('ndim', ctypes.c_int32), ('shape', _i64 * 3),
                ('z_origin', _i64), ('image_stack', ctypes.c_int32), ('height', ctypes.c_int32),
                ('pack_variant', ctypes.c_int32), ('stop_after_graph', ctypes.c_int32),
                ('want_values', ctypes.c_int32), ('time_sweep', ctypes.c_int32), ('wtab', _vp),
                ('spacing_f', ctypes.c_double * 3), ('spacing_i', _i64 * 3), ('spacing_is_int', ctypes.c_int32),
                ('table_ndim', ctypes.c_int32), ('work', _vp), ('work_bytes', _i64), ('out', _vp),
                ('out_bytes', _i64), ('stream', _vp), ('node_cap_hint', _i64), ('hint', _i64 * 8)]


class RunResult(ctypes.Structure):
    """struct skb_run_result (include/skan_b200.h)"""
    _fields_ = [('counts', _i64 * 16), ('off', _i64 * 32), ('work_need', _i64), ('out_need', _i64),
                ('stage_ns', _i64 * 32)]


_SIGNATURES['skb_run_image'] = (_int, [ctypes.POINTER(RunParams), ctypes.POINTER(RunResult)])


class SlabParams(ctypes.Structure):
    """struct skb_slab_params (include/skan_b200.h)"""
    _fields_ = [('image', _vp), ('dtype', ctypes.c_int32), ('pack_variant', ctypes.c_int32), ('global_shape', _i64 * 3),
                ('z0', _i64), ('z1', _i64), ('ze0', _i64), ('nz_ext', _i64), ('want_values', ctypes.c_int32),
                ('time_sweep', ctypes.c_int32), ('wtab', _vp), ('spacing_f', ctypes.c_double * 3),
                ('spacing_i', _i64 * 3), ('spacing_is_int', ctypes.c_int32), ('halo_exchange', ctypes.c_int32),
                ('height', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('work', _vp), ('work_bytes', _i64), ('out', _vp), ('out_bytes', _i64), ('stream', _vp)]


class SlabResult(ctypes.Structure):
    """struct skb_slab_result (include/skan_b200.h)"""
    _fields_ = [('counts', _i64 * 32), ('off', _i64 * 32), ('work_need', _i64), ('out_need', _i64),
                ('exchange_need', _i64), ('host_ns', _i64 * 16), ('dev_ns', _i64 * 16)]


_SIGNATURES['skb_comm_create'] = (_int, [_int, _int, ctypes.c_char_p, _i64, ctypes.POINTER(_vp)])
_SIGNATURES['skb_comm_destroy'] = (_int, [_vp])
_SIGNATURES['skb_comm_reset'] = (_int, [_vp])
_SIGNATURES['skb_comm_allgather_i64'] = (_int, [_vp, _vp, _int, _vp])
_SIGNATURES['skb_run_slab'] = (_int, [ctypes.POINTER(SlabParams), _vp, ctypes.POINTER(SlabResult)])
SLAB_SECTIONS = ['voxels_nodes', 'junction_mst', 'csr_records', 'walk_rank', 'table', 'covered_select', 'cycles_heads',
                 'records', 'emit', 'skeleton_ids', 'stats_table(launch)']
# indices into SlabResult.off / .counts (the enums of skan_b200/csrc/skb_slab.inl)
SLAB_OFF = {n: i for i, n in enumerate(
    ['lin', 'coords', 'values', 'degrees', 'indptr', 'indices', 'data', 'plen', 'pidx', 'pdata', 'pw', 'dist', 'mean',
     'stdev', 't32', 't64', 'tf', 'start_node'])}
SLAB_COUNT = {n: i for i, n in enumerate(
    ['nodes_ext', 'own0', 'own1', 'id_shift', 'node_offset', 'edge_offset', 'edges_owned', 'nodes_total', 'edges_total',
     'regular', 'cycles', 'paths_total', 'entries_total', 'regular_total', 'regular_offset', 'cycle_offset',
     'entry_off_regular', 'entry_off_cycles', 'starts', 'table_records', 'syncs', 'exchanges', 'edges_ext',
     'sweep_ns', 'device_barriers'])}

# indices into RunResult.off / .counts (the enums of skan_b200/csrc/skb_run.inl)
RUN_OFF = {n: i for i, n in enumerate(
    ['lin', 'coords', 'values', 'degrees', 'indptr', 'indices', 'data', 'pindptr', 'pidx', 'pdata', 'pw', 'skel',
     'dist', 'mean', 'stdev', 't32', 't64', 'tf', 'comp_rank'])}
RUN_COUNT = {n: i for i, n in enumerate(
    ['nodes', 'edges', 'paths', 'regular', 'cycles', 'entries', 'junction_edges', 'starts', 'syncs', 'sweep_ns',
     'overlapped', 'retaken', 'junction_nodes'])}
RUN_HINTS = ['junction_edges', 'edges', 'starts', 'regular', 'junction_nodes']   # skb_run_params.hint[i] sizes by the count named

RUN_STAGES = ['start', 'sweep', 'nodes', 'raw_stencil', 'junction_count', 'junction_mst', 'degrees', 'csr_records',
              'walk', 'rank', 'select', 'heads', 'emit', 'cycles', 'skeleton_ids', 'stats', 'table']

EXPORTS = tuple(sorted(_SIGNATURES))


class NativeError(RuntimeError):
    pass


class Library:
    """A loaded C-ABI library; every call checks the status code."""

    def __init__(self, path, *, device_type='cuda'):
        self.path = path
        self.device_type = device_type
        self._dll = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(self._dll, name)
            fn.restype = res
            fn.argtypes = args
        if self._dll.skb_abi_version() != 1:
            raise NativeError(f'{path}: unexpected ABI version')
        self.tile_voxels = int(self._dll.skb_tile_voxels())
        self.launches = 0

    def kernel_launches(self):
        return int(self._dll.skb_kernel_launches())

    def scan_scratch_bytes(self, n):
        return int(self._dll.skb_scan_scratch_bytes(int(n)))

    def call(self, name, *args):
        status = getattr(self._dll, name)(*args)
        self.launches += 1
        if status < 0:
            msg = self._dll.skb_last_error().decode('utf-8', 'replace')
            raise NativeError(f'{name} failed ({status}): {msg}')
        return status


_lib = None


def library():
    """The CUDA library, loaded on first use.  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a). skan_b200 has no CPU fallback.')
        _lib = Library(LIB_PATH, device_type='cuda')
    return _lib


def _install_library_for_tests(lib):
    """tests/ only: route the package through another Library (the host emulation harness of
    tests/emul) so that host-side logic can be unit-tested in a container without a GPU."""
    global _lib
    old = _lib
    _lib = lib
    return old
