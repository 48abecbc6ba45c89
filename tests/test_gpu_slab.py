"""Z-slab mode on real GPUs (one rank per GPU, exchanges through peer-mapped arenas): every rank's share is gathered
on rank 0 and compared with the CPU oracle on the whole volume, array by array and through the result digest.

Every case runs at each world size in WORLDS.  A world size larger than the number of visible GPUs is SKIPPED (and
reported as such), never shrunk to fewer ranks; with SKAN_B200_REQUIRE_RANKS=N in the environment (set it on a
multi-GPU box: `gpurun --gpus 2 -- 'SKAN_B200_REQUIRE_RANKS=2 python -m pytest tests/test_gpu_slab.py -m gpu'`) a
world size <= N that cannot run is a FAILURE."""
import os

import pytest
import torch

from test_slab_gloo import launch

pytestmark = pytest.mark.gpu

WORLDS = (1, 2, 4, 8)
CASES = [
    ['--kind', 'walks', '--shape', '96,64,64', '--halo', '2', '--seed', '11'],
    ['--kind', 'dense12', '--shape', '64,32,32', '--halo', '2', '--seed', '5'],
    ['--kind', 'lines_z', '--shape', '64,33,40', '--halo', '2'],
    ['--kind', 'walks', '--shape', '128,96,80', '--halo', '8', '--seed', '12', '--valued', '1'],
    ['--kind', 'rings', '--shape', '64,22,30', '--halo', '2'],
    ['--kind', 'walks', '--shape', '96,64,64', '--halo', '2', '--seed', '11', '--guard', '1', '--prime', '1'],
    ['--kind', 'dense12', '--shape', '64,32,32', '--halo', '2', '--seed', '5', '--guard', '1', '--prime', '2'],
    # tiny first arenas: every rank gives the call up at a different point, while its peers wait in a device-side barrier
    # (the abort word ends the waits, the guarded kernels are skipped, the read-back reports it, everybody goes again)
    ['--kind', 'walks', '--shape', '96,64,64', '--halo', '2', '--seed', '13', '--small-arenas', '1'],
    # value_is_height: hypot edge weights (graph, junction MST, paths), the far end's pixel value travels with the ranking
    ['--kind', 'walks', '--shape', '96,64,64', '--halo', '2', '--seed', '14', '--valued', '1', '--height', '1'],
    ['--kind', 'dense12', '--shape', '64,32,32', '--halo', '2', '--seed', '6', '--valued', '1', '--height', '1'],
]


def _need(world):
    have = torch.cuda.device_count()
    if have >= world:
        return
    required = int(os.environ.get('SKAN_B200_REQUIRE_RANKS', '0'))
    if world <= required:
        pytest.fail(f'{world} ranks were required (SKAN_B200_REQUIRE_RANKS={required}) but only {have} GPU(s) are visible')
    pytest.skip(f'needs {world} GPUs, {have} visible')


@pytest.mark.parametrize('world', WORLDS)
@pytest.mark.parametrize('args', CASES)
def test_slab_mode_native(world, args):
    _need(world)
    codes, outs = launch(world, *args, backend='nccl', emulate=0)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and f'world={world}' in outs[0]


@pytest.mark.parametrize('world', WORLDS)
@pytest.mark.parametrize('args', [
    ['--kind', 'walks', '--shape', '64,128,128', '--seed', '2'],
    ['--kind', 'dense12', '--shape', '48,64,256', '--seed', '6'],
    ['--kind', 'rings', '--shape', '64,128,128'],
    ['--kind', 'walks', '--shape', '64,128,128', '--seed', '4', '--small-arenas', '1'],
    # not served by the packed-mask exchange of the native runner: the image's boundary planes go through the process group
    ['--kind', 'walks', '--shape', '64,96,80', '--seed', '8', '--valued', '1'],
    ['--kind', 'walks', '--shape', '64,96,80', '--seed', '9', '--valued', '1', '--height', '1'],
])
def test_halo_exchange_from_a_disjoint_partition(world, args):
    """halo='exchange': owned planes only per rank, boundary planes of the packed mask exchanged over NVLink."""
    _need(world)
    codes, outs = launch(world, *args, '--halo-exchange', '1', backend='nccl', emulate=0)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'xhalo=1' in outs[0] and f'world={world}' in outs[0]


@pytest.mark.parametrize('world', (1, 2))
def test_slab_mode_python_sequencing_nccl(world):
    _need(world)
    codes, outs = launch(world, '--kind', 'walks', '--shape', '96,64,64', '--halo', '2', '--seed', '11', '--native', '0',
                         backend='nccl', emulate=0)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'native=0' in outs[0]
