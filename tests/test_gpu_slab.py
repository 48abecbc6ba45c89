"""Z-slab mode on real GPUs over NCCL (run with as many ranks as GPUs are visible, at least 1):
every rank's share is gathered on rank 0 and compared with the CPU oracle on the whole volume."""
import pytest
import torch

from test_slab_gloo import launch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('args', [
    ['--kind', 'walks', '--shape', '96,64,64', '--halo', '2', '--seed', '11'],
    ['--kind', 'dense12', '--shape', '64,32,32', '--halo', '2', '--seed', '5'],
    ['--kind', 'lines_z', '--shape', '64,33,40', '--halo', '2'],
    ['--kind', 'walks', '--shape', '128,96,80', '--halo', '8', '--seed', '12', '--valued', '1'],
    ['--kind', 'rings', '--shape', '64,22,30', '--halo', '2'],
])
def test_slab_mode_nccl(args):
    world = max(1, min(torch.cuda.device_count(), 8))
    codes, outs = launch(world, *args, backend='nccl', emulate=0)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0]


def test_slab_mode_python_sequencing_nccl():
    world = max(1, min(torch.cuda.device_count(), 8))
    codes, outs = launch(world, '--kind', 'walks', '--shape', '96,64,64', '--halo', '2', '--seed', '11', '--native', '0',
                         backend='nccl', emulate=0)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'native=0' in outs[0]


def test_slab_mode_two_ranks_on_one_gpu_is_rejected_or_passes():
    """With a single GPU the 2-rank run shares the device; NCCL refuses duplicate devices, so only
    run the multi-rank case when at least two GPUs are visible."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    codes, outs = launch(2, '--kind', 'walks', '--shape', '64,48,48', '--halo', '5', backend='nccl', emulate=0)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
