"""Z-slab mode (one volume over several ranks) on CPU: world_size 2 and 3 with the gloo backend,
kernels through the host emulation harness, every rank's share gathered on rank 0 and compared
with the CPU oracle run on the whole volume (tests/slab_worker.py)."""
import os
import random
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, 'slab_worker.py')


def launch(world, *args, backend='gloo', emulate=1, timeout=420):
    port = random.randint(20000, 40000)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_PORT=str(port), MASTER_ADDR='127.0.0.1')
        procs.append(subprocess.Popen([sys.executable, WORKER, '--backend', backend, '--emulate', str(emulate), *args],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    return [p.returncode for p in procs], outs


@pytest.mark.parametrize('world,args', [
    (2, ['--kind', 'walks', '--shape', '24,20,28', '--halo', '2']),
    (3, ['--kind', 'walks', '--shape', '48,32,32', '--halo', '5', '--seed', '7', '--valued', '1']),
    (2, ['--kind', 'lines_z', '--shape', '40,21,30', '--halo', '3']),
    (3, ['--kind', 'dense', '--shape', '60,20,20', '--halo', '2', '--seed', '6']),
])
def test_slab_mode_matches_oracle(world, args):
    codes, outs = launch(world, *args)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0]


@pytest.mark.parametrize('world,args', [
    (2, ['--kind', 'walks', '--shape', '24,20,28', '--halo', '2']),
    (3, ['--kind', 'dense', '--shape', '60,20,20', '--halo', '2', '--seed', '6']),
])
def test_slab_mode_python_sequencing_matches_oracle(world, args):
    """The per-stage python sequencing (collectives through torch.distributed) stays covered: it is what
    runs when the ranks are not all on one node."""
    codes, outs = launch(world, *args, '--native', '0')
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'native=0' in outs[0]


@pytest.mark.parametrize('world,args', [
    (2, ['--kind', 'walks', '--shape', '24,20,28', '--halo', '2', '--valued', '1', '--height', '1']),
    (3, ['--kind', 'walks', '--shape', '48,32,32', '--halo', '5', '--seed', '7', '--valued', '1', '--height', '1']),
    (3, ['--kind', 'rings', '--shape', '48,22,30', '--halo', '3', '--valued', '1', '--height', '1']),
    (2, ['--kind', 'dense12', '--shape', '32,16,16', '--halo', '2', '--seed', '4', '--valued', '1', '--height', '1']),
])
def test_value_is_height_in_slab_mode(world, args):
    """Skeleton(value_is_height=True) (csr.py:148-151, 932-948): edge weights hypot(height difference, distance) in the
    graph, in the junction MST (weights travel with the gathered edges) and along the paths; the pixel value of a
    path's far end travels with the ranking records and enters euclidean_distance as an extra coordinate."""
    codes, outs = launch(world, *args)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'native=1' in outs[0]


@pytest.mark.parametrize('world,args', [
    (2, ['--kind', 'walks', '--shape', '24,20,28', '--seed', '5']),                                  # planes of 560 voxels
    (3, ['--kind', 'walks', '--shape', '48,32,32', '--seed', '7', '--valued', '1']),                 # valued image
    (3, ['--kind', 'rings', '--shape', '48,22,30', '--valued', '1', '--height', '1']),               # value_is_height
    (2, ['--kind', 'dense12', '--shape', '32,16,16', '--seed', '4', '--native', '0']),               # python sequencing
])
def test_halo_exchange_of_image_planes_through_the_process_group(world, args):
    """halo='exchange' where the packed-mask exchange of the native runner does not apply (valued images, value_is_height,
    planes that are not whole mask tiles, python sequencing): the two boundary planes of the image travel through
    torch.distributed point-to-point operations, then the extended slab goes the halo='input' way."""
    codes, outs = launch(world, *args, '--halo-exchange', '1')
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'xhalo=1' in outs[0]


def test_native_runner_is_the_default():
    codes, outs = launch(2, '--kind', 'walks', '--shape', '24,20,28', '--halo', '2', '--seed', '3')
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'native=1' in outs[0]


def test_native_runner_restarts_when_arenas_are_too_small():
    """Tiny work / out / exchange arenas on the first call: ranks give up at different points, the others
    are released from their exchange (abort word), everybody grows and runs again."""
    codes, outs = launch(3, '--kind', 'walks', '--shape', '48,32,32', '--halo', '5', '--seed', '7', '--small-arenas', '1')
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'native=1' in outs[0]


@pytest.mark.parametrize('world,args', [
    (2, ['--kind', 'rings', '--shape', '40,22,30', '--halo', '2']),
    (3, ['--kind', 'rings', '--shape', '48,22,30', '--halo', '3', '--valued', '1']),
])
def test_cycles_across_slab_boundaries(world, args):
    """Junction-free cycles that cross one or several slab boundaries, lie astride one, or lie in a boundary
    plane: the ranks gather the uncovered chain voxels and the owner of a ring's minimum node emits it."""
    codes, outs = launch(world, *args)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'native=1' in outs[0]


def test_comm_allgather_between_processes():
    """The shared-memory count channel of the native runner (skb_comm_*), three processes."""
    code = (
        "import os, sys, ctypes, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(HERE)!r}); sys.path.insert(0, {HERE!r})\n"
        "from emul_util import build_emulation_library\n"
        "from skan_b200 import _native\n"
        "lib = _native.Library(build_emulation_library(), device_type='cpu')\n"
        "r, w = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])\n"
        "h = ctypes.c_void_p()\n"
        "lib.call('skb_comm_create', r, w, os.environ['COMM_NAME'].encode(), 1 << 16, ctypes.byref(h))\n"
        "for it in range(200):\n"
        "    v = np.array([r * 1000 + it, it, -r], dtype=np.int64)\n"
        "    out = np.zeros((w, 3), dtype=np.int64)\n"
        "    lib.call('skb_comm_allgather_i64', h, v.ctypes.data, 3, out.ctypes.data)\n"
        "    assert (out[:, 0] == np.arange(w) * 1000 + it).all() and (out[:, 1] == it).all() and (out[:, 2] == -np.arange(w)).all()\n"
        "lib.call('skb_comm_destroy', h)\n"
        "print('COMM-OK')\n")
    from emul_util import build_emulation_library
    build_emulation_library()
    name = 't%08x' % random.getrandbits(32)
    procs = [subprocess.Popen([sys.executable, '-c', code], env=dict(os.environ, RANK=str(r), WORLD_SIZE='3', COMM_NAME=name),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(3)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all('COMM-OK' in o for o in outs), '\n'.join(outs)


def test_junction_clusters_across_slab_boundaries():
    # a dense volume: junction clusters percolate through every slab boundary; the junction MST is
    # solved on the gathered junction edges, so the result must still equal the oracle's
    codes, outs = launch(2, '--kind', 'dense12', '--shape', '40,16,16', '--halo', '2', '--seed', '5')
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0]


@pytest.mark.parametrize('world,args', [
    (2, ['--kind', 'walks', '--shape', '24,128,128', '--seed', '2']),
    (3, ['--kind', 'dense', '--shape', '30,64,256', '--seed', '6']),
    (3, ['--kind', 'rings', '--shape', '48,128,128']),
    (2, ['--kind', 'walks', '--shape', '16,128,128', '--seed', '9', '--small-arenas', '1']),
])
def test_halo_exchange_from_a_disjoint_partition(world, args):
    """halo='exchange': every rank passes its owned planes only; the packed mask of the boundary planes travels through
    the exchange arenas.  The assembled result must equal the oracle's on the whole volume (and so the overlapped-input
    path's, which the tests above compare with the same oracle)."""
    codes, outs = launch(world, *args, '--halo-exchange', '1')
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0] and 'xhalo=1' in outs[0]


@pytest.mark.parametrize('world,args', [
    (2, ['--kind', 'walks', '--shape', '24,20,28', '--halo', '2', '--prime', '1']),
    (3, ['--kind', 'dense', '--shape', '60,20,20', '--halo', '2', '--seed', '6', '--prime', '2']),
    (2, ['--kind', 'dense12', '--shape', '40,16,16', '--halo', '2', '--seed', '5', '--prime', '2']),
    (2, ['--kind', 'walks', '--shape', '24,128,128', '--seed', '2', '--halo-exchange', '1', '--prime', '1']),
    (3, ['--kind', 'rings', '--shape', '48,22,30', '--halo', '3', '--valued', '1', '--guard', '1', '--prime', '1']),
])
def test_capacity_hints_of_the_previous_call(world, args):
    """The second call on a communicator sizes the junction-edge exchange and the edge arrays by the first call's counts and
    saves two synchronisations; hints that fit (--prime 1) and hints that are too small (--prime 2: a much sparser
    volume first) must both give the oracle's result."""
    codes, outs = launch(world, *args)
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0]
