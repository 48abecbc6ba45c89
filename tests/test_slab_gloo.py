"""Z-slab mode (one volume over several ranks) on CPU: world_size 2 and 3 with the gloo backend,
kernels through the host emulation harness, every rank's share gathered on rank 0 and compared
with the CPU oracle run on the whole volume (tests/slab_worker.py)."""
import os
import random
import subprocess
import sys

import pytest

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


def test_junction_clusters_across_slab_boundaries():
    # a dense volume: junction clusters percolate through every slab boundary; the junction MST is
    # solved on the gathered junction edges, so the result must still equal the oracle's
    codes, outs = launch(2, '--kind', 'dense12', '--shape', '40,16,16', '--halo', '2', '--seed', '5')
    assert all(c == 0 for c in codes), '\n'.join(o[-2000:] for o in outs)
    assert 'SLAB-OK' in outs[0]
