"""skeleton_to_nx / nx_to_skeleton (reference csr.py:1393-1480), modelled on the reference's own tests
(test_csr.py:495-587): one edge per path, one node per junction / endpoint, and the round trip through the
regenerated image."""
import warnings

import numpy as np
import pytest
import torch

from golden_util import case

NAMES = ['tinycycle', 'tinyline', 'skeleton0', 'skeleton1', 'skeleton3d', 'skeleton2']


def check(name, to_device):
    import skan_b200
    from skan_b200 import csr
    rec = case(name)
    img = np.asarray(rec['image'])
    sk = skan_b200.Skeleton(to_device(img))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        summary = csr.summarize(sk)          # '-' separated columns, as the reference's tests pass them
    for table in (summary, None):
        g = csr.skeleton_to_nx(sk, table)
        assert g.number_of_edges() == sk.n_paths == int(rec['paths_shape'][0])
        ends = np.unique(np.concatenate([summary['node-id-src'], summary['node-id-dst']]))
        assert g.number_of_nodes() == len(ends)
        assert g.graph['shape'] == tuple(img.shape) and g.graph['dtype'] == img.dtype
    # edges come back in adjacency order, not path order: compare as multisets of (node ids, coordinates)
    got = sorted((tuple(d['indices'].tolist()), tuple(map(tuple, d['path'].tolist()))) for _, _, d in g.edges(data=True))
    want = sorted((tuple(sk.path(k).tolist()), tuple(map(tuple, sk.path_coordinates(k).tolist())))
                  for k in range(sk.n_paths))
    assert got == want
    back = csr.nx_to_skeleton(g)
    np.testing.assert_array_equal(np.asarray(back.skeleton_image), img)
    assert np.array_equal(back.graph.indices, sk.graph.indices) and np.array_equal(back.graph.data, sk.graph.data)


@pytest.mark.parametrize('name', NAMES)
def test_nx_round_trip_emulated(name):
    from emul_util import emulated_device
    with emulated_device():
        check(name, lambda a: a)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_nx_round_trip_gpu(name):
    check(name, lambda a: a)


def test_nx_to_skeleton_errors_emulated():
    # test_csr.py:572-587: arrays, Skeletons and graphs without the attributes raise
    import networkx as nx
    import skan_b200
    from skan_b200 import csr
    from emul_util import emulated_device
    img = np.asarray(case('skeleton0')['image'])
    with emulated_device():
        for wrong in (img, skan_b200.Skeleton(img), nx.Graph()):
            with pytest.raises(Exception):
                csr.nx_to_skeleton(wrong)
