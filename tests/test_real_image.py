"""BASELINE configs[0] stand-in: a real image (the reference's doc/example-data/inf1.tif, thresholded and thinned by
tests/golden/make_real_image.py) with spacing 2.24826 as in benchmarks/bench_skan.py:27, against the outputs of the
unmodified reference: graph, coordinates, degrees, paths and every summarize column."""
import os
import warnings

import numpy as np
import pytest
import torch

from golden_util import check_skeleton_against

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    z = np.load(os.path.join(HERE, 'golden', 'real_image.npz'))
    rec = {k: z[k] for k in z.files}
    shape = tuple(int(x) for x in rec['shape'])
    rec['image'] = np.unpackbits(rec['image_bits'])[:shape[0] * shape[1]].astype(bool).reshape(shape)
    return rec


def check(to_device):
    import skan_b200
    rec = load()
    sk = skan_b200.Skeleton(to_device(rec['image']), spacing=float(rec['spacing']))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        df = skan_b200.summarize(sk, separator='_')
    check_skeleton_against(rec, sk, {c: df[c].to_numpy() for c in df.columns}, exact_weights=True)
    for c, t in zip(rec['summary_columns'], rec['summary_dtypes']):
        assert str(df[str(c)].dtype) == str(t), (c, t)
    g, coords = skan_b200.skeleton_to_csgraph(to_device(rec['image']), spacing=float(rec['spacing']))
    assert np.array_equal(g.indptr, rec['graph_indptr']) and np.array_equal(g.indices, rec['graph_indices'])
    assert np.array_equal(g.data, rec['graph_data'])
    assert np.array_equal(np.stack(coords, axis=1), rec['coordinates'])
    return sk


def test_real_image_emulated():
    from emul_util import emulated_device
    with emulated_device():
        sk = check(lambda a: a)
    assert sk.n_paths == 12423 and sk.graph.shape[0] == 80652


@pytest.mark.gpu
def test_real_image_gpu():
    check(lambda a: torch.from_numpy(a).cuda())
