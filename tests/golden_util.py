"""Helpers shared by the parity tests: golden-vector access and the comparison rules
(bit-exact for integer/index arrays, rtol 1e-12 for fp64 sums; BASELINE.json north_star)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'golden_cases.npz')
RTOL = 1e-12

_cache = {}


def golden():
    if 'z' not in _cache:
        _cache['z'] = np.load(GOLDEN, allow_pickle=False)
    return _cache['z']


def case_names():
    return [str(s) for s in golden()['__names__']]


def case(name):
    z = golden()
    pre = name + '/'
    rec = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    sp = rec['spacing']
    if bool(rec['spacing_is_scalar']):
        sp = sp.item()
    else:
        sp = sp.tolist() if sp.dtype.kind != 'f' else tuple(float(x) for x in sp)
    rec['spacing_arg'] = sp
    rec['height'] = bool(rec['height'])
    return rec


def assert_same_int(a, b, what=''):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    assert np.array_equal(a, b), f'{what}: integer arrays differ'


def assert_close(a, b, what='', rtol=RTOL, atol=0.0):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)


def check_skeleton_against(rec, sk, summary, *, exact_weights=True):
    """sk: object with graph/coordinates/degrees/paths/path_lengths()/path_means()/path_stdev();
    summary: mapping column -> array (underscore separator)."""
    assert_same_int(sk.graph.indptr, rec['graph_indptr'], 'graph.indptr')
    assert_same_int(sk.graph.indices, rec['graph_indices'], 'graph.indices')
    if exact_weights:
        assert np.array_equal(np.asarray(sk.graph.data), rec['graph_data']), 'graph.data not bit-equal'
    else:
        assert_close(sk.graph.data, rec['graph_data'], 'graph.data')
    assert_same_int(sk.coordinates, rec['coordinates'], 'coordinates')
    assert_same_int(sk.degrees, rec['degrees'], 'degrees')
    assert_same_int(sk.paths.indptr, rec['paths_indptr'], 'paths.indptr')
    assert_same_int(sk.paths.indices, rec['paths_indices'], 'paths.indices')
    assert tuple(sk.paths.shape) == tuple(rec['paths_shape'])
    assert_close(sk.paths.data, rec['paths_data'], 'paths.data')
    assert_close(sk.path_lengths(), rec['path_lengths'], 'path_lengths')
    assert_close(sk.path_means(), rec['path_means'], 'path_means')
    m = np.abs(rec['path_means'])
    # stdev = sqrt(E[x^2]-E[x]^2) suffers cancellation: compare variances with atol ~ 1e-12*mean^2
    var_a = np.asarray(sk.path_stdev(), float)**2
    var_b = rec['path_stdev']**2
    assert np.all(np.abs(var_a - var_b) <= 1e-12 * np.maximum(m * m, 1e-300) + 1e-12 * var_b), 'path_stdev'
    cols = [str(c) for c in rec['summary_columns']]
    assert list(summary.keys()) == cols, f'summary columns {list(summary.keys())} vs {cols}'
    for c in cols:
        ref = rec['summary__' + c]
        got = np.asarray(summary[c])
        if ref.dtype.kind in 'iub':
            assert_same_int(got, ref, 'summary ' + c)
        elif c.startswith('stdev'):
            assert np.all(np.abs(got**2 - ref**2) <= 1e-12 * np.maximum(m * m, 1e-300) + 1e-12 * ref**2), c
        else:
            assert_close(got, ref, 'summary ' + c)
