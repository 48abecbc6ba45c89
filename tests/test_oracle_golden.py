"""The oracle (oracle/skan_oracle.py) pinned against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py), and against the known answers of the reference's own
tests (src/skan/test/test_csr.py, test_skeleton_class.py)."""
import numpy as np
import pytest

from golden_util import case, case_names, check_skeleton_against
from oracle import skan_oracle as orc


@pytest.mark.parametrize('name', case_names())
def test_oracle_matches_reference_golden(name):
    rec = case(name)
    if 'error' in rec:
        with pytest.raises(ValueError):
            orc.OracleSkeleton(rec['image'], spacing=rec['spacing_arg'], value_is_height=rec['height'])
        return
    sk = orc.OracleSkeleton(rec['image'], spacing=rec['spacing_arg'], value_is_height=rec['height'])
    summ = orc.summarize(sk, value_is_height=rec['height'])
    check_skeleton_against(rec, sk, summ, exact_weights=True)
    assert np.array_equal(sk.path_label_image(), rec['path_label_image'])
    # dtypes of the summary table
    for c, t in zip(rec['summary_columns'], rec['summary_dtypes']):
        assert str(np.asarray(summ[str(c)]).dtype) == str(t), (c, t)


@pytest.mark.parametrize('name', [n for n in case_names() if not case(n)['height'] and 'error' not in case(n)])
def test_mst_rule_equals_scipy_route(name):
    rec = case(name)
    raw, _ = orc.raw_pixel_graph(rec['image'], rec['spacing_arg'], False)
    a = orc.mst_junctions(raw)
    b = orc.mst_junctions_rule(raw)
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    assert np.array_equal(a.data, b.data)


def test_known_answers_from_reference_tests():
    # test_csr.py:53-63
    sk = orc.OracleSkeleton(case('tinycycle')['image'])
    assert sk.graph.indptr.tolist() == [0, 2, 4, 6, 8]
    assert sk.graph.indices.tolist() == [1, 2, 0, 3, 0, 3, 1, 2]
    assert np.allclose(sk.graph.data, np.sqrt(2))
    # test_skeleton_class.py:18-30 (direction pinned by SURVEY.md Appendix B)
    sk = orc.OracleSkeleton(case('skeleton1')['image'])
    assert [list(map(int, p)) for p in sk.paths_list()] == [
        [7, 5, 0, 1, 2, 3, 4, 6, 10, 9, 12], [7, 8, 12], [7, 11, 13], [12, 14, 15, 16]]
    assert tuple(sk.paths.shape) == (4, 21)
    # test_csr.py:112-115
    sk = orc.OracleSkeleton(case('tinycycle')['image'])
    s = orc.summarize(sk)
    assert s['branch_type'].tolist() == [3] and np.isclose(s['branch_distance'][0], 4 * np.sqrt(2))
    # test_csr.py:118-124
    sk = orc.OracleSkeleton(case('skeleton3d')['image'], spacing=[5, 1, 1])
    s = orc.summarize(sk)
    assert len(s['branch_type']) == 7
    assert (s['node_id_src'][0], s['node_id_dst'][0], s['branch_type'][0]) == (0, 10, 1)
    assert np.isclose(s['branch_distance'][0], 2 * np.sqrt(27))
    # test_csr.py:144-161: the T-junction of skeleton0 loses its diagonal edges
    sk = orc.OracleSkeleton(case('skeleton0')['image'])
    assert sk.graph[2, 5] == 0 and np.all(sk.graph.data == 1.0)
    # test_csr.py:308-328 zero-degree nodes
    img = np.zeros((5, 8), bool); img[1, 1:4] = True; img[3, 5] = True; img[3, 0:3] = True
    sk = orc.OracleSkeleton(img)
    assert sk.n_paths == 2 and (sk.degrees == 0).sum() == 1


@pytest.mark.parametrize('img', [np.zeros((4, 4), bool), np.eye(1, dtype=bool),
                                 np.array([[1, 0, 0, 1]], bool)])
def test_zero_path_skeleton_raises(img):
    # csr.py:442 raises ValueError when there is no path
    with pytest.raises(ValueError):
        orc.OracleSkeleton(img)
