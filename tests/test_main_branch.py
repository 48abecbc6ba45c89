"""summarize(find_main_branch=True): the host-side post-processing of the branch table against golden vectors
produced by the unmodified reference (tests/golden/make_main_branch.py)."""
import os
import warnings

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'main_branch.npz'))
NAMES = sorted({k.split('/')[0] for k in GOLD.files})


def load(name):
    shape = tuple(int(x) for x in GOLD[name + '/shape'])
    img = np.unpackbits(GOLD[name + '/image'])[:int(np.prod(shape))].astype(bool).reshape(shape)
    sp = GOLD[name + '/spacing']
    spacing = float(sp[0]) if sp.size == 1 else tuple(float(x) for x in sp)
    return img, spacing, GOLD[name + '/main']


@pytest.mark.parametrize('name', NAMES)
def test_main_branches_of_the_oracle_table(name):
    from oracle import skan_oracle as orc
    from skan_b200.summary_utils import find_main_branches
    img, spacing, want = load(name)
    rs = orc.summarize(orc.OracleSkeleton(img, spacing=spacing))
    assert np.array_equal(find_main_branches(rs), want)


@pytest.mark.parametrize('name', NAMES)
def test_summarize_find_main_branch_emulated(name):
    import torch
    import skan_b200
    from emul_util import emulated_device
    img, spacing, want = load(name)
    with emulated_device():
        sk = skan_b200.Skeleton(torch.from_numpy(img), spacing=spacing)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            df = skan_b200.summarize(sk, find_main_branch=True)
    assert list(df.columns)[-1] == 'main' and df['main'].dtype == bool
    assert np.array_equal(df['main'].to_numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_summarize_find_main_branch_gpu(name):
    import torch
    import skan_b200
    img, spacing, want = load(name)
    sk = skan_b200.Skeleton(torch.from_numpy(img).cuda(), spacing=spacing)
    df = skan_b200.summarize(sk, find_main_branch=True, separator='_')
    assert np.array_equal(df['main'].to_numpy(), want)
