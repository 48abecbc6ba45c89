import numpy as np

from oracle import skan_oracle as orc
from skan_b200.synth import walks


def make_stack(shape, seed):
    """Images with paths, one empty image, one with isolated pixels only, one touching its borders."""
    B, H, W = shape
    stack = np.zeros(shape, bool)
    for i in range(B):
        if i == 1:
            continue                                   # empty image
        if i == 2:
            stack[i, ::7, ::9] = True                  # isolated pixels only: no path
            continue
        stack[i] = walks((H, W), 6 + i, max(H, W), seed=seed + i, isolated=3, rings=2)
    stack[0, 0, :] = True                              # a line along the image border
    stack[-1, -1, :] = True                            # last row of the last image / first row of nothing
    return stack


def check_stack_against_oracle(stack, spacing):
    from skan_b200 import batch
    df = batch.summarize_images(stack, spacing=spacing)
    seen = set()
    for i in range(stack.shape[0]):
        sub = df[df['image_id'] == i]
        try:
            ref = orc.OracleSkeleton(stack[i], spacing=spacing)
        except ValueError:
            assert len(sub) == 0, f'image {i} has no path in the oracle'
            continue
        seen.add(i)
        rs = orc.summarize(ref)
        assert len(sub) == ref.n_paths, (i, len(sub), ref.n_paths)
        m = np.abs(rs['mean_pixel_value'])
        for k, want in rs.items():
            got = sub[k].to_numpy()
            want = np.asarray(want)
            if want.dtype.kind in 'iub':
                assert np.array_equal(got, want), (i, k)
                assert got.dtype == want.dtype, (i, k, got.dtype, want.dtype)
            elif k.startswith('stdev'):
                assert np.all(np.abs(got**2 - want**2) <= 1e-12 * np.maximum(m * m, 1e-300) + 1e-12 * want**2), (i, k)
            else:
                np.testing.assert_allclose(got, want, rtol=1e-12, err_msg=f'{i} {k}')
    assert seen, 'no image of the stack had a path'
