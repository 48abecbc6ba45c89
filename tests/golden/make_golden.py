"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src/skan through tests/golden/_refshim.py (stand-ins for the three
imports that are not installed), runs reference `Skeleton(...)` and
`summarize(...)` on the reference's own fixtures (src/skan/_testdata.py:6-75), on the small
edge cases of SURVEY.md Appendix B, and on seeded synthetic / dense random inputs, and stores
inputs + outputs in tests/golden/golden_cases.npz.  The GPU box has no /root/reference, so
the tests only read the committed .npz.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _refshim  # noqa: E402
from skan_b200.synth import walks  # noqa: E402


def cases():
    fx = _refshim.reference_fixtures()
    out = []
    for name, img in fx.items():
        if name == 'topograph1d':
            out.append((name + '_height', img, 1, True))
            out.append((name, img, 1, False))
        else:
            out.append((name, img, 1, False))
    out.append(('skeleton3d_sp511', fx['skeleton3d'], [5, 1, 1], False))     # test_csr.py:118-124
    out.append(('skeleton2_sp2', fx['skeleton2'], 2, False))                  # test_csr.py:90-101
    out.append(('skeleton1_sp_f', fx['skeleton1'], 2.5, False))
    out.append(('skeleton0_float', fx['skeleton0'].astype(np.float32) * 3.5, 1, False))
    # Appendix-B edge cases
    out.append(('row_gap', np.array([[1, 1, 0, 1]], bool), 1, False))
    out.append(('block2', np.ones((2, 2), bool), 1, False))
    out.append(('block3', np.ones((3, 3), bool), 1, False))
    plus = np.zeros((3, 3), bool); plus[1, :] = True; plus[:, 1] = True
    out.append(('plus', plus, 1, False))
    out.append(('block3d', np.ones((3, 3, 3), bool), (0.5, 0.1, 0.1), False))
    # seeded synthetic skeletons
    out.append(('walk2d_a', walks((96, 128), 24, 96, seed=11, isolated=6, rings=3), 1, False))
    out.append(('walk2d_b', walks((160, 160), 48, 128, seed=12, isolated=10, rings=4), (1.5, 0.7), False))
    out.append(('walk3d_a', walks((24, 40, 40), 16, 64, seed=13, isolated=5, rings=3), (0.5, 0.1, 0.1), False))
    out.append(('walk3d_b', walks((33, 31, 37), 40, 64, seed=14, isolated=5, rings=2), 1, False))
    # valued images (path_means / stdev over the skeleton image's own values, csr.py:648,792-816)
    rng = np.random.default_rng(21)
    w = walks((96, 96), 24, 96, seed=15, isolated=4, rings=2)
    out.append(('walk2d_f32', (w * (rng.random(w.shape) * 0.9 + 0.1)).astype(np.float32), 1, False))
    out.append(('walk2d_f64_h', (w * (rng.random(w.shape) * 4 + 1)).astype(np.float64), (1.0, 2.0), True))
    out.append(('walk2d_u8', (w * rng.integers(1, 200, w.shape)).astype(np.uint8), 1, False))
    w3 = walks((16, 24, 24), 10, 48, seed=16, isolated=2, rings=1)
    out.append(('walk3d_i16', (w3 * rng.integers(-50, 50, w3.shape)).astype(np.int16), (2, 1, 1), False))
    # dense random blobs: adversarial junction clusters with many equal-weight ties (A.4)
    for i, (shape, p, sp) in enumerate([((12, 12), 0.5, 1), ((16, 16), 0.35, (1.0, 1.0)), ((14, 10), 0.6, (0.3, 0.7)),
                                        ((6, 7, 8), 0.3, 1), ((7, 7, 7), 0.2, (0.5, 0.1, 0.1)),
                                        ((5, 6, 6), 0.45, (1.3, 0.9, 1.1)), ((24, 24), 0.42, 1),
                                        ((9, 9, 9), 0.25, (2, 1, 1))]):
        r = np.random.default_rng(100 + i)
        out.append((f'dense{i}', r.random(shape) < p, sp, False))
    return out


def run_reference(csr, img, spacing, height):
    rec = {}
    try:
        sk = csr.Skeleton(img, spacing=spacing, value_is_height=height)
    except ValueError as e:
        rec['error'] = np.array(str(e))
        return rec
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        df = csr.summarize(sk, value_is_height=height, separator='_')
    rec['graph_indptr'] = sk.graph.indptr
    rec['graph_indices'] = sk.graph.indices
    rec['graph_data'] = sk.graph.data
    rec['coordinates'] = sk.coordinates
    rec['degrees'] = sk.degrees
    rec['paths_indptr'] = sk.paths.indptr
    rec['paths_indices'] = sk.paths.indices
    rec['paths_data'] = sk.paths.data
    rec['paths_shape'] = np.array(sk.paths.shape)
    rec['path_lengths'] = sk.path_lengths()
    rec['path_means'] = sk.path_means()
    rec['path_stdev'] = sk.path_stdev()
    rec['path_label_image'] = sk.path_label_image()
    if sk.pixel_values is not None:
        rec['pixel_values'] = sk.pixel_values
    rec['summary_columns'] = np.array(list(df.columns))
    rec['summary_dtypes'] = np.array([str(t) for t in df.dtypes])
    for c in df.columns:
        rec['summary__' + c] = df[c].to_numpy()
    return rec


def main():
    csr = _refshim.import_reference()
    store = {}
    names = []
    for name, img, spacing, height in cases():
        names.append(name)
        store[f'{name}/image'] = np.asarray(img)
        store[f'{name}/spacing'] = np.asarray(spacing)
        store[f'{name}/spacing_is_scalar'] = np.array(np.isscalar(spacing))
        store[f'{name}/height'] = np.array(bool(height))
        rec = run_reference(csr, np.asarray(img), spacing, height)
        for k, v in rec.items():
            store[f'{name}/{k}'] = v
        print(name, np.asarray(img).shape, 'error' if 'error' in rec else
              f"N={rec['degrees'].size} E={rec['graph_indices'].size} P={rec['paths_indptr'].size - 1}")
    store['__names__'] = np.array(names)
    # error cases pinned by SURVEY.md Appendix B / test_csr.py:246-253
    out = os.path.join(HERE, 'golden_cases.npz')
    np.savez_compressed(out, **store)
    print('wrote', out, os.path.getsize(out), 'bytes')


if __name__ == '__main__':
    main()
