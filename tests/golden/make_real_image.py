"""Golden vectors for a REAL image, the stand-in for BASELINE configs[0] (benchmarks/infected3.npz is an LFS pointer in
the reference tree): doc/example-data/inf1.tif of the reference (1094 x 1536 uint8 SEM image of a spectrin network) is
thresholded and thinned by the test tooling below, and the UNMODIFIED reference runs `Skeleton(skeleton,
spacing=2.24826)` + `summarize` on it as benchmarks/bench_skan.py:27-45 does.  How the mask was thinned is irrelevant
to parity -- the reference and the CUDA path see the same mask.

    python tests/golden/make_real_image.py      ->  tests/golden/real_image.npz   (mask packed to bits + reference outputs)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

SPACING = 2.24826            # bench_skan.py:27


def zhang_suen(img):
    """Zhang-Suen thinning, vectorised (test tooling; any thinning would do)."""
    img = np.pad(img.astype(bool), 1)
    while True:
        changed = False
        for step in (0, 1):
            p2, p3, p4 = img[:-2, 1:-1], img[:-2, 2:], img[1:-1, 2:]
            p5, p6, p7 = img[2:, 2:], img[2:, 1:-1], img[2:, :-2]
            p8, p9 = img[1:-1, :-2], img[:-2, :-2]
            nb = [p2, p3, p4, p5, p6, p7, p8, p9]
            b = sum(x.astype(np.int8) for x in nb)
            a = sum(((~nb[i]) & nb[(i + 1) % 8]).astype(np.int8) for i in range(8))
            cond = img[1:-1, 1:-1] & (b >= 2) & (b <= 6) & (a == 1)
            if step == 0:
                cond &= ~(p2 & p4 & p6) & ~(p4 & p6 & p8)
            else:
                cond &= ~(p2 & p4 & p8) & ~(p2 & p6 & p8)
            if cond.any():
                img[1:-1, 1:-1] &= ~cond
                changed = True
        if not changed:
            return img[1:-1, 1:-1]


def skeleton_of_example():
    from PIL import Image
    from scipy import ndimage as ndi
    im = np.asarray(Image.open('/root/reference/doc/example-data/inf1.tif')).astype(np.float64)
    smooth = ndi.gaussian_filter(im, 1.5)
    local = ndi.uniform_filter(smooth, 31)
    binary = smooth > local + 2.0
    binary = ndi.binary_opening(binary, iterations=1)
    return zhang_suen(binary)


def main():
    import _refshim
    from make_golden import run_reference
    csr = _refshim.import_reference()
    skel = skeleton_of_example()
    print('skeleton', skel.shape, int(skel.sum()), 'pixels', f'{skel.mean() * 100:.2f} %')
    rec = run_reference(csr, skel, SPACING, False)
    g, coords = csr.skeleton_to_csgraph(skel, spacing=SPACING)
    assert np.array_equal(g.indptr, rec['graph_indptr'])
    store = {'image_bits': np.packbits(skel.reshape(-1)), 'shape': np.array(skel.shape), 'spacing': np.array(SPACING)}
    rec.pop('path_label_image')        # V-sized; the small cases cover it
    for k, v in rec.items():
        store[k] = v
    out = os.path.join(HERE, 'real_image.npz')
    np.savez_compressed(out, **store)
    print(f"N={rec['degrees'].size} E={rec['graph_indices'].size} P={rec['paths_indptr'].size - 1}", 'wrote', out,
          os.path.getsize(out), 'bytes')


if __name__ == '__main__':
    main()
