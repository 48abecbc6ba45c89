"""Seeded synthetic skeleton generator.

Used by the tests, the golden-vector script and bench.py to make "neuron-like" inputs of the
BASELINE.json shapes: correlated random walks with Chebyshev-normalised steps (consecutive
voxels differ by at most one per axis, so the lines are 3^d-connected and mostly one voxel
wide), periodic re-seeding of walkers onto other walkers (branching -> multi-voxel junction
clusters), stamped diamond rings (junction-free cycles) and isolated voxels.

The generator produces *coordinates* on the host (vectorised over walkers); `walks_points`
returns them as raveled indices so that a caller can scatter them into a host array
(`walks`) or straight into a device tensor (`walks_device`) without ever building the dense
volume on the host.
"""
from __future__ import annotations

import numpy as np


def walks_points(shape, n_walkers, steps, seed, sigma=0.15, isolated=0, rings=0):
    """Raveled int64 indices (unsorted, with duplicates) of the foreground voxels."""
    shape = tuple(int(s) for s in shape)
    d = len(shape)
    rng = np.random.default_rng(seed)
    hi = np.asarray(shape, dtype=float)
    pos = rng.random((n_walkers, d)) * (hi - 1)
    direc = rng.normal(size=(n_walkers, d))
    direc /= np.linalg.norm(direc, axis=1, keepdims=True) + 1e-300
    strides = np.cumprod((shape[1:] + (1,))[::-1])[::-1].astype(np.int64)
    chunks = []
    for t in range(steps):
        p = np.rint(pos).astype(np.int64)
        ok = np.all((p >= 0) & (p < np.asarray(shape)), axis=1)
        chunks.append((p[ok] * strides).sum(axis=1))
        direc += sigma * rng.normal(size=direc.shape)
        direc /= np.linalg.norm(direc, axis=1, keepdims=True) + 1e-300
        pos += direc / np.abs(direc).max(axis=1, keepdims=True)
        if t % 64 == 63 and n_walkers >= 8:
            k = n_walkers // 8
            who = rng.choice(n_walkers, size=k, replace=False)
            onto = rng.integers(0, n_walkers, size=k)
            pos[who] = pos[onto]
            nd = rng.normal(size=(k, d))
            direc[who] = nd / (np.linalg.norm(nd, axis=1, keepdims=True) + 1e-300)
    if rings and d >= 2:
        # diamond ring of 4 voxels (the reference's `tinycycle`, _testdata.py:6-8) in the last two axes
        c = np.stack([rng.integers(1, max(2, s - 1), size=rings) for s in shape], axis=-1)
        for dy, dx in ((-1, 0), (0, -1), (0, 1), (1, 0)):
            q = c.copy()
            q[:, -2] += dy
            q[:, -1] += dx
            ok = np.all((q >= 0) & (q < np.asarray(shape)), axis=1)
            chunks.append((q[ok] * strides).sum(axis=1))
    if isolated:
        q = np.stack([rng.integers(0, s, size=isolated) for s in shape], axis=-1)
        chunks.append((q * strides).sum(axis=1))
    return np.concatenate(chunks) if chunks else np.zeros(0, np.int64)


def walks(shape, n_walkers, steps, seed, sigma=0.15, isolated=0, rings=0):
    """Dense boolean host array of the given shape."""
    img = np.zeros(int(np.prod(shape)), dtype=bool)
    img[walks_points(shape, n_walkers, steps, seed, sigma, isolated, rings)] = True
    return img.reshape(shape)


def walks_device(shape, n_walkers, steps, seed, device, sigma=0.15, isolated=0, rings=0, dtype=None):
    """Same skeleton, scattered into a torch tensor on `device` (bool unless `dtype` given)."""
    import torch
    pts = walks_points(shape, n_walkers, steps, seed, sigma, isolated, rings)
    v = int(np.prod(shape))
    img = torch.zeros(v, dtype=dtype or torch.bool, device=device)
    idx = torch.from_numpy(pts).to(device)
    img[idx] = True if (dtype is None or dtype == torch.bool) else 1
    return img.reshape(tuple(int(s) for s in shape))


def stack_walks_points(n_images, height, width, walkers_per_image, steps, seed, sigma=0.15, isolated=0):
    """Raveled int64 indices into a (n_images, height, width) stack of independent 2-D skeletons:
    the walks of `walks_points`, vectorised over all images at once (walkers never leave their image,
    branching re-seeds a walker onto another walker of the same image)."""
    rng = np.random.default_rng(seed)
    B, H, W, k = int(n_images), int(height), int(width), int(walkers_per_image)
    n = B * k
    img_of = np.repeat(np.arange(B, dtype=np.int64), k)
    hi = np.array([H - 1, W - 1], dtype=float)
    pos = rng.random((n, 2)) * hi
    direc = rng.normal(size=(n, 2))
    direc /= np.linalg.norm(direc, axis=1, keepdims=True) + 1e-300
    plane = H * W
    chunks = []
    for t in range(steps):
        p = np.rint(pos).astype(np.int64)
        ok = (p[:, 0] >= 0) & (p[:, 0] < H) & (p[:, 1] >= 0) & (p[:, 1] < W)
        chunks.append(img_of[ok] * plane + p[ok, 0] * W + p[ok, 1])
        direc += sigma * rng.normal(size=direc.shape)
        direc /= np.linalg.norm(direc, axis=1, keepdims=True) + 1e-300
        pos += direc / np.abs(direc).max(axis=1, keepdims=True)
        if t % 64 == 63 and k >= 8:
            m = n // 8
            who = rng.choice(n, size=m, replace=False)
            onto = img_of[who] * k + rng.integers(0, k, size=m)
            pos[who] = pos[onto]
            nd = rng.normal(size=(m, 2))
            direc[who] = nd / (np.linalg.norm(nd, axis=1, keepdims=True) + 1e-300)
    if isolated:
        q = rng.integers(0, plane, size=B * isolated) + np.repeat(np.arange(B, dtype=np.int64), isolated) * plane
        chunks.append(q)
    return np.concatenate(chunks) if chunks else np.zeros(0, np.int64)
