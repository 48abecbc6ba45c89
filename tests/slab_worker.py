"""Worker of the Z-slab tests: one process per rank.  Builds the same seeded volume on every rank,
runs skan_b200.slab on this rank's slab and gathers every rank's share on rank 0, which assembles
the global result and compares it with the CPU oracle run on the whole volume."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def make_volume(kind, shape, seed):
    from skan_b200.synth import walks
    if kind == 'walks':
        return walks(shape, max(8, shape[0] * shape[1] // 96), 3 * max(shape), seed=seed, isolated=12, rings=10)
    if kind == 'dense':
        rng = np.random.default_rng(seed)
        return rng.random(shape) < 0.045
    if kind == 'dense12':  # junction clusters that percolate through the volume
        rng = np.random.default_rng(seed)
        return rng.random(shape) < 0.12
    if kind == 'lines_z':   # long straight chains crossing every slab boundary, plus a plane-parallel ring
        img = np.zeros(shape, bool)
        img[:, ::5, ::7] = True
        img[1:-1, 2, 3] = False
        return img
    if kind == 'rings':     # junction-free cycles that cross slab boundaries, touch them, or lie inside a slab
        Z, Y, X = shape
        img = np.zeros(shape, bool)
        zA, zB, xA, xB = 2, Z - 3, 2, X - 3                 # one ring through every slab (x-z plane, corners cut)
        img[zA, 3, xA + 1:xB] = True
        img[zB, 3, xA + 1:xB] = True
        img[zA + 1:zB, 3, xA] = True
        img[zA + 1:zB, 3, xB] = True
        for zc in range(4, Z - 4, 5):                       # 4-voxel rings, some of them astride a boundary
            img[zc - 1, 8, 6] = img[zc, 8, 5] = img[zc, 8, 7] = img[zc + 1, 8, 6] = True
        for zc in range(3, Z - 3, 7):                       # rings lying in one plane (some in a boundary plane)
            img[zc, 12, 11:14] = img[zc, 16, 11:14] = True
            img[zc, 13:16, 10] = img[zc, 13:16, 14] = True
        img[1:-1, 19, 3] = img[2:-2, 19, 9] = True          # and ordinary paths through every slab
        img[Z // 2, 19, 3:10] = True
        return img
    raise ValueError(kind)


def assemble(parts, world):
    """Global arrays from the per-rank shares (sorted by rank)."""
    out = {}
    out['indptr'] = np.concatenate([p['indptr'][:-1] for p in parts] + [parts[-1]['indptr'][-1:]])
    for k in ('indices', 'data', 'coords', 'degrees'):
        out[k] = np.concatenate([p[k] for p in parts])
    out['pidx'] = sum(p['pidx'].astype(np.int64) for p in parts).astype(np.int32)
    out['pdata'] = sum(p['pdata'] for p in parts)
    out['pw'] = sum(p['pw'] for p in parts)
    # global path order: regular paths rank-major, then cycles rank-major
    def cat(key, axis=0):
        reg = [p[key][..., :p['n_regular']] for p in parts]
        cyc = [p[key][..., p['n_regular']:] for p in parts]
        return np.concatenate(reg + cyc, axis=-1)
    for k in ('plen', 'dist', 'mean', 'stdev', 'o32', 'o64', 'of'):
        out[k] = cat(k)
    return out


def compare_with_oracle(img, spacing, glob, height=False):
    from oracle import skan_oracle as orc
    ref = orc.OracleSkeleton(img, spacing=spacing, value_is_height=height)
    rs = orc.summarize(ref, value_is_height=bool(height))
    assert np.array_equal(glob['indptr'], ref.graph.indptr), 'graph.indptr'
    assert np.array_equal(glob['indices'], ref.graph.indices), 'graph.indices'
    assert np.array_equal(glob['data'], ref.graph.data), 'graph.data'
    assert np.array_equal(glob['coords'], ref.coordinates), 'coordinates'
    assert np.array_equal(glob['degrees'], ref.degrees), 'degrees'
    pindptr = np.concatenate([[0], np.cumsum(glob['plen'])])
    assert np.array_equal(pindptr, ref.paths.indptr), 'paths.indptr'
    assert np.array_equal(glob['pidx'], ref.paths.indices), 'paths.indices'
    np.testing.assert_allclose(glob['pdata'], ref.paths.data, rtol=1e-12)
    d = img.ndim
    cols = {'skeleton_id': glob['o32'][0], 'node_id_src': glob['o32'][1], 'node_id_dst': glob['o32'][2],
            'branch_distance': glob['dist'], 'branch_type': glob['o64'][0], 'mean_pixel_value': glob['mean'],
            'stdev_pixel_value': glob['stdev']}
    dh = d + int(bool(height))   # value_is_height: the pixel value is one more coordinate (csr.py:932-948)
    for i in range(d):
        cols[f'image_coord_src_{i}'] = glob['o64'][1 + i]
        cols[f'image_coord_dst_{i}'] = glob['o64'][1 + d + i]
    for i in range(dh):
        cols[f'coord_src_{i}'] = glob['of'][i]
        cols[f'coord_dst_{i}'] = glob['of'][dh + i]
    cols['euclidean_distance'] = glob['of'][2 * dh]
    for k, want in rs.items():
        got = cols[k]
        want = np.asarray(want)
        if want.dtype.kind in 'iub':
            assert np.array_equal(got, want), k
        elif k.startswith('stdev'):
            m = np.abs(rs['mean_pixel_value'])
            assert np.all(np.abs(got**2 - want**2) <= 1e-12 * np.maximum(m * m, 1e-300) + 1e-12 * want**2), k
        else:
            np.testing.assert_allclose(got, want, rtol=1e-12, err_msg=k)
    return ref


def run(rank, world, port, kind, shape, seed, spacing, halo, valued, backend, emulate, native=1, small=0, xhalo=0, prime=0, guard=0,
        height=0):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        import skan_b200
        from skan_b200 import slab
        slab.NATIVE = bool(native)   # 1: skb_run_slab (one native call per rank), 0: the per-stage python sequencing
        if small:   # tiny first arenas: every rank has to retry, at different points (restart protocol)
            slab._INITIAL_ARENAS, slab._INITIAL_EXCHANGE = (4096, 4096), 4096
        img = make_volume(kind, shape, seed)
        if valued:
            rng = np.random.default_rng(seed + 1)
            img = (img * (rng.random(shape) + 0.5)).astype(np.float32)
        z0, z1, ze0, ze1 = slab.slab_bounds(shape[0], world, rank, halo)
        part = np.ascontiguousarray(img[z0:z1] if xhalo else img[ze0:ze1])
        if emulate:
            from emul_util import emulated_device
            cm = emulated_device()
            dev = torch.device('cpu')
        else:
            import contextlib
            cm = contextlib.nullcontext()
            torch.cuda.set_device(rank % torch.cuda.device_count())
            dev = torch.device('cuda', rank % torch.cuda.device_count())
        with cm:
            if guard:   # guard gaps behind every allocation of the runner, arenas pre-filled with a pattern (tests/guard_util.py)
                from guard_util import guarded
                gctx = guarded(0xA5)
                gctx.__enter__()
            # prime: earlier calls on this communicator leave capacity hints behind (1: the same volume, so the hints fit;
            # 2: a much sparser volume first, so the hints are too small and the speculative stages must go again)
            for p in range(prime and 1):
                pimg = img if prime == 1 else (img & (np.random.default_rng(99).random(shape) < 0.3)
                                               if img.dtype == bool else img * (np.random.default_rng(99).random(shape) < 0.3))
                ppart = np.ascontiguousarray(pimg[z0:z1] if xhalo else pimg[ze0:ze1])
                try:
                    if xhalo:
                        slab.skeleton_and_summary_slab(torch.from_numpy(ppart).to(dev), global_shape=shape, z0=z0, z1=z1,
                                                       spacing=spacing, device=dev, halo='exchange')
                    else:
                        slab.skeleton_and_summary_slab(torch.from_numpy(ppart).to(dev), global_shape=shape, z0=z0, z1=z1,
                                                       ze0=ze0, spacing=spacing, device=dev)
                except ValueError:
                    pass   # a volume without any path: every rank raises alike
            if xhalo:   # disjoint partition: the ranks exchange the packed mask of their boundary planes
                res = slab.skeleton_and_summary_slab(torch.from_numpy(part).to(dev), global_shape=shape, z0=z0, z1=z1,
                                                     spacing=spacing, device=dev, halo='exchange',
                                                     value_is_height=bool(height))
            else:
                res = slab.skeleton_and_summary_slab(torch.from_numpy(part).to(dev), global_shape=shape, z0=z0, z1=z1,
                                                     ze0=ze0, spacing=spacing, device=dev, value_is_height=bool(height))
            if guard:
                if dev.type == 'cuda':
                    torch.cuda.synchronize()
                assert gctx.check_gaps() > 10
                gctx.__exit__(None, None, None)
            indptr, indices, data, coords, degrees = slab.owned_graph(res)
            share = dict(indptr=indptr, indices=indices, data=data, coords=coords, degrees=degrees,
                         pidx=res.paths_indices, pdata=res.paths_data, pw=res.paths_weights, plen=res.path_len,
                         dist=res.dist, mean=res.mean, stdev=res.stdev, o32=res.table[0], o64=res.table[1],
                         of=res.table[2])
            share = {k: v.cpu().numpy() for k, v in share.items()}
            share['n_regular'] = res.n_regular
            share['meta'] = dict(rank_rounds=res.rank_rounds, table_rounds=res.table_rounds, n_table=res.n_table,
                                 n_cycles=res.n_cycles, native=hasattr(res, 'counts'))
            from skan_b200 import digest
            dg = digest.digest_slab(res)     # collective: the same on every rank
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(share, gathered, dst=0)
        if rank == 0:
            glob = assemble(gathered, world)
            ref = compare_with_oracle(img, spacing, glob, height=bool(height))
            from oracle import skan_oracle as orc
            want = digest.digest_oracle(ref, orc.summarize(ref, value_is_height=bool(height)))
            assert digest.same(dg, want), f'digest of the slab run differs from the oracle\'s: {dg} != {want}'
            print(f'SLAB-OK world={world} kind={kind} shape={shape} nodes={ref.graph.shape[0]} paths={ref.n_paths} '
                  f'cycles={sum(g["meta"]["n_cycles"] for g in gathered)} table={gathered[0]["meta"]["n_table"]} '
                  f'table_rounds={gathered[0]["meta"]["table_rounds"]} native={int(gathered[0]["meta"]["native"])} xhalo={int(xhalo)}', flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--kind', default='walks')
    ap.add_argument('--shape', default='24,20,28')
    ap.add_argument('--seed', type=int, default=1)
    ap.add_argument('--halo', type=int, default=4)
    ap.add_argument('--valued', type=int, default=0)
    ap.add_argument('--backend', default='gloo')
    ap.add_argument('--emulate', type=int, default=1)
    ap.add_argument('--spacing', default='0.5,0.1,0.1')
    ap.add_argument('--native', type=int, default=1)
    ap.add_argument('--small-arenas', type=int, default=0)
    ap.add_argument('--halo-exchange', type=int, default=0)
    ap.add_argument('--prime', type=int, default=0)
    ap.add_argument('--guard', type=int, default=0)
    ap.add_argument('--height', type=int, default=0, help='value_is_height (needs --valued 1)')
    a = ap.parse_args()
    shape = tuple(int(x) for x in a.shape.split(','))
    spacing = tuple(float(x) for x in a.spacing.split(','))
    run(int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['MASTER_PORT']), a.kind, shape, a.seed,
        spacing, a.halo, bool(a.valued), a.backend, bool(a.emulate), a.native, a.small_arenas, a.halo_exchange, a.prime, a.guard,
        a.height)
