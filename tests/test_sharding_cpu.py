"""world_size-2 (and 3) gloo tests of the multi-GPU plumbing: shard bounds, ragged all-gather, and the
ray / voxel sharding wrappers with a stand-in for the per-rank compute.  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scenedino_b200 import sharding as sh


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 256, 122880, 2097152):
        for w in (1, 2, 3, 4, 8):
            b = sh.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_bounds(4, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(rays):
    """Deterministic per-ray function standing in for the renderer (no GPU here)."""
    d = rays[..., 3:6].sum(-1)
    return {"coarse": {"rgb": torch.stack([d, d * 2, d * 3], -1), "depth": d + 1, "dino_features": rays[..., :4] * 0.5,
                       "weights": rays[..., :2]}}


def _fake_query(xyz):
    return {"sigma": xyz.sum(-1), "dino": xyz.repeat(1, 2)}


def _worker(rank, world, port, n_rays, grid, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        rays = torch.rand(1, n_rays, 11, generator=g)
        out = sh.render_rays_sharded(_fake_render, rays)
        full = _fake_render(rays)
        ok = all(torch.equal(out["coarse"][k], full["coarse"][k]) for k in ("rgb", "depth", "dino_features"))
        ok = ok and "weights" not in out["coarse"]
        X, Y, Z = grid
        xyz = torch.rand(X * Y * Z, 3, generator=g)
        o = sh.query_voxels_sharded(_fake_query, xyz, grid)
        ok = ok and torch.equal(o["sigma"], _fake_query(xyz)["sigma"])
        s, e = o["x_range"]
        ok = ok and torch.equal(o["local"]["dino"], _fake_query(xyz)["dino"][s * Y * Z: e * Y * Z])
        # ragged gather of an odd split along a middle dimension
        t = torch.arange(2 * 7 * 3, dtype=torch.float32).reshape(2, 7, 3)
        sl = sh.shard_slice(7, rank, world)
        ok = ok and torch.equal(sh.all_gather_ragged(t[:, sl].contiguous(), 7, dim=1), t)
        # NeRFRenderer.bind_parallel(gpus=[...]) under a process group: the reference's DataParallel(dim=1) wrapper
        # (renderer/nerf.py:654-658) as one process per GPU -- each rank renders its tile, per-ray outputs are gathered
        import scenedino_b200 as sd
        ren = sd.NeRFRenderer(n_coarse=4)
        wrapped = ren.bind_parallel(torch.nn.Identity(), gpus=list(range(world)))
        assert type(wrapped).__name__ == "_ShardedRenderWrapper" and wrapped.renderer is ren
        wrapped.module.forward = lambda r, **kw: {**_fake_render(r), "state_dict": {"x": r[..., 0]}}
        got = wrapped(rays, want_weights=True)
        ok = ok and all(torch.equal(got["coarse"][k], full["coarse"][k]) for k in ("rgb", "depth", "dino_features", "weights"))
        ok = ok and got["state_dict"]["x"].shape[1] == sh.shard_slice(n_rays, rank, world).stop - sh.shard_slice(n_rays, rank, world).start
        try:
            ren.bind_parallel(torch.nn.Identity(), gpus=list(range(world + 1)))
            ok = False
        except NotImplementedError:
            pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rays,grid", [(2, 101, (5, 4, 3)), (3, 64, (8, 2, 2))])
def test_sharded_wrappers_gloo(world, n_rays, grid):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rays, grid, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def test_voxel_slabs_of_the_grid_concatenate_to_the_whole_grid():
    """Voxel-slab sharding (one x-slab of the SSC grid per rank): the slabs every rank builds for itself -- host restatement
    and C oracle alike -- laid end to end are the whole grid, for even and ragged splits."""
    from oracle import oracle as O
    from scenedino_b200 import synthetic as syn
    dims = (24, 8, 4)
    T = syn.velo_to_cam()
    whole = syn.ssc_voxel_grid(dims=dims)
    assert np.array_equal(whole, O.voxel_grid(T, dims=dims))
    for w in (1, 2, 3, 5, 8):
        slabs = sh.shard_bounds(dims[0], w)
        assert np.array_equal(np.concatenate([syn.ssc_voxel_grid(dims=dims, x_range=b) for b in slabs]), whole)
        assert np.array_equal(np.concatenate([O.voxel_grid(T, dims=dims, x_range=b) for b in slabs]), whole)
