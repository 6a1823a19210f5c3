"""sd_query_points_binned: the tile kernel's 64-d rows left in texel-bin order by TMA tile stores (+ the permutation that
goes with them) against the caller-order query -- bit for bit -- and, fed to the fused SSC head with that permutation,
against the head on caller-order rows.  Needs a B200: run with ``-m gpu``."""
import numpy as np
import pytest
import torch

from helpers import big_query_points
from scenedino_b200 import SdError, ops
from scenedino_b200 import synthetic as syn
from test_gpu_parity import DEV, dev, scenes_from_golden

pytestmark = pytest.mark.gpu


def _check_binned(dsc, dmlp, dp):
    q = ops.query_points(dsc, dmlp, dp, precision=ops.F16, want_rgb=False)
    b = ops.query_points_binned(dsc, dmlp, dp)
    N = dp.shape[0]
    perm = b["perm"].long()
    assert torch.equal(torch.sort(perm).values, torch.arange(N, device=DEV)), "perm is a permutation of the points"
    assert torch.equal(b["sigma"], q["sigma"]) and torch.equal(b["invalid_features"], q["invalid_features"])
    assert torch.equal(b["dino_binned"], q["dino"][perm]), "row r of dino_binned is the feature row of point perm[r]"
    return q, b


@pytest.mark.parametrize("learn_empty", [False, True])
def test_binned_rows_equal_caller_order_rows(golden, learn_empty):
    """70 001 points: a ragged last tile (rows past N are clipped by the copy engine); with learn_empty the rows outside
    the frustum (the projected empty feature instead of taps) leave by the same stores."""
    g = golden("query_big")
    kw = dict(learn_empty=learn_empty, empty_feature=g["empty_feature"] if learn_empty else None)
    _, dsc, _, dmlp = scenes_from_golden(g, feat_dtype=torch.float16, **kw)
    dsc = dsc.project(dmlp)
    pts, _ = big_query_points(g)
    q, _ = _check_binned(dsc, dmlp, dev(pts))
    assert 0 < q["invalid_features"].float().mean() < 1


@pytest.mark.parametrize("map_hw", [(192, 640), (384, 1280)])
def test_binned_ssc_grid_full_size_and_reuse(map_hw):
    """All 2 097 152 voxels of configs[1], both map sizes; then the query again on a NEW map with the sort reused
    (reuse_sorted: what the per-frame SSC loop does) against a fresh caller-order query of that map."""
    Hf, Wf = map_hw
    K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
    mlp_w = syn.make_mlp(0, bias_scale=0.05)
    dmlp = ops.Mlp(*mlp_w, device=DEV)
    dp = dev(syn.ssc_voxel_grid())
    dsc = ops.Scene.from_arrays(syn.make_feature_map(1, 256, Hf, Wf), K, w2c, device=DEV, feat_dtype=torch.float16).project(dmlp)
    _, b = _check_binned(dsc, dmlp, dp)
    dsc2 = ops.Scene.from_arrays(syn.make_feature_map(2, 256, Hf, Wf), K, w2c, device=DEV, feat_dtype=torch.float16).project(dmlp)
    ws_out = {}                                  # the workspace travels in the dict the first call fills
    first = ops.query_points_binned(dsc, dmlp, dp)
    ws_out.update(sigma=first["sigma"], dino_binned=first["dino_binned"], perm=first["perm"],
                  invalid_features=first["invalid_features"].view(torch.uint8))
    ops.query_points_binned(dsc, dmlp, dp, out=ws_out)                       # leaves _workspace in ws_out
    keep_perm = ws_out["perm"].clone()           # (the order inside a bin differs from sort to sort: atomics)
    again = ops.query_points_binned(dsc2, dmlp, dp, out=ws_out, reuse_sorted=True)
    q2 = ops.query_points(dsc2, dmlp, dp, precision=ops.F16, want_rgb=False)
    assert torch.equal(again["perm"], keep_perm)
    assert torch.equal(again["sigma"], q2["sigma"]) and torch.equal(again["dino_binned"], q2["dino"][keep_perm.long()])
    assert not torch.equal(q2["sigma"], b["sigma"])


def test_binned_rows_feed_the_ssc_head():
    """The consumer: sd_ssc_head on binned rows with perm writes every label where the head on caller-order rows does."""
    K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
    dmlp = ops.Mlp(*syn.make_mlp(0, bias_scale=0.05), device=DEV)
    dsc = ops.Scene.from_arrays(syn.make_feature_map(1, 256, 96, 320), K, w2c, device=DEV, feat_dtype=torch.float16).project(dmlp)
    dp = dev(syn.ssc_voxel_grid()[::7].copy())
    q, b = _check_binned(dsc, dmlp, dp)
    head = ops.SscHead(syn.make_expand(3), syn.make_ssc_head(21), device=DEV)
    want = ops.ssc_head(head, q["dino"], want_scores=True)
    got = ops.ssc_head(head, b["dino_binned"], want_scores=True, perm=b["perm"])
    assert torch.equal(got["seg"], want["seg"]) and torch.equal(got["pseudo"], want["pseudo"])
    assert torch.equal(got["scores"], want["scores"])


def test_binned_errors_are_loud(golden):
    g = golden("query")
    _, dsc, _, dmlp = scenes_from_golden(g, feat_dtype=torch.float16)
    pts = dev(g["points"])
    with pytest.raises(SdError, match="projected"):
        ops.query_points_binned(dsc, dmlp, pts)                              # no projection
    few = ops.query_points_binned  # too few points for the tile path: refused, never a silent other path
    dscp = dsc.project(dmlp)
    with pytest.raises(SdError):
        few(dscp, dmlp, pts[:5].contiguous())


def test_btsnet_forward_segmentation_takes_the_binned_path():
    """BTSNet.forward(grid, predict_segmentation=True) (models/bts.py:584-592): with the fused head the 64-d rows stay in
    texel-bin order between the query and sd_ssc_head (launches: 3 sort + tile kernel + head = 5, then tile + head = 2 on
    the next frame of a static grid); labels / sigma equal the caller-order route (materialize_dino_full) wherever the
    head's top-2 gap is clear -- here: bit for bit, both routes feed the head the same rows."""
    import bench
    import scenedino_b200 as sd
    from scenedino_b200 import _abi
    hold = {"map": dev(syn.make_feature_map(1, 256, 96, 320))}
    net = bench.build_net(sd, torch, hold, DEV, "fp16")
    imgs = dev(syn.make_images(2, 1))[None]
    Kc = dev(syn.kitti360_K()[None])[None]
    c2w = dev(np.eye(4, dtype=np.float32)[None])[None]
    net.encode(imgs * 2 - 1, Kc, c2w, ids_encoder=[0], ids_render=[0], images_alt=imgs)
    net.set_scale(0)
    grid = dev(syn.ssc_voxel_grid()[::5].copy())[None]
    net.static_query, net.one_hot_seg = True, False
    with torch.no_grad():
        net.materialize_dino_full = True
        full, _, sig_a, seg_a = net(grid, predict_segmentation=True)
        net._static_cache.clear()
        net.materialize_dino_full = False
        n0 = _abi.launch_count()
        none, _, sig_b, seg_b = net(grid, predict_segmentation=True)
        n1 = _abi.launch_count()
        _, _, sig_c, seg_c = net(grid, predict_segmentation=True)          # static grid: the sort is reused
        n2 = _abi.launch_count()
    assert none is None and full is not None
    assert n1 - n0 == 5 and n2 - n1 == 2, (n1 - n0, n2 - n1)
    assert torch.equal(sig_a, sig_b) and torch.equal(sig_b, sig_c)
    assert torch.equal(seg_a.reshape(-1).to(torch.int64), seg_b.reshape(-1).to(torch.int64)) and torch.equal(seg_b, seg_c)
