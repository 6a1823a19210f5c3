"""Degenerate inputs of the texel sort in front of the tile kernel (csrc/binning.cu): every point in ONE bin, every point
behind the camera, a handful of occupied bins out of thousands, point counts around the range / tile sizes, and a feature map
with more 7x7-texel bins than the sort holds in shared memory (it then sorts by wider bins and the gather kernel walks that
order).  The reference treats every point independently (models/bts.py:476-595), so whatever the order inside the kernels,
every point's result must be the oracle's.  Needs a B200: run with ``-m gpu``."""
import numpy as np
import pytest
import torch

from helpers import TOL_F16, assert_close
from oracle import oracle as O
from scenedino_b200 import _abi, ops
from scenedino_b200 import synthetic as syn
from test_gpu_parity import DEV, dev, g2n

pytestmark = pytest.mark.gpu


def _scene(Hf, Wf, seed=31, C=256):
    feat = syn.make_feature_map(seed, C, Hf, Wf)
    K = syn.kitti360_K()[None]
    w2c = np.eye(4, dtype=np.float32)[None]
    mlp_w = syn.make_mlp(5, 295, 128, 65, bias_scale=0.1)
    osc = O.Scene(feat=feat, K_f=K, w2c_f=w2c)
    dsc = ops.Scene.from_arrays(feat, K, w2c, device=DEV, feat_dtype=torch.float16)
    return osc, dsc, O.Mlp(*mlp_w), ops.Mlp(*mlp_w, device=DEV)


def _check(osc, dsc, omlp, dmlp, pts, launches, sample=4096):
    """device query of all points; oracle on a seeded subset (the whole set when small)"""
    n0 = _abi.launch_count()
    q = ops.query_points(dsc, dmlp, dev(pts), precision=ops.F16, want_rgb=False)
    assert _abi.launch_count() - n0 == launches
    N = len(pts)
    idx = np.arange(N) if N <= sample else np.sort(np.random.RandomState(3).choice(N, sample, replace=False))
    o = O.query_points(osc, omlp, pts[idx], want_rgb=False)
    assert np.array_equal(g2n(q["invalid_features"])[idx], o["invalid_features"])
    assert_close(g2n(q["sigma"])[idx], o["sigma"], TOL_F16, "sigma")
    assert_close(g2n(q["dino"])[idx], o["dino"], TOL_F16, "dino")
    assert torch.isfinite(q["sigma"]).all() and torch.isfinite(q["dino"]).all()
    return q


def _cases(n):
    rs = np.random.RandomState(17)
    one = np.tile(np.array([[1.5, 0.4, 12.0]], np.float32), (n, 1))                                  # one texel, one bin
    behind = (rs.uniform(-1, 1, (n, 3)) * np.array([30, 6, 40]) - np.array([0, 0, 45])).astype(np.float32)   # z < 0: all masked
    few = np.array([[-3.0, 0.2, 8.0], [2.0, -0.5, 20.0], [0.1, 0.8, 50.0]], np.float32)[rs.randint(0, 3, n)]
    few = (few + rs.uniform(-0.02, 0.02, (n, 3))).astype(np.float32)                                 # three clusters
    ramp = np.stack([np.linspace(-20, 20, n), np.full(n, 0.3), np.linspace(3, 60, n)], 1).astype(np.float32)   # sorted walk
    return {"one_bin": one, "behind_camera": behind, "three_clusters": few, "ordered_walk": ramp}


@pytest.mark.parametrize("case", ["one_bin", "behind_camera", "three_clusters", "ordered_walk"])
def test_degenerate_point_distributions(case):
    """70 001 points on the sorted tile path (texel sort = 3 launches + tile kernel): a single bin holding every point, the
    border bins holding every point (all behind the camera: every row takes the masked path), three occupied bins, and
    points already in bin order."""
    osc, dsc, omlp, dmlp = _scene(50, 300)
    dsc = dsc.project(dmlp)
    pts = _cases(70001)[case]
    q = _check(osc, dsc, omlp, dmlp, pts, launches=4)
    if case == "behind_camera":
        assert bool(q["invalid_features"].all())
    if case == "one_bin":
        assert torch.equal(q["sigma"], q["sigma"][:1].expand_as(q["sigma"])), "identical points, identical results"
        b = ops.query_points_binned(dsc, dmlp, dev(pts))
        assert torch.equal(torch.sort(b["perm"].long()).values, torch.arange(len(pts), device=DEV))
        assert torch.equal(b["dino_binned"], q["dino"][b["perm"].long()])


@pytest.mark.parametrize("n", [65535, 65536, 70000, 151553, 303104, 303105])
def test_point_counts_around_range_and_tile_sizes(n):
    """The sort splits the points into at most 2 x (number of SMs) contiguous ranges of whole 1024-point blocks, the tile
    kernel into 128-row tiles: one point less than the smallest query that is sorted at all (sd_query_workspace_bytes: 65 536;
    below it the gather kernel runs alone), that smallest query, a ragged last block, exactly 296 full blocks, and one point
    more (ranges of unequal length)."""
    osc, dsc, omlp, dmlp = _scene(50, 300)
    dsc = dsc.project(dmlp)
    pts = syn.random_points(n, n)
    _check(osc, dsc, omlp, dmlp, pts, launches=4 if n >= 65536 else 1, sample=2048)


def test_map_with_more_bins_than_the_sort_holds():
    """832 x 832 texels = 119 x 119 bins of 7 x 7 > 12 288: the sort falls back to 14 x 14-texel bins, leaves no tile
    records, and the gather kernel walks the sorted order (2 sort launches + field_tc_kernel), projected scene or not."""
    osc, dsc, omlp, dmlp = _scene(832, 832, C=256)
    pts = syn.random_points(41, 230001, box=((-12, 12), (-8, 8), (2, 40)))   # >= 16 points per 7 x 7 bin: the tile path is asked for
    _check(osc, dsc, omlp, dmlp, pts, launches=3, sample=2048)
    _check(osc, dsc.project(dmlp), omlp, dmlp, pts, launches=3, sample=2048)
