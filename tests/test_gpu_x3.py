"""SD_MLP_F32_TC: the projected-map tile kernel with every operand as an fp16 (hi, lo) pair and three tensor-core products
per contraction (field_bin_x3.cu) against the reference's golden outputs and the oracle at the fp32 bar (rel 1e-4).
Needs a B200: run with ``-m gpu``."""
import numpy as np
import pytest
import torch

from helpers import TOL_FP32, assert_close, big_query_points
from oracle import oracle as O
from scenedino_b200 import _abi, ops
from scenedino_b200 import synthetic as syn
from test_gpu_parity import DEV, dev, g2n, scenes_from_golden

pytestmark = pytest.mark.gpu


def test_projection_x3_matches_fp32_matmul(golden):
    g = golden("query")
    _, dsc, _, dmlp = scenes_from_golden(g, learn_empty=True, empty_feature=g["empty_feature"])
    p = dsc.project_x3(dmlp)
    Hf, Wf, C_ = dsc.feat.shape
    nb = Hf * Wf * 256
    blob = p.proj_x3
    off = 32768 + 2 * 2 * 80 * 128
    hi = blob[off:off + nb].view(torch.float16).view(Hf * Wf, 128).double()
    lo = blob[off + nb:off + 2 * nb].view(torch.float16).view(Hf * Wf, 128).double()
    want = dsc.feat.view(Hf * Wf, C_).double() @ torch.from_numpy(g["w_in"][:, :C_]).to(DEV).double().T
    err = ((hi + lo) - want).abs().max().item()
    assert err <= 2e-6 * want.abs().max().item(), err


@pytest.mark.parametrize("tag,learn_empty", [("", False), ("_le", True)])
def test_tile_kernel_x3_vs_reference_big(golden, tag, learn_empty):
    """The reference's own outputs on the 70 001-point query (fixture query_big): masks bit-exact, densities / features of
    the stored subset within 1e-4 -- on the tensor cores; launches: 3 sort + tile kernel."""
    g = golden("query_big")
    kw = dict(learn_empty=learn_empty, empty_feature=g["empty_feature"] if learn_empty else None)
    _, dsc, _, dmlp = scenes_from_golden(g, **kw)
    dscp = dsc.project_x3(dmlp)
    pts, sub = big_query_points(g)
    n0 = _abi.launch_count()
    q = ops.query_points(dscp, dmlp, dev(pts), want_rgb=False, precision=ops.F32TC)
    assert _abi.launch_count() - n0 == 4
    inv = np.unpackbits(g["invalid_features" + tag])[:len(pts)].astype(bool)
    assert np.array_equal(g2n(q["invalid_features"]), inv)
    assert_close(g2n(q["sigma"])[sub], g["sigma" + tag], TOL_FP32, "sigma vs reference")
    assert_close(g2n(q["dino"])[sub], g["dino" + tag], TOL_FP32, "dino vs reference")
    # the fp32 CUDA-core kernel on the same points: the two rel-1e-4 modes agree
    q32 = ops.query_points(dsc, dmlp, dev(pts), want_rgb=False, precision=ops.FP32)
    assert_close(g2n(q["sigma"]), g2n(q32["sigma"]), TOL_FP32, "x3 vs fp32 kernel, sigma")
    assert_close(g2n(q["dino"]), g2n(q32["dino"]), TOL_FP32, "x3 vs fp32 kernel, dino")


@pytest.mark.parametrize("map_hw", [(192, 640), (384, 1280)])
def test_ssc_grid_full_size_x3(map_hw):
    """configs[1] at the fp32 bar on the tensor cores: masks bit-exact on all 2 097 152 voxels, values against the oracle on
    a strided subset, binned rows equal caller-order rows bit for bit, degenerate (behind-camera) voxels finite."""
    Hf, Wf = map_hw
    feat = syn.make_feature_map(1, 256, Hf, Wf)
    K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
    mlp_w = syn.make_mlp(0, bias_scale=0.05)
    pts = syn.ssc_voxel_grid()
    dsc = ops.Scene.from_arrays(feat, K, w2c, device=DEV)
    dmlp = ops.Mlp(*mlp_w, device=DEV)
    dscp = dsc.project_x3(dmlp)
    dp = dev(pts)
    q = ops.query_points(dscp, dmlp, dp, want_rgb=False, precision=ops.F32TC)
    _, _, oinv = O.project(K[0], w2c[0], pts)
    assert np.array_equal(g2n(q["invalid_features"]), oinv)
    sub = np.arange(0, len(pts), 257)
    o = O.query_points(O.Scene(feat=feat, K_f=K, w2c_f=w2c), O.Mlp(*mlp_w), pts[sub])
    assert_close(g2n(q["sigma"])[sub], o["sigma"], TOL_FP32, "sigma")
    assert_close(g2n(q["dino"])[sub], o["dino"], TOL_FP32, "dino")
    assert torch.isfinite(q["sigma"]).all() and torch.isfinite(q["dino"]).all()
    b = ops.query_points_binned(dscp, dmlp, dp, precision=ops.F32TC)
    assert torch.equal(b["sigma"], q["sigma"]) and torch.equal(b["dino_binned"], q["dino"][b["perm"].long()])


def test_btsnet_forward_fp32_tc(golden):
    """BTSNet.forward with sd_precision = "fp32_tc": the reference-shaped surface on the x3 tile kernel; against the fp32
    CUDA-core precision of the same net at 1e-4, masks bit for bit."""
    import bench
    import scenedino_b200 as sd
    hold = {"map": dev(syn.make_feature_map(1, 256, 96, 320))}
    net = bench.build_net(sd, torch, hold, DEV, "fp32_tc", with_head=False)
    imgs = dev(syn.make_images(2, 1))[None]
    Kc = dev(syn.kitti360_K()[None])[None]
    c2w = dev(np.eye(4, dtype=np.float32)[None])[None]
    net.encode(imgs * 2 - 1, Kc, c2w, ids_encoder=[0], ids_render=[0], images_alt=imgs)
    net.set_scale(0)
    grid = dev(syn.ssc_voxel_grid()[::3].copy())[None]
    with torch.no_grad():
        _, inv_a, sig_a, _, st_a = net(grid, only_density=True)
        net.precision = "fp32"
        _, inv_b, sig_b, _, st_b = net(grid, only_density=True)
    assert torch.equal(inv_a, inv_b)
    assert_close(g2n(sig_a), g2n(sig_b), TOL_FP32, "sigma, fp32_tc vs fp32")
    assert_close(g2n(st_a["dino_features"]), g2n(st_b["dino_features"]), TOL_FP32, "dino, fp32_tc vs fp32")
