"""GPU parity of the 768-d head (SURVEY 8d cfg 3 in small; north_star "64-/768-dim").  Needs a B200: ``-m gpu``."""
import numpy as np
import pytest
import torch

from helpers import TOL_F16, TOL_FP32
from scenedino_b200 import _abi, ops
from test_gpu_parity import _check_pass, dev, g2n, scenes_from_golden

pytestmark = pytest.mark.gpu


def test_render_d768_vs_reference(golden):
    """fp32 path: the 768-d head (d_out = 769) with four colour views, coarse pass and the 96-sample fine pass on the
    reference's own merged depths, against the reference's outputs."""
    g = golden("render_d768")
    _, dsc, _, dmlp = scenes_from_golden(g)
    rays = g["rays"][0]
    z = ops.sample_coarse(dev(rays), dev(g["u_coarse"]), dev(g["lin"]), True)
    assert np.array_equal(g2n(z), g["coarse.z_samps"][0])
    for p in ("coarse.", "fine."):
        o = ops.render_pass(dsc, dmlp, dev(rays), dev(g[p + "z_samps"][0]), hard_alpha_cap=True, precision=ops.FP32)
        assert o["dino_features"].shape == (48, 768)
        _check_pass(o, g, p, TOL_FP32)


def test_render_d768_tensor_cores(golden):
    """The same fixture on the tensor-core path (sd_render_pass, SD_MLP_F16_TC, projected scene): ONE fused field kernel
    that composites the 128 hidden units per ray on the tensor cores + the head2 kernel that applies the 768 feature rows
    of W_out to the per-ray sums.  No per-sample feature ever reaches memory: the scratch is 516 B per RAY."""
    g = golden("render_d768")
    _, dsc, _, dmlp = scenes_from_golden(g, feat_dtype=torch.float16)
    dsc = dsc.project(dmlp)
    rays = g["rays"][0]
    for p in ("coarse.", "fine."):
        z = dev(g[p + "z_samps"][0])
        R, K = z.shape
        sc, m = dsc.c(), dmlp.c(ops.F16)
        import ctypes as C
        need = _abi.lib().sd_render_workspace_bytes(C.byref(sc), C.byref(m), R, K)
        assert 0 < need <= R * 129 * 4 + 512, "scratch is per ray, not per sample"
        n0 = _abi.launch_count()
        o = ops.render_pass(dsc, dmlp, dev(rays), z, hard_alpha_cap=True, precision=ops.F16)
        assert _abi.launch_count() - n0 == 2, "field kernel + head2 kernel"
        assert o["dino_features"].shape == (48, 768)
        _check_pass(o, g, p, TOL_F16)


def test_render_d64_hidden_composite_matches_feature_composite(golden):
    """D = 64 through both tensor-core composites: the default one sums the 64 features per ray, the hidden-composite one
    (what D > 64 uses) sums the 128 hidden units and applies W_out afterwards.  The second is exercised through a
    129-output copy of the head whose rows 65.. repeat rows 1..64, so both must agree."""
    g = golden("render_coarse")
    _, dsc, _, dmlp = scenes_from_golden(g, feat_dtype=torch.float16)
    w_out = np.concatenate([g["w_out"], g["w_out"][1:]], 0)            # [129, 128]
    b_out = np.concatenate([g["b_out"], g["b_out"][1:]], 0)
    wide = ops.Mlp(g["w_in"], g["b_in"], w_out, b_out, device="cuda")
    rays, z = dev(g["rays"][0]), dev(g["coarse.z_samps"][0])
    a = ops.render_pass(dsc.project(dmlp), dmlp, rays, z, precision=ops.F16)
    b = ops.render_pass(dsc.project(wide), wide, rays, z, precision=ops.F16)
    assert b["dino_features"].shape[1] == 128
    assert torch.equal(b["dino_features"][:, :64], b["dino_features"][:, 64:])
    for k in ("invalid", "invalid_features"):
        assert torch.equal(a[k], b[k]), k
    for k in ("weights", "alphas", "depth"):                           # the per-sample path is the same arithmetic
        assert float((a[k] - b[k]).abs().max()) <= 1e-5 * max(1.0, float(a[k].abs().max())), k
    _check_pass({**b, "dino_features": b["dino_features"][:, :64]}, g, "coarse.", TOL_F16)
    scale = float(a["dino_features"].abs().mean())
    assert float((a["dino_features"] - b["dino_features"][:, :64]).abs().max()) < 2e-2 * scale * 10


@pytest.mark.parametrize("K,D,nv", [(64, 200, 2), (96, 768, 4), (32, 64, 1)])
def test_render_big_heads_many_tiles_vs_oracle(K, D, nv):
    """Several tiles per CTA (the fixtures above give every CTA at most one): 2 600 rays of a rotated view against the
    oracle, D > 64 on the hidden-composite path incl. a ragged last block of W_out (D = 200), D = 64 on the feature composite."""
    from oracle import oracle as O
    from scenedino_b200 import synthetic as syn
    from helpers import assert_close
    C_, Hf, Wf = 256, 48, 160
    feat = syn.make_feature_map(5, C_, Hf, Wf)
    imgs = syn.make_images(6, nv, 24, 80)
    Km = np.broadcast_to(syn.kitti360_K(), (nv, 3, 3)).copy()
    c2w = np.stack([syn.view_pose_c2w(v) for v in range(nv)])
    w2c = np.linalg.inv(c2w.astype(np.float64)).astype(np.float32)
    mlp_w = syn.make_mlp(2, d_out=D + 1, bias_scale=0.05)
    R = 2600
    rays = syn.image_rays(syn.view_pose_c2w(3), Km[0])[:: (syn.IMG_H * syn.IMG_W) // R][:R]
    rs = np.random.RandomState(K)
    z = np.sort(rs.uniform(3, 80, (R, K)).astype(np.float32), 1)
    osc = O.Scene(feat=feat, K_f=Km[:1], w2c_f=w2c[:1], rgb=imgs, K_c=Km, w2c_c=w2c)
    oo = O.render_pass(osc, O.Mlp(*mlp_w), rays, z, hard_alpha_cap=True)
    dsc = ops.Scene.from_arrays(feat, Km[:1], w2c[:1], imgs, Km, w2c, device="cuda", feat_dtype=torch.float16)
    dmlp = ops.Mlp(*mlp_w, device="cuda")
    dsc = dsc.project(dmlp)
    o = ops.render_pass(dsc, dmlp, dev(rays), dev(z), hard_alpha_cap=True, precision=ops.F16)
    for k in ("weights", "alphas", "depth", "dino_features", "rgb"):
        assert_close(g2n(o[k]), oo[k], TOL_F16, k)
    assert np.array_equal(g2n(o["invalid"]), oo["invalid"]) and np.array_equal(g2n(o["invalid_features"]), oo["invalid_features"])
    # depth keeps a hi/lo pair through the fp16 composite operands: far tighter than the bar
    assert_close(g2n(o["depth"]), oo["depth"], 2e-3, "depth (hi/lo split)")
    o2 = ops.render_pass(dsc, dmlp, dev(rays), dev(z), hard_alpha_cap=True, precision=ops.F16, per_sample=False)
    for k in ("depth", "dino_features", "rgb"):
        assert torch.equal(o2[k], o[k]), k
