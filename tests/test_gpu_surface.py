"""The reference-shaped Python surface (BTSNet / NeRFRenderer / _RenderWrapper) against the golden
dictionaries the unmodified reference produced: same keys, same shapes, values within tolerance,
random draws injected in the reference's call order.  Needs a B200: ``-m gpu``."""
import contextlib

import numpy as np
import pytest
import torch

import scenedino_b200 as sd
from helpers import TOL_F16, TOL_FP32, assert_close, golden_scene_arrays

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


class FakeEncoder(torch.nn.Module):
    """Seeded map standing in for the DINO ViT + DPT encoder (cf. EncoderDummy,
    training/trainer_overfit.py:21-30 in the reference)."""

    def __init__(self, feat):
        super().__init__()
        self.latent_size, self.extra_outs = feat.shape[1], 0
        self.dim_reduction = sd.MlpDimReduction(768, 64, 128)
        self.register_buffer("feat", feat)

    def forward(self, x, ground_truth=False):
        if ground_truth:
            return [torch.zeros(x.shape[0], 8, 2, 2, device=x.device)]
        return [self.feat.clone()]

    def expand_dim(self, f):
        return self.dim_reduction.transform_expand(f)


def build(g, n=1, nv=None, learn_empty=False, precision="fp32"):
    feat, imgs = golden_scene_arrays(g, n=n)
    conf = {"predict_dino": True, "dino_dims": 64, "inv_z": True, "learn_empty": learn_empty, "code_mode": "z",
            "sd_precision": precision}
    code = sd.PositionalEncoding.from_conf({"num_freqs": 6, "freq_factor": 1.5, "include_input": True}, d_in=3)
    enc = FakeEncoder(torch.from_numpy(feat))
    head = sd.make_head({"type": "resnet", "name": "normal_head", "args": {"n_blocks": 0, "d_hidden": 128}},
                        enc.latent_size + code.d_out, 65)
    net = sd.BTSNet(conf, enc, code, {"normal_head": head}, None)
    # weights arrive the way a reference checkpoint would: by state-dict key
    state = {"heads.normal_head.lin_in.weight": g["w_in"], "heads.normal_head.lin_in.bias": g["b_in"],
             "heads.normal_head.lin_out.weight": g["w_out"], "heads.normal_head.lin_out.bias": g["b_out"]}
    if "e_w1" in g:
        state.update({"encoder.dim_reduction.linear_in.weight": g["e_w1"], "encoder.dim_reduction.linear_in.bias": g["e_b1"],
                      "encoder.dim_reduction.linear_out.weight": g["e_w2"], "encoder.dim_reduction.linear_out.bias": g["e_b2"]})
    if learn_empty:
        state["empty_feature"] = g["empty_feature"]
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=False)
    assert not unexpected
    net = net.to(DEV).eval()
    K = g["K"] if n > 1 else g["K"][None]
    c2w = g["c2w"] if n > 1 else g["c2w"][None]
    nv = K.shape[1]
    net.encode(torch.zeros(n, nv, 3, 8, 8, device=DEV), dev(K), dev(c2w), ids_encoder=[0],
               ids_render=list(range(nv)), images_alt=dev(imgs))
    net.set_scale(0)
    return net


@contextlib.contextmanager
def injected_draws(draws):
    """Replays the reference's recorded torch.rand / rand_like / randn_like results in call order."""
    q = [dev(d) for d in draws]
    orig = (torch.rand, torch.rand_like, torch.randn_like)

    def pop(*a, **k):
        return q.pop(0)

    torch.rand = torch.rand_like = torch.randn_like = pop
    try:
        yield
    finally:
        torch.rand, torch.rand_like, torch.randn_like = orig
    assert not q, "the renderer made fewer random draws than the reference"


def check_dict(out, g, prefix, tol):
    keys = [k[len(prefix):] for k in g if k.startswith(prefix)]
    assert sorted(out.keys()) == sorted(keys)
    for k in keys:
        a, b = out[k].detach().cpu().numpy(), g[prefix + k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert a.dtype == b.dtype, (k, a.dtype, b.dtype)
        if k in ("invalid", "invalid_features", "ray_info"):
            assert np.array_equal(a, b), k
        elif k == "rgb_samps":          # bit-equal for unrotated views; torch's CPU bmm rounds rotated ones differently
            assert_close(a, b, TOL_FP32, prefix + k)
        elif k == "z_samps":
            assert np.array_equal(a, b) or np.mean(np.all(a == b, -1)) > 0.99, k
        else:
            assert_close(a, b, tol, prefix + k)


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("fp16", TOL_F16)])
def test_render_wrapper_coarse(golden, precision, tol):
    g = golden("render_coarse")
    net = build(g, precision=precision)
    ren = sd.NeRFRenderer.from_conf({"n_coarse": 64, "n_fine": 0, "lindisp": True, "eval_batch_size": 4096})
    ren.hard_alpha_cap = False                      # demo_utils/utils.py:45
    wrapped = ren.bind_parallel(net, gpus=None).eval()
    with torch.no_grad(), injected_draws([g["u_coarse"]]):
        out = wrapped(dev(g["rays"]), want_weights=True, want_alphas=True, want_z_samps=True, want_rgb_samps=True)
    assert isinstance(out, dict) and set(out.keys()) == {"coarse", "state_dict"}
    check_dict(out["coarse"], g, "coarse.", tol)
    assert set(out["state_dict"].keys()) == {"invalid_features", "dino_features"}
    assert_close(out["state_dict"]["dino_features"].cpu().numpy(), g["state_dict.dino_features"], tol, "state dino")
    # simple_output
    with torch.no_grad(), injected_draws([g["u_coarse"]]):
        rgb, depth = ren.bind_parallel(net, simple_output=True)(dev(g["rays"]))
    assert_close(depth.cpu().numpy(), g["coarse.depth"], tol, "simple depth")
    # empty-ray shortcut (nerf.py:28-32)
    r0, d0 = wrapped(torch.zeros(0, 5, 11, device=DEV))
    assert r0.shape == (0, 3) and d0.shape == (0,)


@pytest.mark.parametrize("name", ["render_fine", "render_fine_lin"])
def test_render_wrapper_fine(golden, name):
    g = golden(name)
    Kc, Kf, Kfd, lindisp, white = [int(v) for v in g["conf"]]
    net = build(g)
    ren = sd.NeRFRenderer.from_conf({"n_coarse": Kc, "n_fine": Kf, "n_fine_depth": Kfd, "lindisp": bool(lindisp),
                                     "depth_std": float(g["depth_std"]), "white_bkgd": bool(white),
                                     "hard_alpha_cap": True})
    wrapped = ren.bind_parallel(net).eval()
    with torch.no_grad(), injected_draws([g["u_coarse"], g["u_fine0"], g["u_fine1"], g["n_depth"]]):
        out = wrapped(dev(g["rays"]), want_weights=True, want_alphas=True, want_z_samps=True, want_rgb_samps=True)
    assert set(out.keys()) == {"coarse", "fine", "state_dict"}
    check_dict(out["coarse"], g, "coarse.", TOL_FP32)
    # fine pass: rows whose importance indices did not flip at a 1-ulp CDF tie must agree
    a, b = out["fine"]["z_samps"].cpu().numpy(), g["fine.z_samps"]
    # the fine samples sit on top of the coarse pass's weights / depth (tolerance-level quantities), so
    # they agree to rounding, except where an importance index flipped at a CDF tie
    same = np.all(np.abs(a - b) <= 1e-5 * np.abs(b), -1)[0]
    assert same.mean() > 0.99
    for k in ("weights", "alphas", "depth", "rgb", "dino_features"):
        assert_close(out["fine"][k].cpu().numpy()[0][same], g["fine." + k][0][same], 3 * TOL_FP32, "fine." + k)


def test_render_wrapper_from_dist(golden):
    g = golden("render_from_dist")
    net = build(g)
    ren = sd.NeRFRenderer.from_conf({"n_coarse": 40, "n_fine": 0, "lindisp": True})
    with torch.no_grad(), injected_draws([g["u0"], g["u1"]]):
        out = ren.bind_parallel(net).eval()(dev(g["rays"]), want_weights=True, want_alphas=True, want_z_samps=True,
                                            want_rgb_samps=True, sample_from_dist=(dev(g["prop_weights"]), dev(g["prop_z"])))
    check_dict(out["coarse"], g, "coarse.", TOL_FP32)


def test_render_wrapper_superbatch(golden):
    g = golden("render_superbatch")
    net = build(g, n=2)
    ren = sd.NeRFRenderer.from_conf({"n_coarse": 32, "n_fine": 0, "lindisp": True, "hard_alpha_cap": True})
    with torch.no_grad(), injected_draws([g["u_coarse"]]):
        out = ren.bind_parallel(net).eval()(dev(g["rays"]), want_weights=True, want_alphas=True, want_z_samps=True,
                                            want_rgb_samps=True)
    check_dict(out["coarse"], g, "coarse.", TOL_FP32)


@pytest.mark.parametrize("tag,learn_empty", [("", False), ("_le", True)])
def test_btsnet_forward(golden, tag, learn_empty):
    g = golden("query")
    net = build(g, learn_empty=learn_empty)
    xyz = dev(g["points"])[None]
    with torch.no_grad():
        rgb, invalid, sigma, extras, state = net(xyz)
        sf, sinv = net.sample_features(xyz)
        col, cinv = net.sample_colors(xyz)
        dino_full, inv2, sigma2, seg = net(xyz[:, :256], predict_segmentation=True)
        rgb0, inv0, sigma0, _, _ = net(xyz, only_density=True)
    N = xyz.shape[1]
    assert extras is None and seg is None and inv2 is None
    assert rgb.shape == (1, N, 6) and invalid.shape == (1, N, 2) and sigma.shape == (1, N, 1)
    assert invalid.dtype == torch.float32 and state["invalid_features"].dtype == torch.bool
    assert state["invalid_features"].shape == (1, N, 1) and state["dino_features"].shape == (1, N, 64)
    assert_close(sigma[0, :, 0].cpu().numpy(), g["sigma" + tag], TOL_FP32, "sigma")
    assert_close(state["dino_features"][0].cpu().numpy(), g["dino" + tag], TOL_FP32, "dino")
    assert np.array_equal(rgb[0].cpu().numpy(), g["rgb" + tag])
    assert np.array_equal(invalid[0].cpu().numpy(), g["invalid" + tag])
    assert np.array_equal(state["invalid_features"][0, :, 0].cpu().numpy(), g["invalid_features" + tag])
    assert sf.shape == (1, N, 1, 295) and sinv.shape == (1, N, 1) and sinv.dtype == torch.bool
    n = g["sample_features" + tag].shape[0]
    assert np.array_equal(sf[0, :n, 0, :256].cpu().numpy(), g["sample_features" + tag][:, :256])
    assert col.shape == (1, 2, N, 3) and cinv.shape == (1, 2, N, 1)
    assert np.array_equal(col.permute(0, 2, 1, 3).reshape(N, 6).cpu().numpy(), g["rgb" + tag])
    assert_close(dino_full[0].cpu().numpy(), g["dino_full" + tag], TOL_FP32, "dino_full")
    assert_close(sigma2[0, :, 0].cpu().numpy(), g["sigma_seg" + tag], TOL_FP32, "sigma (segmentation call)")
    assert rgb0.shape == (1, N, 6) and float(rgb0.abs().sum()) == 0.0
    assert np.array_equal(inv0[0, :, 0].cpu().numpy(), g["invalid_features" + tag].astype(np.float32))


def test_btsnet_forward_large_fp16_uses_projected_map(golden):
    """A big reduced-precision point query (the SSC voxel chunks of sscbench/evaluate_model_sscbench.py:711-717) runs
    on the projected map made once per encode: texel sort + tile kernel, same masks as fp32, values within 2e-2."""
    from scenedino_b200 import _abi
    from scenedino_b200 import synthetic as syn
    g = golden("query")
    net = build(g, learn_empty=True)
    pts = np.concatenate([g["points"], syn.random_points(3, 70000)]).astype(np.float32)
    xyz = dev(pts)[None]
    with torch.no_grad():
        net.precision = "fp32"
        _, _, sigma32, _, st32 = net(xyz, only_density=True)
        net.precision = "fp16"
        _, _, sigma16, _, st16 = net(xyz, only_density=True)
        n0 = _abi.launch_count()
        dino_full, _, sigma_seg, _ = net(xyz, predict_segmentation=True)
        second = _abi.launch_count() - n0
    projs = [st["proj"] for st in net._packed.values() if "proj" in st]
    assert len(projs) == 1 and len(projs[0]) == 1, "one projection per encode, batch element and head"
    assert 4 <= second <= 7, "sort (3 launches) + tile kernel (+ expand_dim), no second projection"
    assert torch.equal(st16["invalid_features"], st32["invalid_features"])
    assert_close(sigma16[0, :, 0].cpu().numpy(), sigma32[0, :, 0].cpu().numpy(), TOL_F16, "sigma fp16 vs fp32")
    assert_close(st16["dino_features"][0].cpu().numpy(), st32["dino_features"][0].cpu().numpy(), TOL_F16, "dino fp16 vs fp32")
    n = len(g["points"])
    assert_close(sigma16[0, :n, 0].cpu().numpy(), g["sigma_le"], TOL_F16, "sigma vs reference")
    assert dino_full.shape == (1, xyz.shape[1], 768) and sigma_seg.shape == (1, xyz.shape[1], 1)
    # the expansion follows the query's precision (tensor-core expand kernel): two reduced-precision stages in a row
    assert_close(dino_full[0, :256].cpu().numpy(), g["dino_full_le"], 5e-2, "dino_full (fp16 query + fp16 expansion) vs reference")
    np.testing.assert_allclose(torch.linalg.norm(dino_full[0], dim=-1).cpu().numpy(), 1.0, rtol=1e-5)


def test_image_ray_sampler_matches_reference(golden):
    """ImageRaySampler.sample (ray_sampler.py:439-513): rays bit-equal to the reference's for a batch of two elements
    (one kernel launch for the whole batch), ground-truth colours / features in the reference's row order, and the
    attribute side effects (channels, height, width learnt from the images)."""
    from scenedino_b200 import synthetic as syn
    g = golden("rays")
    for tag in "ab":
        H, W = (int(x) for x in g[f"{tag}_hw"])
        c2w, proj = g[f"{tag}_c2w"], g[f"{tag}_proj"]
        n, v = c2w.shape[:2]
        imgs = np.stack([syn.make_images(12 + i, v, H, W) for i in range(n)], 0)
        ids = [int(x) for x in g[f"{tag}_ids"]] or None
        s = sd.ImageRaySampler(3.0, 80.0, norm_dir=bool(g[f"{tag}_norm_dir"]), channels=1)
        n0 = sd.launch_count()
        rays, rgb_gt = s.sample(dev(imgs), dev(c2w), dev(proj), image_ids=ids)
        assert sd.launch_count() - n0 == 1
        assert (s.height, s.width, s.channels) == (H, W, 3)
        assert rays.shape == (n, v * H * W, 11) and np.array_equal(rays.cpu().numpy(), g[f"{tag}_rays"])
        want_gt = imgs.transpose(0, 1, 3, 4, 2).reshape(n, v * H * W, 3)
        assert np.array_equal(rgb_gt.cpu().numpy(), want_gt)
        feats = torch.rand(n, v, 16, H // 4, W // 4, device=DEV)
        r2, gt2, dino_gt = s.sample(dev(imgs), dev(c2w), dev(proj), image_ids=ids, dino_features=feats)
        assert torch.equal(r2, rays) and dino_gt.shape == (n, v * (H // 4) * (W // 4), 16)
        assert torch.equal(dino_gt[1, (H // 4) * (W // 4) + 3] if n > 1 else dino_gt[0, 3],
                           feats[1, 1, :, 0, 3] if n > 1 else feats[0, 0, :, 0, 3])


def test_sampler_renderer_reconstruct_round_trip(golden):
    """The demo's call sequence (demo_utils/utils.py:223-231): sample whole views -> render -> reconstruct.  Rendering
    the sampler's rays in one call equals rendering them row by row, and reconstruct lays the result out as images."""
    g = golden("render_coarse")
    net = build(g, precision="fp16")
    ren = sd.NeRFRenderer.from_conf({"n_coarse": 16, "n_fine": 0, "lindisp": True, "eval_batch_size": 1 << 16})
    ren.hard_alpha_cap = False
    wrapped = ren.bind_parallel(net, gpus=None).eval()
    H, W = 12, 40
    s = sd.ImageRaySampler(3.0, 80.0, H, W)
    rays, _ = s.sample(None, dev(g["c2w"][None, :1]), dev(g["K"][None, :1]))
    torch.manual_seed(0)
    u = torch.rand(H * W, 16, device=DEV)
    with torch.no_grad(), injected_draws([u.cpu().numpy()]):
        out = wrapped(rays, want_weights=True, want_alphas=True, want_z_samps=True)
    with torch.no_grad(), injected_draws([u[:W].cpu().numpy()]):
        row0 = wrapped(rays[:, :W], want_weights=True)
    nv = g["K"].shape[0]
    assert nv == 2
    # reference quirk kept: _format_outputs folds invalid_features to [n, R / nv_c, K, nv_c] (nerf.py:586) and reconstruct
    # views it as [n, v, H, W, K, nv_c] (ray_sampler.py:543-547), which only fits with ONE rendered colour view
    with pytest.raises(RuntimeError):
        s.reconstruct({"coarse": dict(out["coarse"])})
    coarse = {k: v for k, v in out["coarse"].items() if k != "invalid_features"}
    img = s.reconstruct({"coarse": coarse, "state_dict": out["state_dict"]})
    assert img["coarse"]["depth"].shape == (1, 1, H, W) and img["coarse"]["rgb"].shape == (1, 1, H, W, nv, 3)
    assert img["coarse"]["dino_features"].shape == (1, 1, H, W, 1, 64) and img["coarse"]["weights"].shape == (1, 1, H, W, 16)
    assert torch.equal(img["coarse"]["depth"][0, 0, 0], row0["coarse"]["depth"][0])
    assert torch.isfinite(img["coarse"]["depth"]).all() and (img["coarse"]["depth"] >= 0).all()


def test_no_silent_fallback(golden):
    g = golden("query")
    net = build(g)
    with pytest.raises(sd.SdError):
        net(torch.zeros(1, 4, 3))                     # CPU tensor
    net.train()                                       # training + autograd: the differentiable (unfused) route, SURVEY 8f-4
    pts = dev(g["points"][:64].reshape(1, -1, 3))
    rgb, invalid, sigma, _, state = net(pts)
    assert sigma.requires_grad and state["dino_features"].requires_grad and not rgb.requires_grad
    net.eval()
    with torch.no_grad():
        rgb_e, invalid_e, sigma_e, _, state_e = net(pts)
    assert torch.equal(rgb, rgb_e) and torch.equal(invalid, invalid_e)
    assert_close(sigma.detach().cpu().numpy(), sigma_e.cpu().numpy(), TOL_FP32, "training-route sigma vs fused kernel")
    assert_close(state["dino_features"].detach().cpu().numpy(), state_e["dino_features"].cpu().numpy(), TOL_FP32, "training-route dino")
    net.train()
    with pytest.raises(NotImplementedError):
        net(pts, predict_segmentation=True)           # the SSC head is evaluation-only


def test_launch_stream_follows_torch_current_stream():
    """Every C-ABI launch goes to torch's CURRENT stream of the tensors' device (heads._stream hands the raw handle over):
    the default stream outside, a side stream inside ``torch.cuda.stream`` -- and work queued there is ordered with it."""
    from scenedino_b200 import heads
    assert (heads._stream().value or 0) == torch.cuda.current_stream().cuda_stream
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        assert (heads._stream().value or 0) == side.cuda_stream
        x = torch.randn(1000, 3, device=DEV)
        code = sd.PositionalEncoding(num_freqs=6, d_in=3, freq_factor=1.5, include_input=True)(x)
        ev = torch.cuda.Event(); ev.record()
    assert (heads._stream().value or 0) == torch.cuda.current_stream().cuda_stream
    ev.synchronize()
    assert code.shape == (1000, 39) and torch.equal(code[:, :3], x)


def test_encode_pose_inverse_replayed_from_a_graph_is_bit_identical(golden):
    """BTSNet.encode inverts the poses with torch's own solver (bts.py:125-126); from the second encode with the same batch
    shape on, that call is replayed from a CUDA graph: the same kernels, so the same bits as the eager call -- for new pose
    values, after a change of shape, and with the switch off."""
    g = golden("query")
    net = build(g)                                           # encode no. 1 (eager)
    nv = g["K"].shape[0]
    K = dev(g["K"][None]); imgs = torch.zeros(1, nv, 3, 8, 8, device=DEV)
    rs = np.random.RandomState(5)

    def pose_batch(n_views):
        out = []
        for _ in range(n_views):
            q, _ = np.linalg.qr(rs.randn(3, 3))
            m = np.eye(4, dtype=np.float32); m[:3, :3] = q.astype(np.float32); m[:3, 3] = rs.uniform(-3, 3, 3)
            out.append(m)
        return dev(np.stack(out)[None])

    for i in range(4):                                       # no. 2 captures, 3.. replay
        c2w = pose_batch(nv)
        net.encode(imgs, K, c2w, ids_encoder=[0], ids_render=list(range(nv)))
        want = torch.linalg.inv_ex(c2w).inverse
        assert torch.equal(net.grid_c_poses_w2c, want) and torch.equal(net.grid_f_poses_w2c, want[:, [0]]), f"encode {i + 2}"
    kinds = {type(v).__name__ for v in net._pose_inverse._by_shape.values()}
    assert kinds == {"tuple"}, f"the solver path was not captured: {net._pose_inverse._by_shape}"
    c2w1 = pose_batch(1)                                     # another shape: eager first, its own graph afterwards
    for _ in range(3):
        net.encode(imgs[:, :1], K[:, :1], c2w1, ids_encoder=[0], ids_render=[0])
        assert torch.equal(net.grid_c_poses_w2c, torch.linalg.inv_ex(c2w1).inverse)
    net.graph_pose_inverse = False
    net.encode(imgs[:, :1], K[:, :1], c2w1, ids_encoder=[0], ids_render=[0])
    assert torch.equal(net.grid_c_poses_w2c, torch.linalg.inv_ex(c2w1).inverse)
