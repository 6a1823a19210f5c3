"""the backward pass of the training step.  The two custom gradient kernels against torch autograd on a plain
PyTorch restatement of the same stage (nerf.py:376-421, bts.py:299-319), and a whole training-mode render through
BTSNet + NeRFRenderer whose gradients (encoder map, head weights, empty feature) match the restatement's.
Needs a B200: run with ``-m gpu``."""
import numpy as np
import pytest
import torch

from scenedino_b200 import ops
from scenedino_b200 import synthetic as syn
from scenedino_b200.autograd import CompositeFn
from test_gpu_parity import DEV, dev

pytestmark = pytest.mark.gpu


def torch_composite(z, sigma, feat, rgb, hard_alpha_cap, white_bkgd):
    """renderer/nerf.py:366-421 in plain torch (fp64 for the comparison)."""
    deltas = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)], -1)
    alphas = 1 - torch.exp(-deltas.abs() * torch.relu(sigma))
    if hard_alpha_cap:
        alphas = torch.cat([alphas[:, :-1], torch.ones_like(alphas[:, :1])], -1)
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], -1)
    T = torch.cumprod(shifted, -1)
    weights = alphas * T[:, :-1]
    depth = (weights * z).sum(-1)
    dino = (weights.unsqueeze(-1) * feat).sum(-2)
    rgb_out = (weights.unsqueeze(-1) * rgb).sum(-2)
    if white_bkgd:
        rgb_out = rgb_out + 1 - weights.sum(1, keepdim=True)
    return weights, alphas, depth, dino, rgb_out


@pytest.mark.parametrize("R,K,D,Crgb", [(257, 64, 64, 3), (33, 96, 8, 12), (19, 1, 5, 3), (64, 37, 100, 6), (5, 130, 64, 3), (3, 200, 16, 3)])
@pytest.mark.parametrize("cap,white", [(False, False), (True, True)])
def test_composite_backward_vs_autograd(R, K, D, Crgb, cap, white):
    g = torch.Generator(device=DEV).manual_seed(R * 1000 + K)
    z = torch.sort(torch.rand((R, K), device=DEV, generator=g) * 50 + 3, dim=1).values
    sigma = (torch.randn((R, K), device=DEV, generator=g) * 0.3).requires_grad_()      # about half negative: relu'd away
    feat = torch.randn((R, K, D), device=DEV, generator=g).requires_grad_()
    rgb = torch.rand((R, K, Crgb), device=DEV, generator=g).requires_grad_()
    outs = CompositeFn.apply(z, sigma, feat, rgb, cap, white)
    ups = [torch.randn(o.shape, device=DEV, generator=g) for o in outs]
    torch.autograd.backward(outs, ups)
    got = [sigma.grad.clone(), feat.grad.clone(), rgb.grad.clone()]
    s64, f64, c64 = (t.detach().double().requires_grad_() for t in (sigma, feat, rgb))
    ref = torch_composite(z.double(), s64, f64, c64, cap, white)
    for o, r in zip(outs, ref):
        assert torch.allclose(o.double(), r, rtol=2e-5, atol=2e-6)
    torch.autograd.backward(ref, [u.double() for u in ups])
    for name, a, b in zip(("sigma", "feat", "rgb"), got, (s64.grad, f64.grad, c64.grad)):
        scale = float(b.abs().max()) + 1e-30
        err = float((a.double() - b).abs().max()) / scale
        assert err < 5e-5, (name, err)


def _net(learn_empty):
    import bench
    import scenedino_b200 as sd
    hold = {}
    net = bench.build_net(sd, torch, hold, DEV, "fp32", with_head=False)
    net.learn_empty = learn_empty
    if learn_empty:
        net.empty_feature = torch.nn.Parameter(torch.randn(256, device=DEV) * 0.3)
    return net, hold


def torch_reference_loss(fmap, empty, head, xyz, K, rgb_w, code_fn, learn_empty):
    """bts.py:271-328 + resnetfc.py:162-199 in plain torch on the same points: F.grid_sample of the map, positional code,
    head, softplus; returns (sigma, dino)."""
    N = xyz.shape[0]
    cam = xyz                                   # identity pose
    z = cam[:, 2:3]
    xy = cam[:, :2] / z.clamp_min(1e-3)
    xy = xy * torch.stack([K[0, 0], K[1, 1]]) + torch.stack([K[0, 2], K[1, 2]])
    invalid = (z[:, 0] <= 1e-3) | (xy[:, 0] < -1) | (xy[:, 0] > 1) | (xy[:, 1] < -1) | (xy[:, 1] > 1)
    samp = torch.nn.functional.grid_sample(fmap[None], xy.clamp(-2, 2)[None, :, None, :], mode="bilinear", padding_mode="border",
                                           align_corners=False)[0, :, :, 0].T      # [N, C]
    if learn_empty:
        samp = torch.where(invalid[:, None], empty[None].expand(N, -1), samp)
    return samp, invalid


def test_training_render_gradients_match_torch(golden):
    """One training-mode render through scenedino_b200.BTSNet + NeRFRenderer (autograd on): the gradient of a scalar loss
    with respect to the encoder map, the head weights and the learned empty feature against the same loss built from
    plain-torch stages (grid_sample + nn.Linear + the torch composite above) on the SAME sample depths."""
    import scenedino_b200 as sd
    for learn_empty in (False, True):
        net, hold = _net(learn_empty)
        fmap = (dev(syn.make_feature_map(3, 256, 24, 80)) * 1.0).requires_grad_()      # [1, C, Hf, Wf]
        hold["map"] = fmap
        imgs = dev(syn.make_images(2, 1))[None]
        Kc = dev(syn.kitti360_K()[None])[None]
        c2w = dev(np.eye(4, dtype=np.float32)[None])[None]
        net.train()
        net.encode(imgs * 2 - 1, Kc, c2w, ids_encoder=[0], ids_render=[0], images_alt=imgs)
        net.set_scale(0)
        ren = sd.NeRFRenderer.from_conf({"n_coarse": 24, "n_fine": 0, "lindisp": True, "hard_alpha_cap": True})
        ren.nan_check = False
        ren.train()
        sampler = sd.ImageRaySampler(z_near=syn.Z_NEAR, z_far=syn.Z_FAR, height=syn.IMG_H, width=syn.IMG_W)
        side = dev(syn.view_pose_c2w(2).astype(np.float32)[None])[None]          # a shifted view: part of the samples leave the frustum
        rays, _ = sampler.sample(None, side, Kc)
        rays = rays[:, ::211][:, :300].contiguous()
        out = ren(net, rays, want_weights=True, want_z_samps=True)
        c = out["coarse"]
        gsel = torch.Generator(device=DEV).manual_seed(5)
        wd = torch.randn(c["dino_features"].shape, device=DEV, generator=gsel)
        loss = (c["depth"] * 0.01).sum() + (c["dino_features"] * wd).sum() + c["rgb"].sum() + (c["weights"] ** 2).sum()
        params = [fmap, net.heads["normal_head"].lin_in.weight, net.heads["normal_head"].lin_out.weight,
                  net.heads["normal_head"].lin_in.bias] + ([net.empty_feature] if learn_empty else [])
        got = torch.autograd.grad(loss, params)
        # ---- the same loss from plain-torch stages on the same sample depths --------------------------------------
        z = c["z_samps"][0].detach()
        R, K_ = z.shape
        r0 = rays[0]
        pts = (r0[:, None, :3] + z[..., None] * r0[:, None, 3:6]).reshape(-1, 3)
        with torch.no_grad():
            feat_k, _ = net.sample_features(pts[None])            # the kernel's positional code (no gradient flows through it)
            code = feat_k[0, :, 0, 256:]
            rgb_s, _ = net.sample_colors(pts[None])
        samp, invalid = torch_reference_loss(fmap[0], net.empty_feature if learn_empty else None, None, pts, Kc[0, 0], None, None, learn_empty)
        head = net.heads["normal_head"]
        o = head.lin_out(torch.relu(head.lin_in(torch.cat([samp, code], -1))))
        sigma, dino = torch.nn.functional.softplus(o[:, 0]).reshape(R, K_), o[:, 1:].reshape(R, K_, -1)
        rgbs = rgb_s[0].permute(1, 0, 2).reshape(R, K_, -1)
        w, a, d, f, col = torch_composite(z, sigma, dino, rgbs, True, False)
        loss_ref = (d * 0.01).sum() + (f * wd[0]).sum() + col.sum() + (w ** 2).sum()
        assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 2e-4 * abs(float(loss_ref)), (float(loss), float(loss_ref))
        want = torch.autograd.grad(loss_ref, params)
        for name, a_, b_ in zip(("map", "w_in", "w_out", "b_in", "empty"), got, want):
            scale = float(b_.abs().max()) + 1e-30
            err = float((a_ - b_).abs().max()) / scale
            assert err < 2e-3, (learn_empty, name, err)
            assert float(b_.abs().max()) > 0, name
