"""Differentiable forms of the two custom stages of the training path (training/base_trainer.py:223-255 calls
loss.backward() through renderer and field).  The fused kernels are forward-only; with autograd on, BTSNet / NeRFRenderer
run the path unfused -- feature gather (sd_sample_features), the head as plain torch modules, compositing (sd_composite) --
and the gradients of the two custom stages come from sd_sample_features_bwd / sd_composite_bwd."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from .heads import _f32c, _ptr, _stream, on_device


class CompositeFn(torch.autograd.Function):
    """(z [B,K], sigma [B,K], feat [B,K,D], rgb [B,K,Crgb]) -> (weights, alphas, depth, dino, rgb_out): nerf.py:376-421."""

    @staticmethod
    def forward(ctx, z, sigma, feat, rgb, hard_alpha_cap: bool, white_bkgd: bool):
        z, sigma, feat, rgb = _f32c(z.detach()), _f32c(sigma.detach()), _f32c(feat.detach()), _f32c(rgb.detach())
        B, K = z.shape
        D, Crgb = feat.shape[-1], rgb.shape[-1]
        f32 = dict(dtype=torch.float32, device=z.device)
        weights, alphas = torch.empty((B, K), **f32), torch.empty((B, K), **f32)
        depth, dino, rgb_out = torch.empty((B,), **f32), torch.empty((B, D), **f32), torch.empty((B, Crgb), **f32)
        cfg = _abi.SdRenderCfg()
        cfg.lindisp, cfg.hard_alpha_cap, cfg.white_bkgd = 0, int(bool(hard_alpha_cap)), int(bool(white_bkgd))
        with on_device(z):
            _abi.check(_abi.lib().sd_composite(_ptr(z), _ptr(sigma), _ptr(feat), _ptr(rgb), B, K, D, Crgb, C.byref(cfg),
                                               _ptr(weights), _ptr(alphas), _ptr(depth), _ptr(dino), _ptr(rgb_out), _stream()),
                       "sd_composite")
        ctx.save_for_backward(z, sigma, feat, rgb)
        ctx.cfg = (int(bool(hard_alpha_cap)), int(bool(white_bkgd)))
        return weights, alphas, depth, dino, rgb_out

    @staticmethod
    def backward(ctx, g_weights, g_alphas, g_depth, g_dino, g_rgb_out):
        z, sigma, feat, rgb = ctx.saved_tensors
        B, K = z.shape
        D, Crgb = feat.shape[-1], rgb.shape[-1]
        cfg = _abi.SdRenderCfg()
        cfg.lindisp, cfg.hard_alpha_cap, cfg.white_bkgd = 0, ctx.cfg[0], ctx.cfg[1]
        gs = [None if g is None else _f32c(g) for g in (g_depth, g_dino, g_rgb_out, g_weights, g_alphas)]
        need_sigma, need_feat, need_rgb = ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        g_sigma = torch.empty_like(sigma) if need_sigma else None
        g_feat = torch.empty_like(feat) if need_feat else None
        g_rgb = torch.empty_like(rgb) if need_rgb else None
        with on_device(z):
            _abi.check(_abi.lib().sd_composite_bwd(_ptr(z), _ptr(sigma), _ptr(feat), _ptr(rgb), B, K, D, Crgb, C.byref(cfg),
                                                   *[_ptr(g) for g in gs], _ptr(g_sigma), _ptr(g_feat), _ptr(g_rgb), _stream()),
                       "sd_composite_bwd")
        return None, g_sigma, g_feat, g_rgb, None, None


class SampleFeaturesFn(torch.autograd.Function):
    """(encoder map [C,Hf,Wf], empty_feature [C] or None, xyz [N,3]) -> (features [N, C + code], invalid [N] uint8):
    BTSNet.sample_features (bts.py:271-328) for one batch element.  ``scene`` is the ctypes scene over the packed fp32
    channels-last copy of the same map (BTSNet._state); ``keep`` holds what its pointers refer to."""

    @staticmethod
    def forward(ctx, fmap, empty_feature, xyz, scene, keep, d_in: int):
        xyz = _f32c(xyz.detach())
        N = xyz.shape[0]
        feat = torch.empty((N, d_in), dtype=torch.float32, device=xyz.device)
        inv = torch.empty((N,), dtype=torch.uint8, device=xyz.device)
        with on_device(xyz):
            _abi.check(_abi.lib().sd_sample_features(C.byref(scene), _ptr(xyz), N, _ptr(feat), _ptr(inv), _stream()),
                       "sd_sample_features")
        ctx.save_for_backward(xyz)
        ctx.scene, ctx.keep, ctx.map_shape = scene, keep, tuple(fmap.shape)
        ctx.has_empty = empty_feature is not None
        ctx.mark_non_differentiable(inv)
        return feat, inv

    @staticmethod
    def backward(ctx, g_feat, _g_inv):
        (xyz,) = ctx.saved_tensors
        C_, Hf, Wf = ctx.map_shape
        g_feat = _f32c(g_feat)
        g_map = torch.zeros((Hf, Wf, C_), dtype=torch.float32, device=xyz.device)
        g_empty = torch.zeros((C_,), dtype=torch.float32, device=xyz.device) if ctx.has_empty else None
        with on_device(xyz):
            _abi.check(_abi.lib().sd_sample_features_bwd(C.byref(ctx.scene), _ptr(xyz), xyz.shape[0], _ptr(g_feat), _ptr(g_map),
                                                         _ptr(g_empty), _stream()), "sd_sample_features_bwd")
        return g_map.permute(2, 0, 1), g_empty, None, None, None, None
