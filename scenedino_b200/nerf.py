"""``NeRFRenderer`` -- host-side mirror of scenedino/renderer/nerf.py:12-658.

Same constructor / ``from_conf`` keys, ``sample_*`` / ``composite`` / ``forward`` signatures, output
dictionary keys and shapes, persistent buffers (``iter_idx``, ``last_sched``) and ``bind_parallel``
wrapper.  Random draws are made with the SAME torch calls, in the same order and shapes as the
reference (so a shared seed gives the same samples); all arithmetic runs in libscenedino_b200:

  sample_coarse            -> sd_sample_coarse            (nerf.py:121-141)
  sample_coarse_from_dist  -> sd_sample_coarse_from_dist  (nerf.py:143-179)
  sample_fine              -> sd_sample_fine              (nerf.py:181-212)
  sample_fine_depth        -> sd_sample_fine_depth        (nerf.py:214-228)
  torch.sort of the merge  -> sd_sort_rows                (nerf.py:490,522)
  composite                -> sd_render_pass (native BTSNet: points, field query and compositing in
                              one call) or model(...) + sd_composite (any other model)  (nerf.py:230-449)
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from .heads import _f32c, _ptr, _stream, device_guard, require_cuda


class DotMap(dict):
    """Attribute-access dict with ``toDict`` -- the subset of dotmap.DotMap the reference uses
    (nerf.py:9,499-509,571)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def toDict(self):
        return {k: (v.toDict() if isinstance(v, DotMap) else v) for k, v in self.items()}


class _RenderWrapper(torch.nn.Module):
    """nerf.py:12-53."""

    def __init__(self, net, renderer, simple_output):
        super().__init__()
        self.net = net
        self.renderer = renderer
        self.simple_output = simple_output

    @device_guard
    def forward(self, rays, want_weights=False, want_alphas=False, want_z_samps=False, want_rgb_samps=False,
                sample_from_dist=None):
        if rays.shape[0] == 0:
            return torch.zeros(0, 3, device=rays.device), torch.zeros(0, device=rays.device)
        so = self.simple_output
        outputs = self.renderer(self.net, rays, want_weights=want_weights and not so,
                                want_alphas=want_alphas and not so, want_z_samps=want_z_samps and not so,
                                want_rgb_samps=want_rgb_samps and not so, sample_from_dist=sample_from_dist)
        if so:
            o = outputs.fine if self.renderer.using_fine else outputs.coarse
            return o.rgb, o.depth
        return outputs.toDict()


class NeRFRenderer(torch.nn.Module):
    def __init__(self, n_coarse=128, n_fine=0, n_fine_depth=0, noise_std=0.0, depth_std=0.01,
                 eval_batch_size=100000, white_bkgd=False, lindisp=False, sched=None, hard_alpha_cap=False,
                 render_mode="volumetric", surface_sigmoid_scale=.1, render_flow=False, normalize_dino=False):
        super().__init__()
        self.n_coarse, self.n_fine = n_coarse, n_fine
        self.n_fine_depth = n_fine_depth
        self.noise_std = noise_std
        self.depth_std = depth_std
        self.eval_batch_size = eval_batch_size  # kept for config compatibility; the fused path does not chunk
        self.white_bkgd = white_bkgd
        self.lindisp = lindisp
        self.using_fine = n_fine > 0
        self.sched = sched if (sched is None or len(sched) > 0) else None
        self.register_buffer("iter_idx", torch.tensor(0, dtype=torch.long), persistent=True)
        self.register_buffer("last_sched", torch.tensor(0, dtype=torch.long), persistent=True)
        self.hard_alpha_cap = hard_alpha_cap
        assert render_mode in ("volumetric", "surface", "neus")
        self.render_mode = render_mode
        self.only_surface_color = (self.render_mode == "surface")
        self.surface_sigmoid_scale = surface_sigmoid_scale
        self.render_flow = render_flow
        self.normalize_dino = normalize_dino
        #: the reference scans six tensors for NaN after every composite and exits the process
        #: (nerf.py:428-432); here the per-ray results are scanned and FloatingPointError is raised.
        self.nan_check = True

    # ---- helpers -----------------------------------------------------------------------------------
    @staticmethod
    def _rays2d(rays):
        require_cuda(rays, "rays")
        if rays.dim() != 2 or rays.shape[1] < 8:
            raise ValueError(f"rays must be [B, >=8], got {tuple(rays.shape)}")
        return _f32c(rays)

    def _cfg(self) -> _abi.SdRenderCfg:
        c = _abi.SdRenderCfg()
        c.lindisp, c.hard_alpha_cap, c.white_bkgd = int(bool(self.lindisp)), int(bool(self.hard_alpha_cap)), int(bool(self.white_bkgd))
        return c

    # ---- sampling ------------------------------------------------------------------------------------
    @device_guard
    def sample_coarse(self, rays):
        """nerf.py:121-141: stratified samples, (B, Kc)."""
        rays = self._rays2d(rays)
        B, Kc = rays.shape[0], self.n_coarse
        step = 1.0 / Kc
        lin = torch.linspace(0, 1 - step, Kc, device=rays.device)
        u = torch.rand_like(torch.empty((B, Kc), dtype=torch.float32, device=rays.device))
        z = torch.empty((B, Kc), dtype=torch.float32, device=rays.device)
        _abi.check(_abi.lib().sd_sample_coarse(_ptr(rays), B, rays.shape[1], _ptr(u), _ptr(lin), Kc,
                                               int(bool(self.lindisp)), _ptr(z), _stream()), "sd_sample_coarse")
        return z

    @device_guard
    def sample_coarse_from_dist(self, rays, weights, z_samp):
        """nerf.py:143-179: resampling of a proposal histogram, (B, Kc), unsorted."""
        rays = self._rays2d(rays)
        B, Kc = rays.shape[0], self.n_coarse
        w, zs = _f32c(weights), _f32c(z_samp)
        Kp = w.shape[-1]
        u0 = torch.rand(B, Kc, dtype=torch.float32, device=rays.device)
        u1 = torch.rand_like(u0, dtype=torch.float32)
        z = torch.empty((B, Kc), dtype=torch.float32, device=rays.device)
        _abi.check(_abi.lib().sd_sample_coarse_from_dist(B, _ptr(w), _ptr(zs), Kp, _ptr(u0), _ptr(u1), Kc,
                                                         int(bool(self.lindisp)), _ptr(z), None, _stream()),
                   "sd_sample_coarse_from_dist")
        return z

    @device_guard
    def sample_fine(self, rays, weights):
        """nerf.py:181-212: importance samples, (B, Kf - Kfd)."""
        rays = self._rays2d(rays)
        B = rays.shape[0]
        w = _f32c(weights)
        Kc, Kf = w.shape[-1], self.n_fine - self.n_fine_depth
        u0 = torch.rand(B, Kf, dtype=torch.float32, device=rays.device)
        u1 = torch.rand_like(u0)
        z = torch.empty((B, Kf), dtype=torch.float32, device=rays.device)
        # the reference divides by self.n_coarse (nerf.py:202), which equals the histogram width
        if Kc != self.n_coarse:
            raise ValueError("sample_fine: weights must have n_coarse columns")
        _abi.check(_abi.lib().sd_sample_fine(_ptr(rays), B, rays.shape[1], _ptr(w), Kc, _ptr(u0), _ptr(u1), Kf,
                                             int(bool(self.lindisp)), _ptr(z), None, _stream()), "sd_sample_fine")
        return z

    @device_guard
    def sample_fine_depth(self, rays, depth):
        """nerf.py:214-228: samples around the expected depth, (B, Kfd)."""
        rays = self._rays2d(rays)
        B, Kfd = rays.shape[0], self.n_fine_depth
        d = _f32c(depth)
        noise = torch.randn_like(torch.empty((B, Kfd), dtype=torch.float32, device=rays.device))
        z = torch.empty((B, Kfd), dtype=torch.float32, device=rays.device)
        _abi.check(_abi.lib().sd_sample_fine_depth(_ptr(rays), B, rays.shape[1], _ptr(d), _ptr(noise), Kfd,
                                                   float(self.depth_std), _ptr(z), _stream()), "sd_sample_fine_depth")
        return z

    @staticmethod
    @device_guard
    def _sort_rows(z):
        z = z.contiguous()
        _abi.check(_abi.lib().sd_sort_rows(_ptr(z), z.shape[0], z.shape[1], _stream()), "sd_sort_rows")
        return z

    # ---- composite -----------------------------------------------------------------------------------
    def composite(self, model, rays, z_samp, coarse=True, sb=0):
        """nerf.py:230-449.  Returns the reference's 10-tuple
        (weights, rgb, depth, alphas, invalid, z_samp, rgbs, ray_info, extras, state_dicts)."""
        return self._composite(model, rays, z_samp, coarse, sb, want_rgb_samps=True)

    @device_guard
    def _composite(self, model, rays, z_samp, coarse, sb, want_rgb_samps):
        with torch.profiler.record_function("renderer_composite"):
            if self.render_mode != "volumetric":
                raise NotImplementedError(f"render_mode={self.render_mode!r}: only 'volumetric' is implemented")
            if self.training and self.noise_std > 0.0:
                raise NotImplementedError("noise_std > 0 in training mode is not implemented")
            rays = self._rays2d(rays)
            z_samp = _f32c(z_samp)
            B, K = z_samp.shape
            if (hasattr(model, "_sd_render_pass") or hasattr(model, "_scene")) and not self._wants_grad(model):
                out = self._composite_native(model, rays, z_samp, max(sb, 1), want_rgb_samps)
            else:
                out = self._composite_generic(model, rays, z_samp, coarse, sb)
            weights, rgb_final, depth, alphas, invalid, rgbs, state = out
            if self.nan_check:
                self._nan_check(depth, rgb_final, z_samp)
            ray_info = rays[:, None, 8:] if rays.shape[-1] > 8 else None
            return weights, rgb_final, depth, alphas, invalid, z_samp, rgbs, ray_info, None, state

    @staticmethod
    def _wants_grad(model) -> bool:
        """Training with autograd on: the fused kernels are forward-only, the pass runs unfused (model(points) with a graph,
        then the composite with its custom backward: scenedino_b200.autograd)."""
        return torch.is_grad_enabled() and bool(getattr(model, "training", False))

    @staticmethod
    def _nan_check(*tensors):
        """nerf.py:428-432 scans six tensors and exits the process; here NaN in the per-ray results raises.  One fused
        reduction and ONE host sync for all tensors (a sum is NaN iff an element is -- or +inf meets -inf, equally bad)."""
        total = sum(t.sum(dtype=torch.float32) for t in tensors if t is not None and t.numel())
        if torch.is_tensor(total) and bool(torch.isnan(total)):
            raise FloatingPointError("NaN in rendered depth / rgb / z_samp (reference: nerf.py:428-432)")

    def _native_setup(self, net, sb, B):
        if B_ := B % sb:
            raise ValueError(f"{B} rays do not split into {sb} scenes ({B_} left over)")
        prec = net._precision()
        st = net._state(prec)
        mlp = net._mlp(prec)
        if st["n"] != sb:
            raise ValueError(f"super-batch {sb} but the field was encoded with batch {st['n']}")
        return prec, st, mlp

    def _forward_native(self, net, rays, sb, want_rgb_samps):
        """NeRFRenderer.forward for a native BTSNet: ONE library call per scene (sd_render_rays: coarse sampling, coarse pass,
        importance + depth samples, merge, sort, fine pass launched back to back, no Python in between).  The random draws
        are made here with the reference's torch calls in the reference's order (nerf.py:134, 193-203, 221)."""
        if self.render_mode != "volumetric":
            raise NotImplementedError(f"render_mode={self.render_mode!r}: only 'volumetric' is implemented")
        if self.training and self.noise_std > 0.0:
            raise NotImplementedError("noise_std > 0 in training mode is not implemented")
        rays = self._rays2d(rays)
        B = rays.shape[0]
        prec, st, mlp = self._native_setup(net, max(sb, 1), B)
        sb = max(sb, 1)
        Bp = B // sb
        Kc, Kf, Kfd = self.n_coarse, (self.n_fine if self.using_fine else 0), (self.n_fine_depth if self.using_fine else 0)
        Kfi = Kf - Kfd
        dev = rays.device
        f32 = dict(dtype=torch.float32, device=dev)
        lin = torch.linspace(0, 1 - 1.0 / Kc, Kc, device=dev)
        u_c = torch.rand_like(torch.empty((B, Kc), **f32))
        u0 = torch.rand(B, Kfi, **f32) if Kfi > 0 else None
        u1 = torch.rand_like(u0) if Kfi > 0 else None
        noise = torch.randn_like(torch.empty((B, Kfd), **f32)) if Kfd > 0 else None
        nv_c, D = st["rgb"].shape[1], mlp.d_out - 1

        def alloc(K):
            return dict(weights=torch.empty((B, K), **f32), alphas=torch.empty((B, K), **f32), depth=torch.empty((B,), **f32),
                        dino=torch.empty((B, D), **f32), rgb=torch.empty((B, 3 * nv_c), **f32), z=torch.empty((B, K), **f32),
                        invalid=torch.empty((B, K, nv_c), **f32), invf=torch.empty((B, K, 1), dtype=torch.uint8, device=dev),
                        rgbs=torch.empty((B, K, 3 * nv_c), **f32) if want_rgb_samps else None)

        passes = [alloc(Kc)] + ([alloc(Kc + Kf)] if Kf > 0 else [])
        use_proj = (prec == _abi.SD_MLP_F16_TC and st["C"] == 256 and mlp.d_hidden == 128
                    and (D > 64 or Bp * Kc >= 65536))
        cfg = self._cfg()
        sp = _abi.SdSampling()
        sp.n_coarse, sp.n_fine, sp.n_fine_depth, sp.depth_std = Kc, Kf, Kfd, float(self.depth_std)
        lib = _abi.lib()
        for b in range(sb):
            sc = net._scene(st, b, net._projection(st, b, mlp) if use_proj else None)
            sl = slice(b * Bp, (b + 1) * Bp)
            outs = []
            for o in passes:
                ro = _abi.SdRenderOut()
                ro.depth, ro.dino, ro.rgb = o["depth"][sl].data_ptr(), o["dino"][sl].data_ptr(), o["rgb"][sl].data_ptr()
                ro.weights, ro.alphas, ro.z_samps = o["weights"][sl].data_ptr(), o["alphas"][sl].data_ptr(), o["z"][sl].data_ptr()
                ro.invalid, ro.invalid_feat = o["invalid"][sl].data_ptr(), o["invf"][sl].data_ptr()
                ro.rgb_samps = o["rgbs"][sl].data_ptr() if want_rgb_samps else None
                outs.append(ro)
            need = lib.sd_render_rays_workspace_bytes(C.byref(sc), C.byref(mlp), C.byref(sp), Bp)
            ws = torch.empty((need,), dtype=torch.uint8, device=dev)
            _abi.check(lib.sd_render_rays(
                C.byref(sc), C.byref(mlp), C.byref(cfg), C.byref(sp), _ptr(rays[sl]), Bp, rays.shape[1], _ptr(u_c[sl]), _ptr(lin),
                _ptr(u0[sl]) if Kfi > 0 else None, _ptr(u1[sl]) if Kfi > 0 else None, _ptr(noise[sl]) if Kfd > 0 else None,
                C.byref(outs[0]), C.byref(outs[1]) if Kf > 0 else None, _ptr(ws), need, _stream()), "sd_render_rays")
        ray_info = rays[:, None, 8:] if rays.shape[-1] > 8 else None
        res = []
        for o in passes:
            if self.nan_check:
                self._nan_check(o["depth"], o["rgb"], o["z"])
            state = {"invalid_features": o["invf"].view(torch.bool), "dino_features": o["dino"]}
            res.append((o["weights"], o["rgb"], o["depth"], o["alphas"], o["invalid"], o["z"], o["rgbs"], ray_info, None, state))
        return res

    def _composite_native(self, net, rays, z, sb, want_rgb_samps):
        prec, st, mlp = self._native_setup(net, sb, rays.shape[0])
        B, K = z.shape
        Bp = B // sb
        nv_c, D = st["rgb"].shape[1], mlp.d_out - 1
        dev = rays.device
        f32 = dict(dtype=torch.float32, device=dev)
        weights, alphas = torch.empty((B, K), **f32), torch.empty((B, K), **f32)
        depth, dino, rgb = torch.empty((B,), **f32), torch.empty((B, D), **f32), torch.empty((B, 3 * nv_c), **f32)
        invalid = torch.empty((B, K, nv_c), **f32)
        invf = torch.empty((B, K, 1), dtype=torch.uint8, device=dev)
        rgbs = torch.empty((B, K, 3 * nv_c), **f32) if want_rgb_samps else None
        cfg = self._cfg()
        lib = _abi.lib()
        # reduced precision: render on the projected map (made once per encode and head): the gather reads 128 projected
        # channels per tap instead of 256 features and layer 1 shrinks to identity + code block
        # (a head with more than 64 feature outputs -- the 768-d variant -- always renders on the projected map: the
        # composite then sums the 128 hidden units per ray and W_out is applied to the sums, whatever D is)
        use_proj = (prec == _abi.SD_MLP_F16_TC and st["C"] == 256 and mlp.d_hidden == 128
                    and (D > 64 or Bp * K >= 65536))
        for b in range(sb):
            sc = net._scene(st, b, net._projection(st, b, mlp) if use_proj else None)
            sl = slice(b * Bp, (b + 1) * Bp)
            need = lib.sd_render_workspace_bytes(C.byref(sc), C.byref(mlp), Bp, K)
            ws = torch.empty((need,), dtype=torch.uint8, device=dev) if need else None
            _abi.check(lib.sd_render_pass(
                C.byref(sc), C.byref(mlp), C.byref(cfg), _ptr(rays[sl]), Bp, rays.shape[1], _ptr(z[sl]), K,
                _ptr(depth[sl]), _ptr(dino[sl]), _ptr(rgb[sl]), _ptr(weights[sl]), _ptr(alphas[sl]),
                _ptr(invalid[sl]), _ptr(invf[sl]), _ptr(rgbs[sl]) if want_rgb_samps else None, None,
                _ptr(ws), need, _stream()), "sd_render_pass")
        state = {"invalid_features": invf.view(torch.bool), "dino_features": dino}
        return weights, rgb, depth, alphas, invalid, rgbs, state

    def _composite_generic(self, model, rays, z, coarse, sb):
        """Any other model: evaluate it in eval_batch_size chunks like nerf.py:268-341, then
        sd_composite."""
        B, K = z.shape
        pts = (rays[:, None, :3] + z.unsqueeze(2) * rays[:, None, 3:6]).reshape(-1, 3)
        info = rays[:, None, 8:].expand(-1, K, -1) if rays.shape[-1] > 8 else None
        if sb > 0:
            pts = pts.reshape(sb, -1, 3)
            info = info.reshape(sb, -1, info.shape[-1]) if info is not None else None
            dim, ebs = 1, (self.eval_batch_size - 1) // sb + 1
        else:
            dim, ebs = 0, self.eval_batch_size
        r_all, i_all, s_all, st_all = [], [], [], []
        infos = torch.split(info, ebs, dim=dim) if info is not None else None
        for i, p in enumerate(torch.split(pts, ebs, dim=dim)):
            rgbs, invalid, sigmas, extras, sd = model(p, coarse=coarse, only_density=self.only_surface_color,
                                                      ray_info=None if infos is None else infos[i],
                                                      render_flow=self.render_flow)
            if extras is not None:
                raise NotImplementedError("models returning extras are not implemented")
            r_all.append(rgbs); i_all.append(invalid); s_all.append(sigmas); st_all.append(sd)
        graph = torch.is_grad_enabled() and any(t.requires_grad for t in s_all)      # training: keep the autograd graph
        keep = (lambda t: t.float().contiguous()) if graph else _f32c
        rgbs = keep(torch.cat(r_all, dim=dim)).reshape(B, K, -1)
        invalid = torch.cat(i_all, dim=dim).reshape(B, K, -1)
        sigmas = keep(torch.cat(s_all, dim=dim)).reshape(B, K)
        state = {k: torch.cat([s[k] for s in st_all], dim=dim) for k in st_all[0].keys()}
        state = {k: v.reshape(B, K, *v.shape[2:]) for k, v in state.items()}
        feat = keep(state["dino_features"])
        D, Crgb = feat.shape[-1], rgbs.shape[-1]
        if torch.is_grad_enabled() and (sigmas.requires_grad or feat.requires_grad or rgbs.requires_grad):
            from .autograd import CompositeFn          # training: the composite with its custom backward (sd_composite_bwd)
            weights, alphas, depth, dino, rgb = CompositeFn.apply(z, sigmas, feat, rgbs, self.hard_alpha_cap, self.white_bkgd)
            state["dino_features"] = dino
            return weights, rgb, depth, alphas, invalid, rgbs, state
        f32 = dict(dtype=torch.float32, device=rays.device)
        weights, alphas = torch.empty((B, K), **f32), torch.empty((B, K), **f32)
        depth, dino, rgb = torch.empty((B,), **f32), torch.empty((B, D), **f32), torch.empty((B, Crgb), **f32)
        cfg = self._cfg()
        _abi.check(_abi.lib().sd_composite(_ptr(z), _ptr(sigmas), _ptr(feat), _ptr(rgbs), B, K, D, Crgb, C.byref(cfg),
                                           _ptr(weights), _ptr(alphas), _ptr(depth), _ptr(dino), _ptr(rgb),
                                           _stream()), "sd_composite")
        state["dino_features"] = dino
        return weights, rgb, depth, alphas, invalid, rgbs, state

    # ---- forward -------------------------------------------------------------------------------------
    @device_guard
    def forward(self, model, rays, want_weights=False, want_alphas=False, want_z_samps=False,
                want_rgb_samps=False, sample_from_dist=None):
        """nerf.py:451-539: rays (SB, B, >=8) -> DotMap(coarse=..., [fine=...], state_dict=...)."""
        with torch.profiler.record_function("renderer_forward"):
            if self.sched is not None and self.last_sched.item() > 0:
                self.n_coarse = self.sched[1][self.last_sched.item() - 1]
                self.n_fine = self.sched[2][self.last_sched.item() - 1]
            assert len(rays.shape) == 3
            sb = rays.shape[0]
            rays = rays.reshape(-1, rays.shape[-1])
            if (sample_from_dist is None and (hasattr(model, "_sd_render_pass") or hasattr(model, "_scene"))
                    and not self._wants_grad(model)):
                passes = self._forward_native(model, rays, sb, want_rgb_samps)
                fmt = dict(want_weights=want_weights, want_alphas=want_alphas, want_z_samps=want_z_samps,
                           want_rgb_samps=want_rgb_samps)
                outputs = DotMap(coarse=self._format_outputs(passes[0], sb, **fmt))
                outputs.state_dict = passes[0][-1]
                if len(passes) > 1:
                    outputs.fine = self._format_outputs(passes[1], sb, **fmt)
                return outputs
            if sample_from_dist is None:
                z_coarse = self.sample_coarse(rays)
            else:
                pw, pz = sample_from_dist
                n = pw.shape[-1]
                z_coarse = self._sort_rows(self.sample_coarse_from_dist(rays, pw.reshape(-1, n), pz.reshape(-1, n)))
            fmt = dict(want_weights=want_weights, want_alphas=want_alphas, want_z_samps=want_z_samps,
                       want_rgb_samps=want_rgb_samps)
            coarse = self._composite(model, rays, z_coarse, True, sb, want_rgb_samps)
            outputs = DotMap(coarse=self._format_outputs(coarse, sb, **fmt))
            outputs.state_dict = coarse[-1]
            if self.using_fine:
                samps = [z_coarse]
                if self.n_fine - self.n_fine_depth > 0:
                    samps.append(self.sample_fine(rays, coarse[0].detach()))
                if self.n_fine_depth > 0:
                    samps.append(self.sample_fine_depth(rays, coarse[2]))
                z_all = self._sort_rows(torch.cat(samps, dim=-1))
                fine = self._composite(model, rays, z_all, False, sb, want_rgb_samps)
                outputs.fine = self._format_outputs(fine, sb, **fmt)
            return outputs

    def _format_outputs(self, rendered_outputs, superbatch_size, want_weights=False, want_alphas=False,
                        want_z_samps=False, want_rgb_samps=False):
        """nerf.py:541-598 (shapes, keys and the invalid_features reshape quirk included)."""
        weights, rgb_final, depth, alphas, invalid, z_samps, rgb_samps, ray_info, extras, state_dict = rendered_outputs
        n_smps = weights.shape[-1]
        out_d_rgb, out_d_i = rgb_final.shape[-1], invalid.shape[-1]
        out_d_dino = state_dict["dino_features"].shape[-1]
        sb = superbatch_size
        if sb > 0:
            rgb_final = rgb_final.reshape(sb, -1, out_d_rgb)
            depth = depth.reshape(sb, -1)
            invalid = invalid.reshape(sb, -1, n_smps, out_d_i)
        ret = DotMap(rgb=rgb_final, depth=depth, invalid=invalid)
        if ray_info is not None:
            ret.ray_info = ray_info.reshape(sb, -1, ray_info.shape[-1])
        if extras is not None:
            ret.extras = extras.reshape(sb, -1, extras.shape[-1])
        if want_weights:
            ret.weights = weights.reshape(sb, -1, n_smps)
        if want_alphas:
            ret.alphas = alphas.reshape(sb, -1, n_smps)
        if want_z_samps:
            ret.z_samps = z_samps.reshape(sb, -1, n_smps)
        if want_rgb_samps:
            ret.rgb_samps = rgb_samps.reshape(sb, -1, n_smps, out_d_rgb)
        if "dino_features" in state_dict:
            ret.dino_features = state_dict["dino_features"].reshape(sb, -1, out_d_dino)
        if "invalid_features" in state_dict:
            ret.invalid_features = state_dict["invalid_features"].reshape(sb, -1, n_smps, out_d_i)
        return ret

    def sched_step(self, steps=1):
        """nerf.py:600-620."""
        if self.sched is None:
            return
        self.iter_idx += steps
        while self.last_sched.item() < len(self.sched[0]) and self.iter_idx.item() >= self.sched[0][self.last_sched.item()]:
            self.n_coarse = self.sched[1][self.last_sched.item()]
            self.n_fine = self.sched[2][self.last_sched.item()]
            print("INFO: NeRF sampling resolution changed on schedule ==> c", self.n_coarse, "f", self.n_fine)
            self.last_sched += 1

    @classmethod
    def from_conf(cls, conf, white_bkgd=False, eval_batch_size=100000):
        """nerf.py:622-639 (note lindisp defaults to True here, False in __init__)."""
        return cls(
            conf.get("n_coarse", 128), conf.get("n_fine", 0), n_fine_depth=conf.get("n_fine_depth", 0),
            noise_std=conf.get("noise_std", 0.0), depth_std=conf.get("depth_std", 0.01),
            white_bkgd=conf.get("white_bkgd", white_bkgd), lindisp=conf.get("lindisp", True),
            eval_batch_size=conf.get("eval_batch_size", eval_batch_size), sched=conf.get("sched", None),
            hard_alpha_cap=conf.get("hard_alpha_cap", False), render_mode=conf.get("render_mode", "volumetric"),
            surface_sigmoid_scale=conf.get("surface_sigmoid_scale", 1), render_flow=conf.get("render_flow", False),
            normalize_dino=conf.get("normalize_dino", False))

    def bind_parallel(self, net, gpus=None, simple_output=False):
        """nerf.py:641-658.  The reference wraps the renderer in ``nn.DataParallel(dim=1)`` when ``gpus`` has more than one
        entry -- one process, rays split along dim 1.  Here multi-GPU is one process PER GPU (torch.distributed): with an
        initialised process group whose size equals ``len(gpus)`` the returned module renders this rank's contiguous tile of
        the rays and all-gathers the per-ray outputs (scenedino_b200.sharding.render_rays_sharded), so that every rank gets
        the full result like DataParallel's caller does; without a process group the request is an error."""
        wrapped = _RenderWrapper(net, self, simple_output=simple_output)
        if gpus is None or len(gpus) <= 1:
            return wrapped
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() == len(gpus)):
            raise NotImplementedError(
                f"bind_parallel(gpus={list(gpus)}): scenedino_b200 runs one process per GPU -- launch {len(gpus)} ranks with "
                "torchrun and call bind_parallel on each (rays are sharded by scenedino_b200.sharding), not nn.DataParallel")
        return _ShardedRenderWrapper(wrapped)


class _ShardedRenderWrapper(torch.nn.Module):
    """The multi-process counterpart of ``DataParallel(_RenderWrapper, dim=1)`` (nerf.py:654-658): every rank renders its
    contiguous tile of rays [n, R/world, .] and the outputs are all-gathered along the ray dimension."""

    def __init__(self, wrapped: _RenderWrapper):
        super().__init__()
        self.module = wrapped
        self.net, self.renderer = wrapped.net, wrapped.renderer

    def forward(self, rays, **kwargs):
        from .sharding import all_gather_ragged, shard_slice
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        R = rays.shape[1]
        sl = shard_slice(R, rank, world)
        local = self.module(rays[:, sl].contiguous(), **kwargs)
        if isinstance(local, tuple):                       # simple_output: (rgb, depth)
            return tuple(all_gather_ragged(t, R, dim=1) for t in local)
        out = {}
        for level, part in local.items():
            if level == "state_dict":
                out[level] = part                            # per-sample state stays sharded, like the 64-d voxel features
                continue
            out[level] = {k: (all_gather_ragged(v, R, dim=1) if torch.is_tensor(v) and v.dim() >= 2 and v.shape[1] == sl.stop - sl.start
                              else v) for k, v in part.items()}
        return out
