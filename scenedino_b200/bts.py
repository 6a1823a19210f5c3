"""``BTSNet`` -- host-side mirror of the reference field model (scenedino/models/bts.py:22-595).

Same constructor, attribute names (``encoder``, ``code_xyz``, ``heads``, ``empty_feature``: the
state-dict keys reference checkpoints carry), ``encode`` / ``sample_features`` / ``sample_colors`` /
``forward`` signatures and return conventions.  The DINO encoder stays a PyTorch module; everything
per 3-D point (projection, frustum mask, bilinear gather, positional code, MLP head, softplus, colour
lookup) runs in libscenedino_b200 through the C ABI.  There is no PyTorch fallback for those stages.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _abi
from .heads import (MlpDimReduction, ResnetFC, _f32c, _ptr, _stream, device_guard, expand_precision, note_field_precision,
                    require_cuda)

PRECISIONS = {"fp32": _abi.SD_MLP_FP32, "fp16": _abi.SD_MLP_F16_TC, "fp32_tc": _abi.SD_MLP_F32_TC}


class _PoseInverse:
    """``torch.linalg.inv_ex(poses).inverse`` for the [n, v, 4, 4] poses of ``BTSNet.encode`` (bts.py:125-126), replayed
    from a CUDA graph from the second call with a given shape on.  The solver path behind the 4 x 4 inverse is eleven small
    kernels and 0.9 ms of host time per call -- half of a 1.9 ms SSC frame; a replay issues the very same kernels (the same
    bits) in ~20 us.  Anything unusual -- poses that require grad, an enclosing capture, a capture that fails -- takes the
    eager call."""

    def __init__(self):
        self._by_shape = {}

    @staticmethod
    def _eager(poses):
        return torch.linalg.inv_ex(poses).inverse

    def __call__(self, poses: torch.Tensor) -> torch.Tensor:
        if (not poses.is_cuda or poses.requires_grad or poses.dtype != torch.float32 or poses.numel() == 0
                or poses.device.index != torch.cuda.current_device() or torch.cuda.is_current_stream_capturing()):
            return self._eager(poses)
        key = (tuple(poses.shape), poses.device)
        ent = self._by_shape.get(key)
        if ent is None:                   # first call with this shape: eager (it also creates the solver handles)
            self._by_shape[key] = "seen"
            return self._eager(poses)
        if ent == "seen":
            try:
                with torch.inference_mode(False):     # a plain tensor: later encodes may run in or outside inference mode
                    static_in = torch.empty(poses.shape, dtype=poses.dtype, device=poses.device)
                static_in.copy_(poses)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    static_out = self._eager(static_in)
                ent = (graph, static_in, static_out)
            except Exception:             # noqa: BLE001 -- a solver path that cannot be captured: stay eager, same results
                ent = "eager"
            self._by_shape[key] = ent
        if ent == "eager":
            return self._eager(poses)
        graph, static_in, static_out = ent
        try:
            static_in.copy_(poses)
            graph.replay()
            return static_out.clone()
        except RuntimeError:              # (e.g. autograd-mode restrictions on the static buffers): same results, eagerly
            self._by_shape[key] = "eager"
            return self._eager(poses)


class BTSNet(nn.Module):
    def __init__(self, conf, encoder: nn.Module, code_xyz, heads: dict, final_pred_head: str | None = None,
                 uncertainty_predictor: nn.Module | None = None, ren_nc=None,
                 downstream_head: nn.Module | None = None):
        super().__init__()
        self.encoder = encoder
        self.code_xyz = code_xyz
        self.heads = nn.ModuleDict(heads)
        self.uncertainty_predictor = uncertainty_predictor
        self.extra_outs = getattr(self.encoder, "extra_outs", 0)
        self.final_pred_head = final_pred_head if final_pred_head else list(self.heads.keys())[0]
        self.requires_bottleneck_feats = False
        self.use_viewdirs = conf.get("use_viewdirs", False)
        self.d_min, self.d_max = conf.get("z_near", 3), conf.get("z_far", 80)
        self.learn_empty = conf.get("learn_empty", True)
        self.empty_empty = conf.get("empty_empty", False)
        self.inv_z = conf.get("inv_z", True)
        self.color_interpolation = conf.get("color_interpolation", "bilinear")
        self.code_mode = conf.get("code_mode", "z")
        self.flip_augmentation = conf.get("flip_augmentation", False)
        self.return_sample_depth = conf.get("return_sample_depth", False)
        self.sample_color = conf.get("sample_color", True)
        self.predict_dino = conf.get("predict_dino", False)

        d_in = self.encoder.latent_size + self.code_xyz.d_out
        if self.sample_color and self.predict_dino:
            d_out = 1 + conf.get("dino_dims", 16)
        elif self.sample_color:
            d_out = 1
        else:
            d_out = 4
        self._d_in, self._d_out = d_in, d_out

        # what the fused kernels implement = what every shipped config selects
        # (configs/model/dino_downsampler.yaml:30-62)
        if self.code_mode != "z":
            raise NotImplementedError(f"code_mode={self.code_mode!r}: only 'z' is implemented")
        if self.use_viewdirs or not self.sample_color or self.color_interpolation != "bilinear":
            raise NotImplementedError("use_viewdirs / sample_color=False / non-bilinear colours are not implemented")
        if self.extra_outs:
            raise NotImplementedError("encoder.extra_outs > 0 is not implemented")
        head = self.heads[self.final_pred_head]
        if not isinstance(head, ResnetFC) and not (hasattr(head, "lin_in") and hasattr(head, "lin_out")
                                                   and getattr(head, "n_blocks", 0) == 0):
            raise NotImplementedError("the prediction head must be a ResnetFC with n_blocks=0")

        if self.learn_empty:
            self.empty_feature = nn.Parameter(torch.randn((self.encoder.latent_size,), requires_grad=True))
        self._scale = 0
        self.downstream_head = downstream_head
        self.gt_classes = downstream_head.gt_classes if downstream_head is not None else None

        #: "fp32": CUDA-core FFMA head on the fp32 map (rel 1e-4 parity mode);
        #: "fp16": tcgen05 head (half operands, fp32 accumulate) on the fp16 map (rel 2e-2, the throughput mode);
        #: "auto": fp16 under torch autocast (the dtype the reference runs the head in there), fp32 otherwise.
        self.precision = conf.get("sd_precision", "auto")
        self.encode_loss_features = True
        #: forward(predict_segmentation=True) returns the 768-d expansion as its first element like the reference
        #: (bts.py:585); False skips it (first element None) -- the SSC evaluation only keeps sigma and seg
        self.materialize_dino_full = True
        #: forward(predict_segmentation=True) returns the labels one-hot [n, N, gt_classes] int64 like the reference
        #: (bts.py:588: 152 B per voxel); False returns them as they leave the kernel: uint8 [n, N]
        self.one_hot_seg = True
        #: the caller queries the SAME points tensor (same storage, unchanged contents) against the SAME encoder camera
        #: frame after frame -- the SSC evaluation does (sscbench/evaluate_model_sscbench.py:270-279 builds the grid
        #: once) --: the texel sort of the points is then kept across encode() calls and only the tile kernel runs
        #: (sd_query_points_sorted).  Off by default: the library cannot see in-place edits of the points.  True trusts the
        #: caller on the camera too; "check" compares the encoder camera at every encode (one small device->host read-back).
        self.static_query = False
        self._static_cache = {}
        #: the pose inverse of encode() replayed from a CUDA graph (same kernels, same bits, no 0.9 ms of host time per
        #: frame); False calls torch.linalg.inv_ex every time
        self.graph_pose_inverse = True
        self._pose_inverse = _PoseInverse()
        self._ids_cache = {}
        self._packed = {}
        self._head_packed = None
        self.grid_f_features = None

    # ---- trivial accessors (bts.py:103-110) ------------------------------------------------------
    def set_scale(self, scale):
        self._scale = scale

    def get_scale(self):
        return self._scale

    def compute_grid_transforms(self, *args, **kwargs):
        pass

    def reset_static_query(self):
        """Drops the point sorts kept under ``static_query``."""
        self._static_cache.clear()

    def _take_views(self, x: torch.Tensor, ids):
        """``x[:, ids]`` for a list of view ids, with the index tensor kept on the device of ``x``"""
        if torch.is_tensor(ids) or not x.is_cuda:
            return x[:, ids]
        key = (tuple(int(i) for i in ids), x.device)
        idx = self._ids_cache.get(key)
        if idx is None:
            if len(self._ids_cache) >= 64:
                self._ids_cache.clear()
            idx = self._ids_cache[key] = torch.tensor(key[0], dtype=torch.long, device=x.device)
        return x[:, idx]

    # ---- encode (bts.py:112-259) -----------------------------------------------------------------
    def encode(self, images, Ks, poses_c2w, ids_encoder=None, ids_render=None, ids_loss=None, images_alt=None,
               combine_ids=None, color_frame_filter=None, loss_feature_grid_shift=None):
        if combine_ids is not None or color_frame_filter is not None:
            raise NotImplementedError("combine_ids / color_frame_filter are not implemented")
        if loss_feature_grid_shift is not None and tuple(loss_feature_grid_shift) != (0, 0):
            raise NotImplementedError("loss_feature_grid_shift is not implemented")
        if self.flip_augmentation and self.training:
            raise NotImplementedError("flip_augmentation is a training-time option; not implemented")
        with torch.autocast(device_type=images.device.type, enabled=False):
            # bts.py:125-126 calls torch.inverse, which reads the LU status back (a device sync per encode).  inv_ex is the
            # same routine -- the same bits -- without the read-back; rigid poses are never singular.  From the second
            # encode with the same batch shape on, the call is replayed from a CUDA graph (_PoseInverse).
            poses_f = poses_c2w.float()
            poses_w2c = self._pose_inverse(poses_f) if self.graph_pose_inverse else torch.linalg.inv_ex(poses_f).inverse

        # the reference indexes with the Python lists themselves (bts.py:128-141): each such index is a host -> device copy
        # of the list followed by a stream synchronisation -- six per encode, and the host then waits for the previous
        # frame's kernels.  The same index tensors, made once per (ids, device), give the same copies without the wait.
        take = self._take_views
        if ids_encoder is None:
            images_encoder, Ks_encoder, poses_w2c_encoder = images, Ks, poses_w2c
        else:
            images_encoder, Ks_encoder, poses_w2c_encoder = take(images, ids_encoder), take(Ks, ids_encoder), take(poses_w2c, ids_encoder)
        images_loss = images if ids_loss is None else take(images, ids_loss)
        images = images_alt if images_alt is not None else images * 0.5 + 0.5
        if ids_render is None:
            images_render, Ks_render, poses_w2c_render = images, Ks, poses_w2c
        else:
            images_render, Ks_render, poses_w2c_render = take(images, ids_render), take(Ks, ids_render), take(poses_w2c, ids_render)

        n_, nv_, c_, h_, w_ = images_encoder.shape
        image_latents_ms = self.encoder(images_encoder.reshape(n_ * nv_, c_, h_, w_))
        if self.encode_loss_features:
            nl_, nvl_ = images_loss.shape[:2]
            loss_ms = self.encoder(images_loss.reshape(nl_ * nvl_, *images_loss.shape[2:]), ground_truth=True)
            self.grid_l_loss_features = [l.view(nl_, nvl_, -1, *l.shape[-2:]) for l in loss_ms]
        else:
            self.grid_l_loss_features = None

        hf, wf = image_latents_ms[0].shape[-2:]
        for l in image_latents_ms:
            if tuple(l.shape[-2:]) != (hf, wf):
                raise NotImplementedError("multi-scale encoder outputs of different sizes are not implemented")
        # bts.py:217-222 copies every map through F.interpolate(size=same); here the one re-layout
        # pass (sd_featmap_pack, planar -> channels-last) replaces that copy, lazily per scale/dtype.
        self.grid_f_features = [l.view(n_, nv_, -1, hf, wf) for l in image_latents_ms]
        self.grid_f_extra = None
        self.grid_f_Ks = Ks_encoder
        self.grid_f_poses_w2c = poses_w2c_encoder
        self.grid_f_combine = None
        self.grid_c_imgs = images_render.detach()
        self.grid_c_Ks = Ks_render
        self.grid_c_poses_w2c = poses_w2c_render
        self.grid_c_combine = None
        self.color_frame_filter = None
        self._packed = {}

    # ---- C-ABI plumbing ----------------------------------------------------------------------------
    def sd_tensors(self):
        """Tensors whose device every call must share (heads.on_device)."""
        head = self.heads[self.final_pred_head]
        ts = [head.lin_in.weight]
        if self.grid_f_features is not None:
            ts.append(self.grid_f_features[self._scale])
        return ts

    def _precision(self) -> int:
        p = self.precision
        if p == "auto":
            p = "fp16" if torch.is_autocast_enabled() else "fp32"
        if p not in PRECISIONS:
            raise ValueError(f"precision must be one of fp32|fp32_tc|fp16|auto, got {self.precision!r}")
        note_field_precision(PRECISIONS[p])
        return PRECISIONS[p]

    def _state(self, precision: int):
        """Packed, contiguous device state for the current scale: (feat_nhwc, dtype, cams...)."""
        if self.grid_f_features is None:
            raise RuntimeError("BTSNet.encode must be called before querying the field")
        dt = _abi.SD_F16 if precision == _abi.SD_MLP_F16_TC else _abi.SD_F32
        key = (self._scale, dt)
        st = self._packed.get(key)
        if st is None:
            fmap = self.grid_f_features[self._scale]
            require_cuda(fmap, "the encoder feature map")
            n, nv, c, hf, wf = fmap.shape
            if nv != 1:
                raise NotImplementedError("the default head supports exactly one encoder view (ids_encoder=[0])")
            src = _f32c(fmap)
            dst = torch.empty((n, nv, hf, wf, c), device=src.device,
                              dtype=torch.float16 if dt == _abi.SD_F16 else torch.float32)
            _abi.check(_abi.lib().sd_featmap_pack(_ptr(src), n * nv, c, hf, wf, _ptr(dst), dt, _stream()),
                       "sd_featmap_pack")
            st = dict(
                feat=dst, dt=dt, n=n, C=c, Hf=hf, Wf=wf,
                K_f=_f32c(self.grid_f_Ks).reshape(n, nv, 3, 3).contiguous(),
                w2c_f=_f32c(self.grid_f_poses_w2c).reshape(n, nv, 4, 4).contiguous(),
                rgb=_f32c(self.grid_c_imgs),
                K_c=_f32c(self.grid_c_Ks).contiguous(), w2c_c=_f32c(self.grid_c_poses_w2c).contiguous(),
            )
            if st["rgb"].shape[2] != 3:
                raise NotImplementedError("colour views must have 3 channels")
            if self.static_query == "check":   # one small read-back (a host sync) per encode: the kept sort is only valid for this camera
                st["cam_sig"] = bytes(torch.cat([st["K_f"].reshape(-1), st["w2c_f"].reshape(-1)]).cpu().numpy().tobytes())
            self._packed[key] = st
        return st

    def _projection(self, st, b: int, mlp: _abi.SdMlp):
        """Blob of sd_field_project (fp16 state) / sd_field_project_x3 (fp32 state, SD_MLP_F32_TC) for batch element ``b``
        and this head, made once per encode: the map pushed through the feature columns of the head's first layer, for the
        projected-map tile kernel."""
        # the blob bakes in W_in[:, :C] and W_feat . empty_feature: key it on the pack generation of the head (a repack
        # can land on a recycled allocator address) and on the identity / version of empty_feature
        cache = st.setdefault("proj", {})
        pm = self._packed_mlp()
        ef = self.empty_feature if self.learn_empty else None
        key = (b, id(pm), pm.generation, None if ef is None else (ef.data_ptr(), ef._version))
        for k in [k for k in cache if k[0] == b and k != key]:
            del cache[k]                      # stale projections of this batch element
        blob = cache.get(key)
        if blob is None:
            sc = self._scene(st, b)
            lib = _abi.lib()
            x3 = mlp.precision == _abi.SD_MLP_F32_TC
            nbytes = (lib.sd_field_project_x3_bytes if x3 else lib.sd_field_project_bytes)(C.byref(sc))
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=st["feat"].device)
            off = (-raw.data_ptr()) % 1024
            blob = raw[off:off + nbytes]
            _abi.check((lib.sd_field_project_x3 if x3 else lib.sd_field_project)(C.byref(sc), C.byref(mlp), _ptr(blob), nbytes, _stream()),
                       "sd_field_project_x3" if x3 else "sd_field_project")
            cache[key] = blob
        return blob

    def _scene(self, st, b: int, proj=None) -> _abi.SdScene:
        s = _abi.SdScene()
        if proj is not None:
            if st["dt"] == _abi.SD_F16:
                s.feat_proj = proj.data_ptr()
            else:                              # (the fp32 state's projection is the x3 one)
                s.feat_proj_x3 = proj.data_ptr()
        s.feat = st["feat"][b].data_ptr()
        s.feat_dtype = st["dt"]
        s.nv_f, s.C, s.Hf, s.Wf = 1, st["C"], st["Hf"], st["Wf"]
        s.K_f = st["K_f"][b].data_ptr()
        s.w2c_f = st["w2c_f"][b].data_ptr()
        rgb = st["rgb"]
        s.rgb = rgb[b].data_ptr()
        s.nv_c, s.Hc, s.Wc = rgb.shape[1], rgb.shape[3], rgb.shape[4]
        s.K_c = st["K_c"][b].data_ptr()
        s.w2c_c = st["w2c_c"][b].data_ptr()
        s.d_min, s.d_max, s.inv_z = float(self.d_min), float(self.d_max), int(bool(self.inv_z))
        s.num_freqs = int(self.code_xyz.num_freqs)
        s.freq_factor = float(getattr(self.code_xyz, "freq_factor", float(self.code_xyz.freqs[0])))
        s.include_input = int(bool(self.code_xyz.include_input))
        s.learn_empty = int(bool(self.learn_empty))
        if self.learn_empty:
            self._empty_f32 = _f32c(self.empty_feature)
            s.empty_feature = self._empty_f32.data_ptr()
        return s

    def _packed_mlp(self):
        head = self.heads[self.final_pred_head]
        if isinstance(head, ResnetFC):
            return head._packed
        if self._head_packed is None:
            from .heads import PackedMlp
            self._head_packed = PackedMlp()
        return self._head_packed

    def _mlp(self, precision: int) -> _abi.SdMlp:
        head = self.heads[self.final_pred_head]
        return self._packed_mlp().get(head.lin_in, head.lin_out, precision)

    def _check_points(self, xyz):
        require_cuda(xyz, "xyz")
        if xyz.dim() != 3 or xyz.shape[-1] != 3:
            raise ValueError(f"xyz must be [n, n_pts, 3], got {tuple(xyz.shape)}")
        return _f32c(xyz)

    def _wants_grad(self) -> bool:
        """Training mode with autograd on: the differentiable (unfused) path, SURVEY 8f-4."""
        return torch.is_grad_enabled() and self.training

    def _forward_train(self, xyz, only_density=False):
        """BTSNet.forward (bts.py:476-595) with a gradient: feature gather (custom backward into the encoder map and the
        learned empty feature), the head as torch modules, colours from the images (no gradient).  Same returns."""
        from .autograd import SampleFeaturesFn
        xyz = self._check_points(xyz)
        st = self._state(_abi.SD_MLP_FP32)
        n, N, _ = xyz.shape
        fmap = self.grid_f_features[self._scale]                  # [n, 1, C, Hf, Wf], part of the encoder's graph
        d_in = st["C"] + self.code_xyz.d_out
        head = self.heads[self.final_pred_head]
        feats, invs = [], []
        for b in range(n):
            sc = self._scene(st, b)
            f, inv = SampleFeaturesFn.apply(fmap[b, 0], self.empty_feature if self.learn_empty else None, xyz[b], sc,
                                            (st, getattr(self, "_empty_f32", None)), d_in)
            feats.append(f); invs.append(inv)
        feat = torch.stack(feats, 0)                              # [n, N, C + code]
        invalid_features = torch.stack(invs, 0).view(torch.bool).unsqueeze(-1)        # [n, N, 1]
        mlp_output = head.lin_out(head.activation(head.lin_in(feat)))                 # resnetfc.py:162-199, n_blocks = 0
        sigma = torch.nn.functional.softplus(mlp_output[..., :1])
        dino = mlp_output[..., 1:]
        if only_density:
            rgb = torch.zeros((n, N, st["rgb"].shape[1] * 3), device=xyz.device)
            invalid = invalid_features.to(sigma.dtype)
        else:
            with torch.no_grad():
                rgb, invalid_colors = self.sample_colors(xyz)     # (n, nv, N, 3), (n, nv, N, 1)
            nv_ = rgb.shape[1]
            rgb = rgb.permute(0, 2, 1, 3).reshape(n, N, nv_ * 3)
            invalid_colors = invalid_colors.permute(0, 2, 1, 3).reshape(n, N, nv_)
            invalid = (invalid_colors | torch.all(invalid_features, dim=-1)[..., None]).to(rgb.dtype)
        state_dict = {"invalid_features": invalid_features.flatten(0, 1)[None], "dino_features": dino}
        return rgb, invalid, sigma, None, state_dict

    # ---- BTSNet.sample_features (bts.py:271-328) ---------------------------------------------------
    @device_guard
    def sample_features(self, xyz):
        xyz = self._check_points(xyz)
        st = self._state(_abi.SD_MLP_FP32)
        n, N, _ = xyz.shape
        d = st["C"] + self.code_xyz.d_out
        feat = torch.empty((n, N, 1, d), dtype=torch.float32, device=xyz.device)
        inv = torch.empty((n, N, 1), dtype=torch.uint8, device=xyz.device)
        lib = _abi.lib()
        for b in range(n):
            sc = self._scene(st, b)
            _abi.check(lib.sd_sample_features(C.byref(sc), _ptr(xyz[b]), N, _ptr(feat[b]), _ptr(inv[b]), _stream()),
                       "sd_sample_features")
        return feat, inv.view(torch.bool)

    # ---- BTSNet.sample_colors (bts.py:330-441) -----------------------------------------------------
    @device_guard
    def sample_colors(self, xyz, **kwargs):
        if kwargs.get("render_flow", False):
            raise NotImplementedError("render_flow is not implemented")
        xyz = self._check_points(xyz)
        st = self._state(_abi.SD_MLP_FP32)
        n, N, _ = xyz.shape
        nv = st["rgb"].shape[1]
        rgb = torch.empty((n, N, nv, 3), dtype=torch.float32, device=xyz.device)
        inv = torch.empty((n, N, nv), dtype=torch.uint8, device=xyz.device)
        lib = _abi.lib()
        for b in range(n):
            sc = self._scene(st, b)
            _abi.check(lib.sd_sample_colors(C.byref(sc), _ptr(xyz[b]), N, _ptr(rgb[b]), _ptr(inv[b]), _stream()),
                       "sd_sample_colors")
        return rgb.permute(0, 2, 1, 3), inv.view(torch.bool).permute(0, 2, 1).unsqueeze(-1)

    # ---- BTSNet.forward (bts.py:476-595) -----------------------------------------------------------
    @device_guard
    def forward(self, xyz: torch.Tensor, **kwargs):
        only_density = kwargs.get("only_density", False)
        predict_segmentation = kwargs.get("predict_segmentation", False)
        prediction_mode = kwargs.get("prediction_mode", "stego_kmeans")
        if kwargs.get("render_flow", False):
            raise NotImplementedError("render_flow is not implemented")
        if not self.predict_dino:
            raise NotImplementedError("predict_dino=False heads are not implemented")
        if self._wants_grad():
            if predict_segmentation:
                raise NotImplementedError("predict_segmentation with autograd is not implemented (the SSC head is evaluation-only)")
            head = self.heads[self.final_pred_head]
            if not (hasattr(head, "lin_in") and hasattr(head, "lin_out") and getattr(head, "n_blocks", 0) == 0):
                raise NotImplementedError("the differentiable path needs a ResnetFC head with n_blocks = 0")
            return self._forward_train(xyz, only_density=only_density)
        with torch.profiler.record_function("model_inference"):
            xyz = self._check_points(xyz)
            prec = self._precision()
            st = self._state(prec)
            mlp = self._mlp(prec)
            n, N, _ = xyz.shape
            nv_c = st["rgb"].shape[1]
            D = mlp.d_out - 1
            dev = xyz.device
            sigma = torch.empty((n, N, 1), dtype=torch.float32, device=dev)
            dino = torch.empty((n, N, D), dtype=torch.float32, device=dev)
            invf = torch.empty((n, N, 1), dtype=torch.uint8, device=dev)
            want_colors = not (predict_segmentation or only_density)
            rgb = torch.empty((n, N, 3 * nv_c), dtype=torch.float32, device=dev) if want_colors else None
            invalid = torch.empty((n, N, nv_c), dtype=torch.float32, device=dev) if want_colors else None
            lib = _abi.lib()
            # big reduced-precision queries (SSC voxel chunks) run on the projected map: made once per encode and head
            hf, wf = st["Hf"], st["Wf"]
            use_proj = (prec in (_abi.SD_MLP_F16_TC, _abi.SD_MLP_F32_TC) and st["C"] == 256 and mlp.d_hidden == 128 and D <= 64
                        and N >= 16 * ((hf + 5) // 7 + 1) * ((wf + 5) // 7 + 1))
            # predict_segmentation with the fused head: the 64-d rows only feed sd_ssc_head, which takes a permutation --
            # they stay in the tile kernel's own (texel-bin) order and leave the SM by TMA tile stores
            # (sd_query_points_binned); nothing caller-visible changes, seg / sigma come back in the caller's order
            head = self.downstream_head
            fused = (predict_segmentation and head is not None and hasattr(head, "forward_reduced")
                     and isinstance(getattr(self.encoder, "dim_reduction", None), MlpDimReduction))
            binned = fused and use_proj and D == 64 and not self.materialize_dino_full
            perm = torch.empty((n, N), dtype=torch.int32, device=dev) if binned else None
            for b in range(n):
                sc = self._scene(st, b, self._projection(st, b, mlp) if use_proj else None)
                need = lib.sd_query_workspace_bytes(C.byref(sc), C.byref(mlp), N)
                skey = (b, xyz.data_ptr(), xyz._version, N, need, st.get("cam_sig")) if (self.static_query and use_proj and need) else None
                kept = self._static_cache.get(skey) if skey else None
                if binned and need:
                    ws, mask = kept if kept is not None else (torch.empty((need,), dtype=torch.uint8, device=dev), None)
                    _abi.check(lib.sd_query_points_binned(
                        C.byref(sc), C.byref(mlp), _ptr(xyz[b]), N, _ptr(sigma[b]), _ptr(dino[b]), _ptr(perm[b]),
                        _ptr(invf[b]), _ptr(ws), need, int(kept is not None), _stream()), "sd_query_points_binned")
                    if kept is not None:
                        invf[b].copy_(mask)
                    elif skey:
                        if len(self._static_cache) >= 8:
                            self._static_cache.clear()
                        self._static_cache[skey] = (ws, invf[b].clone())
                    continue
                if binned:          # (too few points for the sorted tile path: plain query, identity order)
                    perm = None
                    binned = False
                if kept is not None:         # same points, same camera: the sort (and the frustum mask) of the first call
                    ws, mask = kept
                    _abi.check(lib.sd_query_points_sorted(
                        C.byref(sc), C.byref(mlp), _ptr(xyz[b]), N, _ptr(sigma[b]), _ptr(dino[b]),
                        _ptr(rgb[b]) if want_colors else None, _ptr(invalid[b]) if want_colors else None,
                        _ptr(ws), need, _stream()), "sd_query_points_sorted")
                    invf[b].copy_(mask)
                    continue
                ws = torch.empty((need,), dtype=torch.uint8, device=dev) if need else None
                _abi.check(lib.sd_query_points(
                    C.byref(sc), C.byref(mlp), _ptr(xyz[b]), N, _ptr(sigma[b]), _ptr(dino[b]),
                    _ptr(rgb[b]) if want_colors else None, _ptr(invalid[b]) if want_colors else None,
                    _ptr(invf[b]), _ptr(ws), need, _stream()), "sd_query_points")
                if skey:
                    if len(self._static_cache) >= 8:
                        self._static_cache.clear()
                    self._static_cache[skey] = (ws, invf[b].clone())
            invalid_features = invf.view(torch.bool)

        if predict_segmentation:  # bts.py:528-533, 584-592
            if self.materialize_dino_full or not fused:
                with expand_precision(prec):      # the expansion runs in the precision of the query that feeds it
                    dino_full = self.encoder.expand_dim(dino)
            else:
                dino_full = None                  # 3 KB per voxel (6.4 GB per SSC grid) nobody reads: sscbench keeps sigma + seg
            seg = None
            if head is not None:
                # scenedino_b200.SemanticHead: expansion + STEGO head + cosine argmax + pseudo-label LUT in ONE kernel
                # (sd_ssc_head) starting from the 64-d features; any other head module is called like the reference does
                if fused and binned:            # rows in texel-bin order: one launch per scene, labels land at perm[r]
                    seg8 = torch.empty((n, N), dtype=torch.uint8, device=dev)
                    for b in range(n):
                        head.forward_reduced(dino[b], self.encoder.dim_reduction, mode=prediction_mode, perm=perm[b], out=seg8[b])
                    seg = seg8 if not self.one_hot_seg else torch.nn.functional.one_hot(seg8.to(torch.int64), self.gt_classes)
                elif fused and not self.one_hot_seg:
                    seg = head.forward_reduced(dino, self.encoder.dim_reduction, mode=prediction_mode,
                                               out=torch.empty((n * N,), dtype=torch.uint8, device=dev)).view(n, N)
                else:
                    seg = (head.forward_reduced(dino, self.encoder.dim_reduction, mode=prediction_mode) if fused
                           else head(dino_full, mode=prediction_mode))
                    seg = torch.nn.functional.one_hot(seg, self.gt_classes)
            return dino_full, None, sigma, seg
        if only_density:  # bts.py:570-572
            rgb = torch.zeros((n, N, nv_c * 3), device=dev)
            invalid = invalid_features.to(sigma.dtype)
        state_dict = {"invalid_features": invalid_features.flatten(0, 1)[None], "dino_features": dino}
        return rgb, invalid, sigma, None, state_dict
