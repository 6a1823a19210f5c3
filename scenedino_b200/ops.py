"""Functional layer over the C ABI: one Python function per entry point of
include/scenedino_b200.h, taking CUDA tensors.  ``BTSNet`` / ``NeRFRenderer`` are the reference-shaped
surface; this module is what the parity tests and bench.py drive directly.

Every function launches on ``torch.cuda.current_stream()`` and returns freshly allocated tensors.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _abi
from .heads import _f32c, _ptr, _stream, device_guard, on_device, require_cuda

FP32, F16 = _abi.SD_MLP_FP32, _abi.SD_MLP_F16_TC
F32TC = _abi.SD_MLP_F32_TC      # rel 1e-4 on the tensor cores: fp16 (hi, lo) operand pairs, three products per contraction


def _dev(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32).contiguous()


@device_guard
def featmap_pack(nchw: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """[n, C, H, W] fp32 -> [n, H, W, C] fp32 | fp16 (sd_featmap_pack)."""
    require_cuda(nchw, "feature map")
    src = _f32c(nchw)
    n, c, h, w = src.shape
    dst = torch.empty((n, h, w, c), dtype=dtype, device=src.device)
    if dtype not in (torch.float32, torch.float16):
        raise ValueError("feature maps are packed as float32 or float16")
    dt = _abi.SD_F16 if dtype == torch.float16 else _abi.SD_F32
    _abi.check(_abi.lib().sd_featmap_pack(_ptr(src), n, c, h, w, _ptr(dst), dt, _stream()), "sd_featmap_pack")
    return dst


class Mlp:
    """Packed two-layer head (nn.Linear layout in: w_in [H, d_in], w_out [d_out, H])."""

    def __init__(self, w_in, b_in, w_out, b_out, device="cuda", precision: int = FP32):
        self.w = [_dev(t, device) for t in (w_in, b_in, w_out, b_out)]
        self.d_hidden, self.d_in = self.w[0].shape
        self.d_out = self.w[2].shape[0]
        self.precision = precision
        lib = _abi.lib()
        nbytes = lib.sd_mlp_pack_bytes(self.d_in, self.d_hidden, self.d_out)
        with on_device(self.w):
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.w[0].device)
            off = (-raw.data_ptr()) % 1024
            self.blob = raw[off:off + nbytes]
            _abi.check(lib.sd_mlp_pack(*[_ptr(t) for t in self.w], self.d_in, self.d_hidden, self.d_out, _ptr(self.blob),
                                       _stream()), "sd_mlp_pack")

    def sd_tensors(self):
        return [self.blob]

    def c(self, precision: int | None = None) -> _abi.SdMlp:
        m = _abi.SdMlp()
        m.packed = self.blob.data_ptr()
        m.d_in, m.d_hidden, m.d_out = self.d_in, self.d_hidden, self.d_out
        m.precision = self.precision if precision is None else precision
        return m


@dataclass
class Scene:
    """Device-side state of one batch element (what BTSNet.encode stashes, bts.py:246-257)."""
    feat: torch.Tensor                  # [Hf, Wf, C] channels-last, fp32 or fp16
    K_f: torch.Tensor                   # [1,3,3]
    w2c_f: torch.Tensor                 # [1,4,4]
    rgb: torch.Tensor | None = None     # [nv_c,3,Hc,Wc]
    K_c: torch.Tensor | None = None
    w2c_c: torch.Tensor | None = None
    d_min: float = 3.0
    d_max: float = 80.0
    inv_z: bool = True
    num_freqs: int = 6
    freq_factor: float = 1.5
    include_input: bool = True
    learn_empty: bool = False
    empty_feature: torch.Tensor | None = None
    proj: torch.Tensor | None = None    # blob of sd_field_project (tied to one head), see project()
    proj_x3: torch.Tensor | None = None  # blob of sd_field_project_x3 (F32TC queries), see project_x3()

    @classmethod
    def from_arrays(cls, feat_nchw, K_f, w2c_f, rgb=None, K_c=None, w2c_c=None, device="cuda",
                    feat_dtype=torch.float32, **kw):
        """feat_nchw [1, C, Hf, Wf] (reference layout) is packed channels-last on the device."""
        f = featmap_pack(_dev(feat_nchw, device), feat_dtype)[0]
        s = cls(feat=f, K_f=_dev(K_f, device).reshape(1, 3, 3), w2c_f=_dev(w2c_f, device).reshape(1, 4, 4), **kw)
        if rgb is not None:
            s.rgb, s.K_c, s.w2c_c = _dev(rgb, device), _dev(K_c, device), _dev(w2c_c, device)
        if s.empty_feature is not None:
            s.empty_feature = _dev(s.empty_feature, device)
        return s

    def sd_tensors(self):
        return [t for t in (self.feat, self.K_f, self.w2c_f, self.rgb, self.K_c, self.w2c_c, self.empty_feature, self.proj, self.proj_x3)
                if t is not None]

    def with_feat_dtype(self, feat_nchw, dtype):
        import dataclasses
        return dataclasses.replace(self, feat=featmap_pack(_dev(feat_nchw, self.feat.device), dtype)[0])

    def project(self, mlp: "Mlp") -> "Scene":
        """Pushes the fp16 map through the feature columns of ``mlp``'s first layer (sd_field_project), once per
        encode; queries with that head then run the projected-map tile kernel.  Returns a copy of the scene."""
        import dataclasses
        if self.feat.dtype != torch.float16:
            raise ValueError("project() needs the fp16 channels-last map")
        sc, m = dataclasses.replace(self, proj=None).c(), mlp.c(F16)
        lib = _abi.lib()
        nbytes = lib.sd_field_project_bytes(C.byref(sc))
        with on_device(self, mlp):
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.feat.device)
            off = (-raw.data_ptr()) % 1024
            blob = raw[off:off + nbytes]
            _abi.check(lib.sd_field_project(C.byref(sc), C.byref(m), _ptr(blob), nbytes, _stream()), "sd_field_project")
        return dataclasses.replace(self, proj=blob)

    def project_x3(self, mlp: "Mlp") -> "Scene":
        """The projection for F32TC queries (sd_field_project_x3): fp32 map -> P in fp32 -> (hi, lo) fp16 maps + (hi, lo)
        weight images of ``mlp``.  Returns a copy of the scene."""
        import dataclasses
        if self.feat.dtype != torch.float32:
            raise ValueError("project_x3() needs the fp32 channels-last map")
        sc, m = dataclasses.replace(self, proj=None, proj_x3=None).c(), mlp.c(F32TC)
        lib = _abi.lib()
        nbytes = lib.sd_field_project_x3_bytes(C.byref(sc))
        with on_device(self, mlp):
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.feat.device)
            off = (-raw.data_ptr()) % 1024
            blob = raw[off:off + nbytes]
            _abi.check(lib.sd_field_project_x3(C.byref(sc), C.byref(m), _ptr(blob), nbytes, _stream()), "sd_field_project_x3")
        return dataclasses.replace(self, proj_x3=blob)

    @property
    def nv_c(self) -> int:
        return 0 if self.rgb is None else self.rgb.shape[0]

    @property
    def code_dim(self) -> int:
        return (3 if self.include_input else 0) + 6 * self.num_freqs

    def c(self) -> _abi.SdScene:
        s = _abi.SdScene()
        s.feat = self.feat.data_ptr()
        s.feat_dtype = _abi.SD_F16 if self.feat.dtype == torch.float16 else _abi.SD_F32
        s.nv_f = 1
        s.Hf, s.Wf, s.C = self.feat.shape
        s.K_f, s.w2c_f = self.K_f.data_ptr(), self.w2c_f.data_ptr()
        if self.rgb is not None:
            s.rgb = self.rgb.data_ptr()
            s.nv_c, _, s.Hc, s.Wc = self.rgb.shape
            s.K_c, s.w2c_c = self.K_c.data_ptr(), self.w2c_c.data_ptr()
        s.d_min, s.d_max, s.inv_z = self.d_min, self.d_max, int(self.inv_z)
        s.num_freqs, s.freq_factor, s.include_input = self.num_freqs, self.freq_factor, int(self.include_input)
        s.learn_empty = int(self.learn_empty)
        if self.empty_feature is not None:
            s.empty_feature = self.empty_feature.data_ptr()
        if self.proj is not None:
            s.feat_proj = self.proj.data_ptr()
        if self.proj_x3 is not None:
            s.feat_proj_x3 = self.proj_x3.data_ptr()
        return s


def _e(shape, ref, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=ref.device)


@device_guard
def project_points(K, w2c, xyz):
    xyz = _f32c(xyz); require_cuda(xyz, "xyz")
    N = xyz.shape[0]
    xy, z, inv = _e((N, 2), xyz), _e((N,), xyz), _e((N,), xyz, torch.uint8)
    _abi.check(_abi.lib().sd_project_points(_ptr(_f32c(K)), _ptr(_f32c(w2c)), _ptr(xyz), N, _ptr(xy), _ptr(z), _ptr(inv),
                                            _stream()), "sd_project_points")
    return xy, z, inv.view(torch.bool)


@device_guard
def sample_features(scene: Scene, xyz):
    xyz = _f32c(xyz); require_cuda(xyz, "xyz")
    N = xyz.shape[0]
    feat = _e((N, scene.feat.shape[-1] + scene.code_dim), xyz)
    inv = _e((N,), xyz, torch.uint8)
    sc = scene.c()
    _abi.check(_abi.lib().sd_sample_features(C.byref(sc), _ptr(xyz), N, _ptr(feat), _ptr(inv), _stream()),
               "sd_sample_features")
    return feat, inv.view(torch.bool)


@device_guard
def sample_colors(scene: Scene, xyz):
    xyz = _f32c(xyz); require_cuda(xyz, "xyz")
    N = xyz.shape[0]
    rgb, inv = _e((N, 3 * scene.nv_c), xyz), _e((N, scene.nv_c), xyz, torch.uint8)
    sc = scene.c()
    _abi.check(_abi.lib().sd_sample_colors(C.byref(sc), _ptr(xyz), N, _ptr(rgb), _ptr(inv), _stream()), "sd_sample_colors")
    return rgb, inv.view(torch.bool)


@device_guard
def mlp_forward(mlp: Mlp, x, precision=None):
    x = _f32c(x); require_cuda(x, "x")
    out = _e((x.shape[0], mlp.d_out), x)
    m = mlp.c(precision)
    _abi.check(_abi.lib().sd_mlp_forward(C.byref(m), _ptr(x), x.shape[0], _ptr(out), _stream()), "sd_mlp_forward")
    return out


@device_guard
def gen_rays(c2w, proj, H: int, W: int, z_near: float, z_far: float, frame_ids=None, norm_dir: bool = True,
             xy_shift=(0.0, 0.0), out=None):
    """c2w [V,4,4], proj [V,3,3] (CUDA) -> rays [V*H*W, 11] of one batch element (sd_gen_rays; util.gen_rays +
    ImageRaySampler.sample's frame-id / pixel columns)."""
    c2w = _f32c(c2w); require_cuda(c2w, "c2w")
    proj = _dev(proj, c2w.device)
    V = c2w.shape[0]
    if c2w.shape[1:] != (4, 4) or proj.shape != (V, 3, 3):
        raise ValueError(f"gen_rays: c2w {tuple(c2w.shape)} / proj {tuple(proj.shape)} are not [V,4,4] / [V,3,3]")
    ids = None if frame_ids is None else _dev(frame_ids, c2w.device).reshape(-1)
    if ids is not None and ids.numel() != V:
        raise ValueError("gen_rays: one frame id per view")
    if out is None:
        out = _e((V * H * W, 11), c2w)
    elif out.shape != (V * H * W, 11) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("gen_rays: out must be a contiguous float32 [V*H*W, 11] tensor")
    _abi.check(_abi.lib().sd_gen_rays(_ptr(c2w), _ptr(proj), _ptr(ids) if ids is not None else None, V, H, W, z_near, z_far,
                                      int(bool(norm_dir)), float(xy_shift[0]), float(xy_shift[1]), _ptr(out), _stream()),
               "sd_gen_rays")
    return out


def gen_voxel_grid(T, dims=(256, 256, 32), voxel_size=0.2, origin=(0.0, -25.6, -2.0), x_range=None, device="cuda", out=None):
    """Voxel centres of an SSC grid in the camera frame, made on the device (sd_gen_voxel_grid): [N,3] fp32, bit-identical
    to ``synthetic.ssc_voxel_grid`` and through it to the reference's grid (tests/golden/voxel_grid.npz).  ``T``: [3..4, 4]
    float64 lidar -> camera (host array); ``voxel_size`` is used as a double, as in the reference."""
    import numpy as np
    x0, x1 = (0, dims[0]) if x_range is None else x_range
    Th = np.ascontiguousarray(np.asarray(T, np.float64)[:3, :4])
    org = np.ascontiguousarray(np.asarray(origin, np.float32))
    n = (x1 - x0) * dims[1] * dims[2]
    if out is None:
        out = torch.empty((n, 3), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _abi.check(_abi.lib().sd_gen_voxel_grid(org.ctypes.data, float(voxel_size), dims[0], dims[1], dims[2], x0, x1,
                                                Th.ctypes.data, _ptr(out), _stream()), "sd_gen_voxel_grid")
    return out


@device_guard
def expand_dim(mlp: Mlp, f, precision: int = FP32):
    """MlpDimReduction.transform_expand; ``precision=F16`` takes the tensor-core kernel (64 -> 128 -> k*128 heads)."""
    f = _f32c(f); require_cuda(f, "f")
    out = _e((f.shape[0], mlp.d_out), f)
    m = mlp.c(precision)
    _abi.check(_abi.lib().sd_expand_dim(C.byref(m), _ptr(f), f.shape[0], _ptr(out), _stream()), "sd_expand_dim")
    return out


@device_guard
def query_points(scene: Scene, mlp: Mlp, xyz, want_rgb=True, precision=None, out=None, binned=True):
    """BTSNet.forward for one scene -> dict(sigma[N], dino[N,D], rgb[N,3nv_c], invalid[N,nv_c],
    invalid_features[N] bool).  ``out`` lets bench.py reuse output buffers."""
    xyz = _f32c(xyz); require_cuda(xyz, "xyz")
    N = xyz.shape[0]
    D = mlp.d_out - 1
    if out is None:
        out = dict(sigma=_e((N,), xyz), dino=_e((N, D), xyz), invalid_features=_e((N,), xyz, torch.uint8))
        if want_rgb and scene.nv_c:
            out.update(rgb=_e((N, 3 * scene.nv_c), xyz), invalid=_e((N, scene.nv_c), xyz))
    sc, m = scene.c(), mlp.c(precision)
    lib = _abi.lib()
    need = lib.sd_query_workspace_bytes(C.byref(sc), C.byref(m), N) if binned else 0
    ws = out.get("_workspace")
    if need and (ws is None or ws.numel() < need):
        ws = out["_workspace"] = torch.empty((need,), dtype=torch.uint8, device=xyz.device)
    _abi.check(lib.sd_query_points(C.byref(sc), C.byref(m), _ptr(xyz), N, _ptr(out["sigma"]), _ptr(out["dino"]),
                                   _ptr(out.get("rgb")), _ptr(out.get("invalid")), _ptr(out["invalid_features"]),
                                   _ptr(ws) if need else None, need, _stream()), "sd_query_points")
    res = {k: v for k, v in out.items() if not k.startswith("_")}
    res["invalid_features"] = out["invalid_features"].view(torch.bool)
    return res


@device_guard
def query_points_sorted(scene: Scene, mlp: Mlp, xyz, out: dict):
    """The query again for unchanged points and cameras (sd_query_points_sorted): reuses the texel sort left in
    ``out["_workspace"]`` by an earlier ``query_points(..., out=out)``; only the tile kernel runs."""
    xyz = _f32c(xyz); require_cuda(xyz, "xyz")
    ws = out.get("_workspace")
    if ws is None:
        raise ValueError("query_points_sorted needs the `out` dict of an earlier query_points call (its workspace)")
    sc, m = scene.c(), mlp.c(F16)
    _abi.check(_abi.lib().sd_query_points_sorted(C.byref(sc), C.byref(m), _ptr(xyz), xyz.shape[0], _ptr(out["sigma"]),
                                                 _ptr(out["dino"]), _ptr(out.get("rgb")), _ptr(out.get("invalid")),
                                                 _ptr(ws), ws.numel(), _stream()), "sd_query_points_sorted")
    res = {k: v for k, v in out.items() if not k.startswith("_")}
    res["invalid_features"] = out["invalid_features"].view(torch.bool)
    return res


@device_guard
def query_points_binned(scene: Scene, mlp: Mlp, xyz, out=None, reuse_sorted=False, precision=F16):
    """The point query with the 64-d features left in texel-bin order (sd_query_points_binned): dict(sigma [N],
    invalid_features [N] -- caller's order --, dino_binned [N,64], perm [N] int32: row r of dino_binned belongs to point
    perm[r]).  ``dino_binned[argsort(perm)]`` is bit-identical to ``query_points(...)["dino"]``.  ``reuse_sorted``: the
    workspace in ``out`` holds the sort of the same points and cameras (only the tile kernel runs)."""
    xyz = _f32c(xyz); require_cuda(xyz, "xyz")
    N = xyz.shape[0]
    if out is None:
        out = dict(sigma=_e((N,), xyz), dino_binned=_e((N, 64), xyz), perm=_e((N,), xyz, torch.int32),
                   invalid_features=_e((N,), xyz, torch.uint8))
    sc, m = scene.c(), mlp.c(precision)
    lib = _abi.lib()
    need = lib.sd_query_workspace_bytes(C.byref(sc), C.byref(m), N)
    ws = out.get("_workspace")
    if need and (ws is None or ws.numel() < need):
        if reuse_sorted:
            raise ValueError("query_points_binned(reuse_sorted=True) needs the `out` dict of an earlier call (its workspace)")
        ws = out["_workspace"] = torch.empty((need,), dtype=torch.uint8, device=xyz.device)
    _abi.check(lib.sd_query_points_binned(C.byref(sc), C.byref(m), _ptr(xyz), N, _ptr(out["sigma"]), _ptr(out["dino_binned"]),
                                          _ptr(out.get("perm")), _ptr(out["invalid_features"]), _ptr(ws) if need else None,
                                          need, int(bool(reuse_sorted)), _stream()), "sd_query_points_binned")
    res = {k: v for k, v in out.items() if not k.startswith("_")}
    res["invalid_features"] = out["invalid_features"].view(torch.bool)
    return res


class QueryGraph:
    """One point query (fixed scene, head, points and output buffers) captured into a CUDA graph: replaying it costs one
    launch instead of seven (the sort's memset and four kernels, the field kernel) -- for callers that query the same
    grid frame after frame, like the SSC evaluation (sscbench/evaluate_model_sscbench.py:270-279 builds the grid once).
    The scene tensors, ``xyz`` and ``out`` must stay alive and in place; their CONTENTS may change between replays."""

    def __init__(self, scene: Scene, mlp: Mlp, xyz, out: dict, want_rgb: bool = False, precision=None, binned_out: bool = False):
        self._keep = (scene, mlp, xyz, out)
        if binned_out:       # features in texel-bin order + perm (query_points_binned); ``out`` holds dino_binned / perm
            call = lambda: query_points_binned(scene, mlp, xyz, out=out)
        else:
            call = lambda: query_points(scene, mlp, xyz, want_rgb=want_rgb, precision=precision, out=out)
        side = torch.cuda.Stream(device=xyz.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # first calls: module load, function attributes, workspace allocation
            for _ in range(2):
                call()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = _abi.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            call()
        self.launches = _abi.launch_count() - n0     # kernels inside one replay

    def replay(self):
        self.graph.replay()


@device_guard
def sample_coarse(rays, u, lin, lindisp=True):
    rays, u, lin = _f32c(rays), _f32c(u), _f32c(lin)
    R, Kc = u.shape
    z = _e((R, Kc), rays)
    _abi.check(_abi.lib().sd_sample_coarse(_ptr(rays), R, rays.shape[1], _ptr(u), _ptr(lin), Kc, int(lindisp), _ptr(z),
                                           _stream()), "sd_sample_coarse")
    return z


@device_guard
def sample_fine(rays, weights, u0, u1, lindisp=True):
    rays, weights, u0, u1 = _f32c(rays), _f32c(weights), _f32c(u0), _f32c(u1)
    R, Kc = weights.shape
    Kf = u0.shape[1]
    z, inds = _e((R, Kf), rays), _e((R, Kf), rays, torch.int32)
    _abi.check(_abi.lib().sd_sample_fine(_ptr(rays), R, rays.shape[1], _ptr(weights), Kc, _ptr(u0), _ptr(u1), Kf,
                                         int(lindisp), _ptr(z), _ptr(inds), _stream()), "sd_sample_fine")
    return z, inds


@device_guard
def sample_fine_depth(rays, depth, noise, depth_std):
    rays, depth, noise = _f32c(rays), _f32c(depth), _f32c(noise)
    R, Kfd = noise.shape
    z = _e((R, Kfd), rays)
    _abi.check(_abi.lib().sd_sample_fine_depth(_ptr(rays), R, rays.shape[1], _ptr(depth), _ptr(noise), Kfd,
                                               float(depth_std), _ptr(z), _stream()), "sd_sample_fine_depth")
    return z


@device_guard
def sample_coarse_from_dist(weights, z_samp, u0, u1, lindisp=True):
    weights, z_samp, u0, u1 = _f32c(weights), _f32c(z_samp), _f32c(u0), _f32c(u1)
    R, Kp = weights.shape
    Kc = u0.shape[1]
    z, inds = _e((R, Kc), weights), _e((R, Kc), weights, torch.int32)
    _abi.check(_abi.lib().sd_sample_coarse_from_dist(R, _ptr(weights), _ptr(z_samp), Kp, _ptr(u0), _ptr(u1), Kc,
                                                     int(lindisp), _ptr(z), _ptr(inds), _stream()),
               "sd_sample_coarse_from_dist")
    return z, inds


@device_guard
def sort_rows(z):
    z = _f32c(z).clone()
    _abi.check(_abi.lib().sd_sort_rows(_ptr(z), z.shape[0], z.shape[1], _stream()), "sd_sort_rows")
    return z


def _cfg(lindisp=True, hard_alpha_cap=False, white_bkgd=False):
    c = _abi.SdRenderCfg()
    c.lindisp, c.hard_alpha_cap, c.white_bkgd = int(lindisp), int(hard_alpha_cap), int(white_bkgd)
    return c


@device_guard
def composite(z, sigma, feat, rgb=None, hard_alpha_cap=False, white_bkgd=False):
    z, sigma, feat = _f32c(z), _f32c(sigma), _f32c(feat)
    R, K = z.shape
    D = feat.shape[-1]
    Crgb = 0 if rgb is None else rgb.shape[-1]
    rgb = None if rgb is None else _f32c(rgb)
    out = dict(weights=_e((R, K), z), alphas=_e((R, K), z), depth=_e((R,), z), dino=_e((R, D), z),
               rgb=_e((R, Crgb), z) if Crgb else None)
    cfg = _cfg(True, hard_alpha_cap, white_bkgd)
    _abi.check(_abi.lib().sd_composite(_ptr(z), _ptr(sigma), _ptr(feat), _ptr(rgb), R, K, D, Crgb, C.byref(cfg),
                                       _ptr(out["weights"]), _ptr(out["alphas"]), _ptr(out["depth"]), _ptr(out["dino"]),
                                       _ptr(out["rgb"]), _stream()), "sd_composite")
    return out


@device_guard
def render_pass(scene: Scene, mlp: Mlp, rays, z, hard_alpha_cap=False, white_bkgd=False, want_rgb_samps=False,
                want_sigma=True, per_sample=True, precision=None, out=None):
    """One NeRFRenderer.composite call for one scene (sd_render_pass)."""
    rays, z = _f32c(rays), _f32c(z)
    R, K = z.shape
    D, nv_c = mlp.d_out - 1, scene.nv_c
    if out is None:
        out = dict(depth=_e((R,), z), dino_features=_e((R, D), z), rgb=_e((R, 3 * nv_c), z))
        if per_sample:
            out.update(weights=_e((R, K), z), alphas=_e((R, K), z), invalid=_e((R, K, nv_c), z),
                       invalid_features=_e((R, K), z, torch.uint8))
            if want_sigma:
                out["sigma"] = _e((R, K), z)
        if want_rgb_samps:
            out["rgb_samps"] = _e((R, K, 3 * nv_c), z)
    sc, m, cfg = scene.c(), mlp.c(precision), _cfg(True, hard_alpha_cap, white_bkgd)
    lib = _abi.lib()
    need = lib.sd_render_workspace_bytes(C.byref(sc), C.byref(m), R, K)
    ws = torch.empty((need,), dtype=torch.uint8, device=z.device) if need else None
    _abi.check(lib.sd_render_pass(C.byref(sc), C.byref(m), C.byref(cfg), _ptr(rays), R, rays.shape[1], _ptr(z), K,
                                  _ptr(out["depth"]), _ptr(out["dino_features"]), _ptr(out["rgb"]),
                                  _ptr(out.get("weights")), _ptr(out.get("alphas")), _ptr(out.get("invalid")),
                                  _ptr(out.get("invalid_features")), _ptr(out.get("rgb_samps")), _ptr(out.get("sigma")),
                                  _ptr(ws), need, _stream()), "sd_render_pass")
    res = dict(out)
    if "invalid_features" in res:
        res["invalid_features"] = res["invalid_features"].view(torch.bool)
    res["z_samps"] = z
    return res


class SscHead:
    """Folded weights of the expansion + unsupervised SSC head (sd_ssc_head_pack): ``expand`` = (w1 [128,64], b1, w2 [768,128],
    b2) of MlpDimReduction, ``head`` = dict(wl, bl, wn1, bn1, wn2, bn2, centres, lut) as synthetic.make_ssc_head lays it out."""

    def __init__(self, expand, head, device="cuda"):
        w1, b1, w2, b2 = [_dev(t, device) for t in expand]
        hs = [_dev(head[k], device) for k in ("wl", "bl", "wn1", "bn1", "wn2", "bn2", "centres")]
        lut = torch.as_tensor(head["lut"]).to(device=device, dtype=torch.int64).contiguous()
        self.d_mid, self.n_cls = hs[2].shape[0], hs[6].shape[0]
        lib = _abi.lib()
        nbytes = lib.sd_ssc_head_pack_bytes(w1.shape[1], w1.shape[0], w2.shape[0], self.d_mid, hs[0].shape[0], self.n_cls)
        if nbytes == 0:
            raise _abi.SdError("SscHead: unsupported head shape")
        with on_device(w1, lut):
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=w1.device)
            off = (-raw.data_ptr()) % 1024
            self.blob = raw[off:off + nbytes]
            _abi.check(lib.sd_ssc_head_pack(_ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), *[_ptr(h) for h in hs], _ptr(lut), w1.shape[1],
                                            w1.shape[0], w2.shape[0], self.d_mid, hs[0].shape[0], self.n_cls, _ptr(self.blob),
                                            _stream()), "sd_ssc_head_pack")

    def sd_tensors(self):
        return [self.blob]


@device_guard
def ssc_head(head: SscHead, f, want_scores=True, perm=None, out=None):
    """f [N,64] -> dict(seg [N] uint8, pseudo [N] uint8, scores [N,n_cls]) (sd_ssc_head)."""
    f = _f32c(f); require_cuda(f, "f")
    N = f.shape[0]
    if out is None:
        out = dict(seg=_e((N,), f, torch.uint8))
        if want_scores:
            out.update(pseudo=_e((N,), f, torch.uint8), scores=_e((N, head.n_cls), f))
    _abi.check(_abi.lib().sd_ssc_head(_ptr(head.blob), head.d_mid, head.n_cls, _ptr(f), _ptr(perm), N, _ptr(out["seg"]),
                                      _ptr(out.get("pseudo")), _ptr(out.get("scores")), _stream()), "sd_ssc_head")
    return out


@device_guard
def positional_encoding(x, num_freqs=6, freq_factor=1.5, include_input=True):
    x = _f32c(x); require_cuda(x, "x")
    N, d_in = x.shape
    out = _e((N, (d_in if include_input else 0) + 2 * num_freqs * d_in), x)
    _abi.check(_abi.lib().sd_positional_encoding(_ptr(x), N, d_in, num_freqs, float(freq_factor), int(bool(include_input)),
                                                 _ptr(out), _stream()), "sd_positional_encoding")
    return out
