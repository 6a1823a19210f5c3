"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front-end of the C oracle (oracle/sd_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (scenedino_b200/) must never do so.

All arrays are numpy float32 / uint8 / int32, C-contiguous, one scene (batch element) at a time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "libsd_oracle.so")
_lib = None


ORACLE_VERSION = 2          # sdo_version() of sd_oracle.c


def build(force: bool = False) -> str:
    """Compiles the oracle with the recipe in oracle/Makefile (gcc, a second or two).  Rebuilt when the source is newer than
    the library or the library was built from another version of the source (stamp file: copied trees lose timestamps)."""
    src = os.path.join(_HERE, "sd_oracle.c")
    stamp = os.path.join(os.path.dirname(_LIB_PATH), ".version")
    fresh = (os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src) and os.path.exists(stamp)
             and open(stamp).read().strip() == str(ORACLE_VERSION))
    if force or not fresh:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
        with open(stamp, "w") as f:
            f.write(str(ORACLE_VERSION))
    return _LIB_PATH


class _Scene(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p), ("nv_f", C.c_int), ("C", C.c_int), ("Hf", C.c_int), ("Wf", C.c_int),
        ("K_f", C.c_void_p), ("w2c_f", C.c_void_p),
        ("rgb", C.c_void_p), ("nv_c", C.c_int), ("Hc", C.c_int), ("Wc", C.c_int),
        ("K_c", C.c_void_p), ("w2c_c", C.c_void_p),
        ("d_min", C.c_float), ("d_max", C.c_float), ("inv_z", C.c_int), ("num_freqs", C.c_int),
        ("freq_factor", C.c_float), ("include_input", C.c_int), ("learn_empty", C.c_int),
        ("empty_feature", C.c_void_p),
    ]


class _Mlp(C.Structure):
    _fields_ = [("w_in", C.c_void_p), ("b_in", C.c_void_p), ("w_out", C.c_void_p),
                ("b_out", C.c_void_p), ("d_in", C.c_int), ("d_hidden", C.c_int), ("d_out", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.sdo_version.restype = C.c_int
        if _lib.sdo_version() != ORACLE_VERSION:
            raise RuntimeError(f"oracle library version {_lib.sdo_version()}, oracle.py expects {ORACLE_VERSION}: "
                               "delete oracle/build/ and run again")
        _lib.sdo_num_threads.restype = C.c_int
        _lib.sdo_code_dim.restype = C.c_int
        _lib.sdo_query_points.restype = C.c_int
        _lib.sdo_render_pass.restype = C.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().sdo_num_threads())


def set_num_threads(n: int) -> None:
    lib().sdo_set_num_threads(C.c_int(int(n)))


@dataclass
class Scene:
    """One batch element of what BTSNet.encode stashes (models/bts.py:246-257)."""
    feat: np.ndarray            # [nv_f, C, Hf, Wf]
    K_f: np.ndarray             # [nv_f, 3, 3]
    w2c_f: np.ndarray           # [nv_f, 4, 4]
    rgb: np.ndarray | None = None   # [nv_c, 3, Hc, Wc]
    K_c: np.ndarray | None = None
    w2c_c: np.ndarray | None = None
    d_min: float = 3.0
    d_max: float = 80.0
    inv_z: bool = True
    num_freqs: int = 6
    freq_factor: float = 1.5
    include_input: bool = True
    learn_empty: bool = False
    empty_feature: np.ndarray | None = None
    _keep: list = field(default_factory=list, repr=False)

    def c(self) -> _Scene:
        self.feat = _f32(self.feat); self.K_f = _f32(self.K_f); self.w2c_f = _f32(self.w2c_f)
        nv_f, Cc, Hf, Wf = self.feat.shape
        s = _Scene()
        s.feat = _p(self.feat); s.nv_f, s.C, s.Hf, s.Wf = nv_f, Cc, Hf, Wf
        s.K_f = _p(self.K_f); s.w2c_f = _p(self.w2c_f)
        if self.rgb is not None:
            self.rgb = _f32(self.rgb); self.K_c = _f32(self.K_c); self.w2c_c = _f32(self.w2c_c)
            s.rgb = _p(self.rgb); s.nv_c, _, s.Hc, s.Wc = self.rgb.shape
            s.K_c = _p(self.K_c); s.w2c_c = _p(self.w2c_c)
        else:
            s.nv_c = 0
        s.d_min, s.d_max, s.inv_z = self.d_min, self.d_max, int(self.inv_z)
        s.num_freqs, s.freq_factor, s.include_input = self.num_freqs, self.freq_factor, int(self.include_input)
        s.learn_empty = int(self.learn_empty)
        if self.empty_feature is not None:
            self.empty_feature = _f32(self.empty_feature)
            s.empty_feature = _p(self.empty_feature)
        return s

    @property
    def nv_c(self) -> int:
        return 0 if self.rgb is None else self.rgb.shape[0]

    @property
    def code_dim(self) -> int:
        return (3 if self.include_input else 0) + 6 * self.num_freqs


@dataclass
class Mlp:
    """ResnetFC with n_blocks = 0 (models/prediction_heads/resnetfc.py:90-96)."""
    w_in: np.ndarray   # [d_hidden, d_in]
    b_in: np.ndarray
    w_out: np.ndarray  # [d_out, d_hidden]
    b_out: np.ndarray

    def c(self) -> _Mlp:
        self.w_in = _f32(self.w_in); self.b_in = _f32(self.b_in)
        self.w_out = _f32(self.w_out); self.b_out = _f32(self.b_out)
        m = _Mlp()
        m.w_in, m.b_in, m.w_out, m.b_out = _p(self.w_in), _p(self.b_in), _p(self.w_out), _p(self.b_out)
        m.d_hidden, m.d_in = self.w_in.shape
        m.d_out = self.w_out.shape[0]
        return m

    @property
    def d_out(self) -> int:
        return self.w_out.shape[0]


# ----------------------------------------------------------------------------------------------
def project(K, w2c, xyz):
    """pinhole.py:40-112 for one view -> (xy[N,2] unclamped, z[N], invalid[N] bool)."""
    K, w2c, xyz = _f32(K), _f32(w2c), _f32(xyz)
    N = xyz.shape[0]
    xy = np.empty((N, 2), np.float32); z = np.empty(N, np.float32); inv = np.empty(N, np.uint8)
    lib().sdo_project(_p(K), _p(w2c), _p(xyz), C.c_int64(N), _p(xy), _p(z), _p(inv))
    return xy, z, inv.astype(bool)


def encode_xyz(scene: Scene, xy_clamped, z):
    xy_clamped, z = _f32(xy_clamped), _f32(z)
    N = z.shape[0]
    s = scene.c()
    code = np.empty((N, scene.code_dim), np.float32)
    lib().sdo_encode_xyz(C.byref(s), _p(xy_clamped), _p(z), C.c_int64(N), _p(code))
    return code


def grid_sample(src, xy):
    """F.grid_sample(bilinear, border, align_corners=False); src [C,H,W], xy [N,2] -> [N,C]."""
    src, xy = _f32(src), _f32(xy)
    Cc, H, W = src.shape
    N = xy.shape[0]
    out = np.empty((N, Cc), np.float32)
    lib().sdo_grid_sample(_p(src), Cc, H, W, _p(xy), C.c_int64(N), _p(out))
    return out


def sample_features(scene: Scene, xyz):
    """BTSNet.sample_features (bts.py:271-328) -> (feat[N,nv_f,C+code], invalid[N,nv_f] bool)."""
    xyz = _f32(xyz); N = xyz.shape[0]
    s = scene.c()
    d = s.C + scene.code_dim
    out = np.empty((N, s.nv_f, d), np.float32); inv = np.empty((N, s.nv_f), np.uint8)
    lib().sdo_sample_features(C.byref(s), _p(xyz), C.c_int64(N), _p(out), _p(inv))
    return out, inv.astype(bool)


def mlp_forward(mlp: Mlp, x):
    x = _f32(x); N = x.shape[0]
    m = mlp.c()
    out = np.empty((N, m.d_out), np.float32)
    lib().sdo_mlp_forward(C.byref(m), _p(x), C.c_int64(N), _p(out))
    return out


def query_points(scene: Scene, mlp: Mlp, xyz, want_rgb=True, want_raw=False):
    """BTSNet.forward (bts.py:476-595) -> dict(sigma[N], dino[N,D], rgb[N,3nv_c], invalid[N,nv_c],
    invalid_features[N] bool[, raw[N,d_out]])."""
    xyz = _f32(xyz); N = xyz.shape[0]
    s, m = scene.c(), mlp.c()
    D = m.d_out - 1
    sigma = np.empty(N, np.float32); dino = np.empty((N, D), np.float32)
    invf = np.empty(N, np.uint8)
    rgb = inv = None
    if want_rgb and s.nv_c > 0:
        rgb = np.empty((N, 3 * s.nv_c), np.float32); inv = np.empty((N, s.nv_c), np.float32)
    raw = np.empty((N, m.d_out), np.float32) if want_raw else None
    rc = lib().sdo_query_points(C.byref(s), C.byref(m), _p(xyz), C.c_int64(N), _p(sigma), _p(dino),
                                _p(rgb), _p(inv), _p(invf), _p(raw))
    if rc != 0:
        raise ValueError("oracle: unsupported configuration (nv_f must be 1, d_in = C + code)")
    out = dict(sigma=sigma, dino=dino, rgb=rgb, invalid=inv, invalid_features=invf.astype(bool))
    if want_raw:
        out["raw"] = raw
    return out


def sample_coarse(rays, u, lin, lindisp=True):
    rays, u, lin = _f32(rays), _f32(u), _f32(lin)
    R, Kc = u.shape
    z = np.empty((R, Kc), np.float32)
    lib().sdo_sample_coarse(_p(rays), C.c_int64(R), rays.shape[1], _p(u), _p(lin), Kc, int(lindisp), _p(z))
    return z


def sample_fine(rays, weights, u0, u1, lindisp=True):
    rays, weights, u0, u1 = _f32(rays), _f32(weights), _f32(u0), _f32(u1)
    R, Kc = weights.shape
    Kf = u0.shape[1]
    z = np.empty((R, Kf), np.float32); inds = np.empty((R, Kf), np.int32)
    lib().sdo_sample_fine(_p(rays), C.c_int64(R), rays.shape[1], _p(weights), Kc, _p(u0), _p(u1), Kf,
                          int(lindisp), _p(z), _p(inds))
    return z, inds


def sample_fine_depth(rays, depth, noise, depth_std):
    rays, depth, noise = _f32(rays), _f32(depth), _f32(noise)
    R, Kfd = noise.shape
    z = np.empty((R, Kfd), np.float32)
    lib().sdo_sample_fine_depth(_p(rays), C.c_int64(R), rays.shape[1], _p(depth), _p(noise), Kfd,
                                C.c_float(depth_std), _p(z))
    return z


def sample_coarse_from_dist(weights, z_samp, u0, u1, lindisp=True):
    weights, z_samp, u0, u1 = _f32(weights), _f32(z_samp), _f32(u0), _f32(u1)
    R, Kp = weights.shape
    Kc = u0.shape[1]
    z = np.empty((R, Kc), np.float32); inds = np.empty((R, Kc), np.int32)
    lib().sdo_sample_coarse_from_dist(C.c_int64(R), _p(weights), _p(z_samp), Kp, _p(u0), _p(u1), Kc,
                                      int(lindisp), _p(z), _p(inds))
    return z, inds


def sort_rows(z):
    z = _f32(z).copy()
    lib().sdo_sort_rows(_p(z), C.c_int64(z.shape[0]), z.shape[1])
    return z


def composite(z, sigma, feat, rgb, hard_alpha_cap=False, white_bkgd=False):
    z, sigma, feat = _f32(z), _f32(sigma), _f32(feat)
    R, K = z.shape
    D = feat.shape[-1]
    Crgb = 0 if rgb is None else rgb.shape[-1]
    rgb = None if rgb is None else _f32(rgb)
    w = np.empty((R, K), np.float32); a = np.empty((R, K), np.float32)
    depth = np.empty(R, np.float32); dino = np.empty((R, D), np.float32)
    rgb_out = np.empty((R, Crgb), np.float32) if Crgb else None
    lib().sdo_composite(_p(z), _p(sigma), _p(feat), _p(rgb), C.c_int64(R), K, D, Crgb,
                        int(hard_alpha_cap), int(white_bkgd), _p(w), _p(a), _p(depth), _p(dino),
                        _p(rgb_out))
    return dict(weights=w, alphas=a, depth=depth, dino=dino, rgb=rgb_out)


def render_pass(scene: Scene, mlp: Mlp, rays, z, hard_alpha_cap=False, white_bkgd=False,
                want_rgb_samps=False):
    """One NeRFRenderer.composite call (nerf.py:230-449) for one scene."""
    rays, z = _f32(rays), _f32(z)
    R, K = z.shape
    s, m = scene.c(), mlp.c()
    D = m.d_out - 1
    nv_c = s.nv_c
    w = np.empty((R, K), np.float32); a = np.empty((R, K), np.float32)
    depth = np.empty(R, np.float32); dino = np.empty((R, D), np.float32)
    rgb = np.empty((R, 3 * nv_c), np.float32)
    inv = np.empty((R, K, nv_c), np.float32); invf = np.empty((R, K), np.uint8)
    sig = np.empty((R, K), np.float32)
    rs = np.empty((R, K, 3 * nv_c), np.float32) if want_rgb_samps else None
    rc = lib().sdo_render_pass(C.byref(s), C.byref(m), _p(rays), C.c_int64(R), rays.shape[1], _p(z), K,
                               int(hard_alpha_cap), int(white_bkgd), _p(w), _p(a), _p(depth), _p(dino),
                               _p(rgb), _p(inv), _p(invf), _p(rs), _p(sig))
    if rc != 0:
        raise ValueError("oracle: unsupported configuration")
    out = dict(weights=w, alphas=a, depth=depth, dino_features=dino, rgb=rgb, invalid=inv,
               invalid_features=invf.astype(bool), z_samps=z, sigma=sig)
    if want_rgb_samps:
        out["rgb_samps"] = rs
    return out


def render_rays(scene: Scene, mlp: Mlp, rays, *, lin, u_coarse, u_fine0=None, u_fine1=None,
                n_depth=None, depth_std=0.01, lindisp=True, hard_alpha_cap=False, white_bkgd=False):
    """NeRFRenderer.forward for one scene (nerf.py:451-539): coarse pass and, when fine uniforms
    are given, importance/depth samples merged and a second pass."""
    z_c = sample_coarse(rays, u_coarse, lin, lindisp)
    coarse = render_pass(scene, mlp, rays, z_c, hard_alpha_cap, white_bkgd)
    out = {"coarse": coarse}
    parts = [z_c]
    if u_fine0 is not None and u_fine0.shape[1] > 0:
        z_f, inds = sample_fine(rays, coarse["weights"], u_fine0, u_fine1, lindisp)
        parts.append(z_f); out["fine_inds"] = inds
    if n_depth is not None and n_depth.shape[1] > 0:
        parts.append(sample_fine_depth(rays, coarse["depth"], n_depth, depth_std))
    if len(parts) > 1:
        z_all = sort_rows(np.concatenate(parts, axis=1))
        out["fine"] = render_pass(scene, mlp, rays, z_all, hard_alpha_cap, white_bkgd)
    return out


def expand_dim(f, w1, b1, w2, b2):
    """MlpDimReduction.transform_expand (dim_reduction.py:22-25)."""
    f, w1, b1, w2, b2 = _f32(f), _f32(w1), _f32(b1), _f32(w2), _f32(b2)
    N = f.shape[0]
    out = np.empty((N, w2.shape[0]), np.float32)
    lib().sdo_expand_dim(_p(f), C.c_int64(N), w1.shape[1], w1.shape[0], w2.shape[0], _p(w1), _p(b1),
                         _p(w2), _p(b2), _p(out))
    return out


def gen_rays(c2w, proj, H, W, z_near, z_far, frame_ids=None, norm_dir=True, xy_shift=(0.0, 0.0)):
    """ImageRaySampler.sample's rays for one batch element (ray_sampler.py:439-486, util.py:113-158,253-285):
    c2w [V,4,4], proj [V,3,3] -> [V*H*W, 11]."""
    c2w, proj = _f32(c2w), _f32(proj)
    V = c2w.shape[0]
    ids = None if frame_ids is None else _f32(frame_ids)
    out = np.empty((V * H * W, 11), np.float32)
    lib().sdo_gen_rays(_p(c2w), _p(proj), _p(ids) if ids is not None else None, V, H, W, C.c_float(z_near),
                       C.c_float(z_far), int(bool(norm_dir)), C.c_float(xy_shift[0]), C.c_float(xy_shift[1]), _p(out))
    return out


def ssc_head(x, wl, bl, wn1, bn1, wn2, bn2, centres, lut):
    """SemanticHead.forward(mode="stego_kmeans") (semantic_head.py:107-112, 285-305, 308-373):
    x [N, d_in] -> (seg [N] int64, pseudo [N] int64, inner products [N, n_cls])."""
    x, wl, bl, wn1, bn1, wn2, bn2, centres = (_f32(a) for a in (x, wl, bl, wn1, bn1, wn2, bn2, centres))
    lut = np.ascontiguousarray(lut, dtype=np.int64)
    N, d_in = x.shape
    d_mid, d_code, n_cls = wn1.shape[0], wl.shape[0], centres.shape[0]
    seg, pseudo = np.empty(N, np.int64), np.empty(N, np.int64)
    ip = np.empty((N, n_cls), np.float32)
    rc = lib().sdo_ssc_head(_p(x), C.c_int64(N), d_in, d_mid, d_code, n_cls, _p(wl), _p(bl), _p(wn1), _p(bn1), _p(wn2), _p(bn2),
                            _p(centres), _p(lut), _p(seg), _p(pseudo), _p(ip))
    if rc:
        raise MemoryError("sdo_ssc_head")
    return seg, pseudo, ip


def voxel_grid(T, dims=(256, 256, 32), voxel_size=0.2, origin=(0.0, -25.6, -2.0), x_range=None):
    """The SSC voxel grid in the camera frame as the reference builds it (sscbench/evaluate_model_sscbench.py:270-278,
    point_utils.py:46-67, fusion.py:205-219, 407-411): [N, 3] fp32, 'ij' order.  ``T``: [3..4, 4] float64 lidar -> camera."""
    x0, x1 = (0, dims[0]) if x_range is None else x_range
    Th = np.ascontiguousarray(np.asarray(T, np.float64)[:3, :4])
    org = np.ascontiguousarray(np.asarray(origin, np.float32))
    out = np.empty(((x1 - x0) * dims[1] * dims[2], 3), np.float32)
    lib().sdo_voxel_grid(_p(org), C.c_double(float(voxel_size)), int(dims[0]), int(dims[1]), int(dims[2]), int(x0), int(x1),
                         Th.ctypes.data_as(C.c_void_p), _p(out))
    return out
