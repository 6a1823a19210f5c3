"""Synthetic KITTI-360-shaped inputs for tests and bench.py (there is no dataset / checkpoint access).

Shapes and constants follow the reference's inference set-up:
  * normalised intrinsics           demo_utils/utils.py:102-105
  * z_near / z_far = 3 / 80         demo_utils/utils.py:49
  * image 192 x 640, feature map 256 x 384 x 1280 (ViT-B/8) or 256 x 192 x 640 (DINOv2 B/14)
  * rays = origin(3) dir(3, unit) near far frame_id px py   common/ray_sampler.py:476-484
  * SSC grid 256 x 256 x 32 @ 0.2 m, origin (0,-25.6,-2) lidar frame, moved to the camera frame
    by the hard-coded KITTI-360 calibration    sscbench/point_utils.py:17-82,84-137

Everything is generated from numpy ``RandomState`` seeds so that fixtures only need to store
outputs.  Pure numpy: usable on the CPU box and the GPU box alike.
"""
from __future__ import annotations

import math

import numpy as np

Z_NEAR, Z_FAR = 3.0, 80.0
IMG_H, IMG_W = 192, 640


def kitti360_K() -> np.ndarray:
    return np.array([[0.7849, 0.0, -0.0312], [0.0, 2.9391, 0.2701], [0.0, 0.0, 1.0]], np.float32)


def view_pose_c2w(v: int) -> np.ndarray:
    """View 0 is the identity; extra views move 0.6 m right on odd v, 1 m forward per pair and yaw
    by +-5 degrees per pair."""
    T = np.eye(4, dtype=np.float64)
    if v > 0:
        yaw = math.radians(5.0) * (v // 2) * (1 if v % 2 == 0 else -1)
        c, s = math.cos(yaw), math.sin(yaw)
        T[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        T[:3, 3] = [0.6 * (v % 2), 0.0, 1.0 * (v // 2)]
    return T.astype(np.float32)


def make_mlp(seed: int = 0, d_in: int = 295, d_hidden: int = 128, d_out: int = 65,
             bias_scale: float = 0.0):
    """kaiming_normal_(fan_in) weights as in resnetfc.py:90-96; the reference zero-inits biases,
    ``bias_scale`` > 0 draws them instead so that the bias path is exercised."""
    rs = np.random.RandomState(seed)
    w_in = (rs.standard_normal((d_hidden, d_in)) * math.sqrt(2.0 / d_in)).astype(np.float32)
    w_out = (rs.standard_normal((d_out, d_hidden)) * math.sqrt(2.0 / d_hidden)).astype(np.float32)
    b_in = (rs.standard_normal(d_hidden) * bias_scale).astype(np.float32)
    b_out = (rs.standard_normal(d_out) * bias_scale).astype(np.float32)
    return w_in, b_in, w_out, b_out


def make_expand(seed: int = 3, d_red: int = 64, d_lat: int = 128, d_full: int = 768):
    """nn.Linear default-init-like weights for MlpDimReduction (dim_reduction.py:15-20)."""
    rs = np.random.RandomState(seed)
    b1, b2 = 1.0 / math.sqrt(d_red), 1.0 / math.sqrt(d_lat)
    w1 = rs.uniform(-b1, b1, (d_lat, d_red)).astype(np.float32)
    c1 = rs.uniform(-b1, b1, d_lat).astype(np.float32)
    w2 = rs.uniform(-b2, b2, (d_full, d_lat)).astype(np.float32)
    c2 = rs.uniform(-b2, b2, d_full).astype(np.float32)
    return w1, c1, w2, c2


def make_feature_map(seed: int, C: int, Hf: int, Wf: int) -> np.ndarray:
    """[1, C, Hf, Wf] standard-normal stand-in for the encoder output (NCHW like the reference)."""
    rs = np.random.RandomState(seed)
    return rs.standard_normal((1, C, Hf, Wf)).astype(np.float32)


def make_images(seed: int, nv: int, H: int = IMG_H, W: int = IMG_W) -> np.ndarray:
    """[nv, 3, H, W] in [0,1] -- what BTSNet.encode stores as grid_c_imgs (images*0.5+0.5)."""
    rs = np.random.RandomState(seed)
    return rs.uniform(0.0, 1.0, (nv, 3, H, W)).astype(np.float32)


def image_rays(c2w: np.ndarray, K: np.ndarray, H: int = IMG_H, W: int = IMG_W,
               z_near: float = Z_NEAR, z_far: float = Z_FAR, frame_id: float = 0.0) -> np.ndarray:
    """All pixel-centre rays of one view, [H*W, 11] (ImageRaySampler.sample, ray_sampler.py:439-513
    via util.gen_rays / unproj_map, util.py:113-158,253-285)."""
    c2w = c2w.astype(np.float32); K = K.astype(np.float32)
    pw, ph = np.float32(2.0 / W), np.float32(2.0 / H)
    xs = np.linspace(-1 + 0.5 * pw, 1 - 0.5 * pw, W, dtype=np.float32)
    ys = np.linspace(-1 + 0.5 * ph, 1 - 0.5 * ph, H, dtype=np.float32)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    x = (X - K[0, 2]) / K[0, 0]
    y = (Y - K[1, 2]) / K[1, 1]
    d = np.stack([x, y, np.ones_like(x)], -1).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    dw = (d.reshape(-1, 3) @ c2w[:3, :3].T).astype(np.float32)
    n = H * W
    rays = np.empty((n, 11), np.float32)
    rays[:, 0:3] = c2w[:3, 3]
    rays[:, 3:6] = dw
    rays[:, 6] = z_near
    rays[:, 7] = z_far
    rays[:, 8] = frame_id
    rays[:, 9] = X.reshape(-1)
    rays[:, 10] = Y.reshape(-1)
    return rays


_CAM2VELO = np.array([
    0.04307104361, -0.08829286498, 0.995162929, 0.8043914418,
    -0.999004371, 0.007784614041, 0.04392796942, 0.2993489574,
    -0.01162548558, -0.9960641394, -0.08786966659, -0.1770225824], np.float64).reshape(3, 4)


def velo_to_cam() -> np.ndarray:
    c2v = np.concatenate([_CAM2VELO, np.array([[0, 0, 0, 1.0]])], 0)
    return np.linalg.inv(c2v)


def ssc_voxel_grid(dims=(256, 256, 32), voxel_size: float = 0.2, origin=(0.0, -25.6, -2.0),
                   x_range: tuple[int, int] | None = None) -> np.ndarray:
    """Voxel centres of the SSCBench grid in the camera frame, flattened 'ij' order, [N,3] fp32.
    ``x_range`` selects a slab of x indices (used for multi-GPU voxel-slab sharding)."""
    x0, x1 = (0, dims[0]) if x_range is None else x_range
    org = np.asarray(origin, np.float32).astype(np.float64)      # vol_origin.astype(np.float32), fusion.py:207
    vs = float(voxel_size)                                        # a Python float in the reference: a double inside numba

    def centres(o, idx):
        # vox2world (sscbench/fusion.py:205-219, numba): vol_origin[j] + (vox_size * vox_coords[i, j]) + vox_size * offsets[j]
        # with fp32 origin / index and double size / offset -> double arithmetic, ONE rounding when stored to the fp32 array
        return ((o + vs * idx.astype(np.float32).astype(np.float64)) + vs * 0.5).astype(np.float32)

    px = centres(org[0], np.arange(x0, x1))
    py = centres(org[1], np.arange(dims[1]))
    pz = centres(org[2], np.arange(dims[2]))
    P = np.stack(np.meshgrid(px, py, pz, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    T = velo_to_cam()
    Ph = np.hstack([P, np.ones((len(P), 1), np.float32)])  # rigid_transform (fusion.py:407-411)
    out = (T @ Ph.T).T[:, :3]
    return np.ascontiguousarray(out, dtype=np.float32)


def random_points(seed: int, n: int, box=((-30, 30), (-6, 6), (-5, 70))) -> np.ndarray:
    """Points around the camera: a good share out of frustum and some behind the camera."""
    rs = np.random.RandomState(seed)
    lo = np.array([b[0] for b in box], np.float32); hi = np.array([b[1] for b in box], np.float32)
    return (rs.uniform(0, 1, (n, 3)).astype(np.float32) * (hi - lo) + lo).astype(np.float32)


def make_ssc_head(seed: int, d_in: int = 768, d_code: int = 64, n_cls: int = 27, gt_cls: int = 19) -> dict:
    """Seeded parameters of the unsupervised SSC head (downstream_head/semantic_head.py: StegoClusterHead 1x1-conv
    weights at nn.Conv2d's default scale, KMeansParamHead centres and pseudo-label LUT)."""
    rs = np.random.RandomState(seed)

    def u(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rs.uniform(-b, b, shape).astype(np.float32)

    return dict(wl=u((d_code, d_in), d_in), bl=u((d_code,), d_in), wn1=u((d_in, d_in), d_in), bn1=u((d_in,), d_in),
                wn2=u((d_code, d_in), d_in), bn2=u((d_code,), d_in), centres=rs.randn(n_cls, d_code).astype(np.float32),
                lut=rs.randint(0, gt_cls, n_cls).astype(np.int64))


def ssc_head_inputs(dino_full: np.ndarray, dino_full_le: np.ndarray) -> np.ndarray:
    """Rows fed to the head in tests/golden/ssc_head.npz: the reference's own 768-d expansions (unit rows), some of them
    scaled (the head re-normalises), and one all-zero row (norm clamp)."""
    x = np.concatenate([dino_full, dino_full_le], 0).astype(np.float32)
    return np.concatenate([x, (x[:64] * np.float32(3.5)).astype(np.float32), np.zeros((1, x.shape[1]), np.float32)], 0)
