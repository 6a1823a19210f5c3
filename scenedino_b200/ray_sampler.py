"""Whole-image ray sampler on the device.

Mirror of the reference's ``ImageRaySampler`` (scenedino/common/ray_sampler.py:421-607): same constructor,
``sample`` and ``reconstruct`` signatures, same tensor shapes, same attribute side effects (``channels``, ``height``
and ``width`` are learnt from the first images seen).  The reference builds the rays per batch element from a
dozen eager ops (util.gen_rays / util.unproj_map, common/util.py:113-158,253-285); here ONE kernel launch
(``sd_gen_rays``) writes the ``[n, v*H*W, 11]`` tensor for the whole batch from the poses and intrinsics, bit-identical
to torch's CPU result.  ``reconstruct`` only reshapes (views), as in the reference.
"""
from __future__ import annotations

from math import isqrt

import torch

from . import ops

# render-dict entries and the trailing dimensions they keep after [n, v_in, H, W]
#   "K" = samples per ray, "V" = rendered colour views, "C" = channels, "*" = the entry's own last dimension
_PER_RAY_LAYOUT = {
    "rgb": ("V", "C"),
    "weights": ("K",),
    "depth": (),
    "invalid": ("K", "V"),
    "invalid_features": ("K", "V"),
    "alphas": ("K",),
    "z_samps": ("K",),
    "rgb_samps": ("K", "V", "C"),
    "ray_info": ("*",),
    "extras": ("*",),
    "dino_features": (1, "*"),
}


class RaySampler:
    """ray_sampler.py:11-20."""

    def __init__(self, z_near: float, z_far: float) -> None:
        self.z_near = z_near
        self.z_far = z_far

    def sample(self, images, poses, projs):
        raise NotImplementedError

    def reconstruct(self, render_dict):
        raise NotImplementedError


def _rows_channels_last(x: torch.Tensor) -> torch.Tensor:
    """[n, v, c, h, w] -> [n, v*h*w, c] (the ground-truth layout next to the rays, ray_sampler.py:488-502)."""
    n, v, c, h, w = x.shape
    return x.permute(0, 1, 3, 4, 2).reshape(n, v * h * w, c)


class ImageRaySampler(RaySampler):
    def __init__(self, z_near: float, z_far: float, height: int | None = None, width: int | None = None,
                 channels: int = 3, norm_dir: bool = True, dino_upscaled: bool = False) -> None:
        super().__init__(z_near, z_far)
        self.height = height
        self.width = width
        self.channels = channels
        self.norm_dir = norm_dir
        self.dino_upscaled = dino_upscaled

    def sample(self, images, poses, projs, image_ids=None, dino_features=None, dino_artifacts=None):
        """images [n,v,c,h,w] | None, poses [n,v,4,4] (CUDA), projs [n,v,3,3] -> (rays [n, v*H*W, 11], rgb_gt
        [n, v*H*W, c] | None[, dino_gt [n, v*ph*pw, dc]])."""
        n, v = poses.shape[:2]
        if images is not None:
            self.channels = images.shape[2]
        if self.height is None:
            self.height, self.width = images.shape[-2:]
        H, W = self.height, self.width

        if image_ids is None:
            ids = torch.arange(v, device=poses.device, dtype=torch.float32)
        else:
            ids = torch.as_tensor(image_ids, device=poses.device, dtype=torch.float32)
        # batch elements are just more views to the kernel: [n*v] poses -> [n*v*H*W, 11]
        rays = ops.gen_rays(poses.reshape(n * v, 4, 4), projs.reshape(n * v, 3, 3), H, W, float(self.z_near),
                            float(self.z_far), frame_ids=ids.repeat(n), norm_dir=self.norm_dir)
        rays = rays.view(n, v * H * W, 11).to(poses.dtype)

        rgb_gt = None
        if images is not None:
            rgb_gt = _rows_channels_last(images.reshape(n, -1, self.channels, H, W))
        if dino_features is not None:
            dc, ph, pw = dino_features.shape[2:]
            return rays, rgb_gt, _rows_channels_last(dino_features.reshape(n, -1, dc, ph, pw))
        return rays, rgb_gt

    def reconstruct(self, render_dict, channels=None, dino_channels=None):
        """Views every per-ray entry of the render dict(s) back as images (ray_sampler.py:515-607)."""
        H, W = self.height, self.width
        n = v_in = None
        for name, part in render_dict.items():
            if not isinstance(part, dict) or "rgb" not in part:
                continue
            if channels is None:
                channels = self.channels
            n, n_rays, v_c = part["rgb"].shape
            v_in = n_rays // (H * W)
            dims = {"V": v_c // channels, "C": channels, "K": part["weights"].shape[-1]}
            for key, tail in _PER_RAY_LAYOUT.items():
                if key not in part:
                    continue
                t = part[key]
                shape = [t.shape[-1] if d == "*" else dims.get(d, d) for d in tail]
                part[key] = t.view(n, v_in, H, W, *shape)
            render_dict[name] = part

        if "rgb_gt" in render_dict:
            render_dict["rgb_gt"] = render_dict["rgb_gt"].view(n, v_in, H, W, channels)
        if "dino_gt" in render_dict:
            gt = render_dict["dino_gt"]
            dc = gt.shape[-1]
            if self.dino_upscaled:
                render_dict["dino_gt"] = gt.view(n, v_in, H, W, dc)
            else:
                # the reference infers the patch size from the element count (ray_sampler.py:590-594)
                ps = isqrt((n * v_in * H * W * dc) // gt.numel())
                render_dict["dino_gt"] = gt.view(n, v_in, H // ps, W // ps, dc)
            if "dino_artifacts" in render_dict:
                art = render_dict["dino_artifacts"]
                ps = isqrt((n * v_in * H * W * dc) // art.numel())
                render_dict["dino_artifacts"] = art.view(n, v_in, H // ps, W // ps, dc)
        return render_dict
