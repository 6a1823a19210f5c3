"""Kernel-level breakdown of one bench render configuration (torch.profiler / CUPTI sees the library's kernels too):
    python profiles/dev_cfg.py cfg4|cfg1|cfg3"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import scenedino_b200 as sd  # noqa: E402
from scenedino_b200 import synthetic as syn  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(1)
    Hf, Wf, views, n_coarse, n_fine, d_out, nv_c, subset, expand = {
        "cfg1": (384, 1280, [0], 64, 0, 65, 1, 4096, False),
        "cfg4": (192, 640, [1], 32, 0, 65, 1, None, True),
        "cfg3": (384, 1280, [0, 1, 2, 3], 64, 32, 769, 4, None, False)}[tag]
    hold = {"map": torch.randn((1, 256, Hf, Wf), device=dev, generator=g)}
    rnet = bench.build_net(sd, torch, hold, dev, "fp16", d_out=d_out, with_head=False, seed=0)
    ren = sd.NeRFRenderer.from_conf({"n_coarse": n_coarse, "n_fine": n_fine, "n_fine_depth": 0, "lindisp": True,
                                     "hard_alpha_cap": n_fine > 0})
    ren.nan_check = False
    wrapped = ren.bind_parallel(rnet, gpus=None).eval()
    imgs = torch.from_numpy(syn.make_images(2, nv_c)).to(dev)[None]
    Kc = torch.from_numpy(np.broadcast_to(syn.kitti360_K(), (nv_c, 3, 3)).copy()).to(dev)[None]
    c2w = torch.from_numpy(np.stack([syn.view_pose_c2w(v) for v in range(nv_c)])).to(dev)[None]
    rnet.encoder.dim_reduction.precision = "fp16"
    rnet.encode(imgs * 2 - 1, Kc, c2w, ids_encoder=[0], ids_render=list(range(nv_c)), images_alt=imgs)
    rnet.set_scale(0)
    vp = torch.from_numpy(np.stack([syn.view_pose_c2w(v) for v in views])).to(dev)[None]
    vK = torch.from_numpy(np.broadcast_to(syn.kitti360_K(), (len(views), 3, 3)).copy()).to(dev)[None]
    sampler = sd.ImageRaySampler(z_near=syn.Z_NEAR, z_far=syn.Z_FAR, height=syn.IMG_H, width=syn.IMG_W)
    sel = torch.randperm(len(views) * syn.IMG_H * syn.IMG_W, device=dev, generator=g)[:subset] if subset else None

    def step():
        rays, _ = sampler.sample(None, vp, vK)
        if sel is not None:
            rays = rays[:, sel].contiguous()
        with torch.no_grad():
            out = wrapped(rays)
        lvl = out["fine"] if n_fine > 0 else out["coarse"]
        if expand:
            return rnet.encoder.expand_dim(lvl["dino_features"])
        return lvl

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        step()
    b.record()
    torch.cuda.synchronize()
    print(tag, "ms/step", a.elapsed_time(b) / 5)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
