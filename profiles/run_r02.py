"""Small single-purpose workloads for the round-2 ncu captures (one mode per process, a handful of launches each):
    python profiles/run_r02.py query|binned|ssc|render64|render768|expand [n_iter]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedino_b200 import ops  # noqa: E402
from scenedino_b200 import synthetic as syn  # noqa: E402


def main():
    mode = sys.argv[1]
    n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(1)
    Kc = syn.kitti360_K()
    if mode in ("query", "ssc", "binned", "x3"):
        N = 1 << 21
        if mode == "ssc":
            f = torch.randn((N, 64), device=dev, generator=g) * 0.5
            head = ops.SscHead(syn.make_expand(3), syn.make_ssc_head(21), device=dev)
            o = dict(seg=torch.empty((N,), dtype=torch.uint8, device=dev))
            for _ in range(n_iter):
                ops.ssc_head(head, f, want_scores=False, out=o)
        else:
            feat = ops.featmap_pack(torch.randn((1, 256, 384, 1280), device=dev, generator=g), torch.float32 if mode == "x3" else torch.float16)
            sc = ops.Scene(feat=feat[0], K_f=torch.from_numpy(Kc[None]).to(dev), w2c_f=torch.eye(4, device=dev)[None])
            mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F32TC if mode == "x3" else ops.F16)
            sc = sc.project_x3(mlp) if mode == "x3" else sc.project(mlp)
            pts = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
            out = None
            for _ in range(n_iter):
                if mode == "binned":
                    r = ops.query_points_binned(sc, mlp, pts, out=out)
                    if out is None:
                        out = dict(r); out["invalid_features"] = out["invalid_features"].view(torch.uint8)
                elif mode == "x3":
                    out = ops.query_points(sc, mlp, pts, want_rgb=False, precision=ops.F32TC, out=out)
                else:
                    out = ops.query_points(sc, mlp, pts, want_rgb=False, out=out)
    elif mode == "expand":
        N = 1 << 19
        f = torch.randn((N, 64), device=dev, generator=g) * 0.5
        mlp = ops.Mlp(*syn.make_expand(3), device=dev)
        for _ in range(n_iter):
            ops.expand_dim(mlp, f, precision=ops.F16)
    elif mode in ("render64", "render768"):
        D, K, nv, (Hf, Wf) = (64, 32, 1, (192, 640)) if mode == "render64" else (768, 96, 4, (384, 1280))
        fm = ops.featmap_pack(torch.randn((1, 256, Hf, Wf), device=dev, generator=g), torch.float16)
        c2w = np.stack([syn.view_pose_c2w(v) for v in range(nv)])
        w2c = np.linalg.inv(c2w.astype(np.float64)).astype(np.float32)
        sc = ops.Scene(feat=fm[0], K_f=torch.from_numpy(Kc[None]).to(dev), w2c_f=torch.from_numpy(w2c[:1]).to(dev),
                       rgb=torch.from_numpy(syn.make_images(2, nv)).to(dev),
                       K_c=torch.from_numpy(np.broadcast_to(Kc, (nv, 3, 3)).copy()).to(dev), w2c_c=torch.from_numpy(w2c).to(dev))
        mlp = ops.Mlp(*syn.make_mlp(0, d_out=D + 1), device=dev, precision=ops.F16)
        sc = sc.project(mlp)
        R = syn.IMG_H * syn.IMG_W
        view = torch.from_numpy(syn.view_pose_c2w(1).astype(np.float32)).to(dev)[None]
        rays = ops.gen_rays(view, torch.from_numpy(Kc[None].astype(np.float32)).to(dev), syn.IMG_H, syn.IMG_W, syn.Z_NEAR, syn.Z_FAR)
        lin = torch.linspace(0, 1 - 1.0 / K, K, device=dev)
        u = torch.rand((R, K), device=dev, generator=g)
        z = torch.sort(ops.sample_coarse(rays, u, lin, True), dim=1).values.contiguous()
        ro = None
        for _ in range(n_iter):
            ro = ops.render_pass(sc, mlp, rays, z, per_sample=False, out=ro)
    torch.cuda.synchronize()
    print("ok", mode)


if __name__ == "__main__":
    main()
