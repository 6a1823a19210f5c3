"""Round-2 device timings (CUDA events, warm, median): the fused expansion + SSC head, the expansion alone, the field query
feeding them, and full-image renders for D = 64 (both tensor-core composites) and D = 768 (hidden composite + head2).
    python profiles/time_r02.py > gpurun_out/time_r02.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedino_b200 import ops  # noqa: E402
from scenedino_b200 import synthetic as syn  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    dev = "cuda:0"
    out = {"env": {k: os.environ.get(k) for k in ("SD_TC_HCOMP",)}}
    N = 1 << 21
    f = torch.randn((N, 64), device=dev) * 0.5
    expand, hw = syn.make_expand(3), syn.make_ssc_head(21)
    head = ops.SscHead(expand, hw, device=dev)
    o = dict(seg=torch.empty((N,), dtype=torch.uint8, device=dev))
    ms = timed(lambda: ops.ssc_head(head, f, want_scores=False, out=o))
    out["ssc_head_2M"] = {"ms": ms, "Gvox_s": N / ms / 1e6, "exec_tflops": N * 364544 / ms / 1e9,
                          "ref_algo_tflops": N * (212992 + 1376256 + 2 * 64 * 27) / ms / 1e9}
    mlp_e = ops.Mlp(*expand, device=dev)
    ms = timed(lambda: ops.expand_dim(mlp_e, f[: N // 4], precision=ops.F16), n=5)
    out["expand_tc_512k"] = {"ms": ms}
    del f
    torch.cuda.empty_cache()
    # renders
    g = torch.Generator(device=dev).manual_seed(1)
    for name, (Hf, Wf, K, D, nv) in {"vitb8_K64_D64": (384, 1280, 64, 64, 1), "dinov2_K32_D64": (192, 640, 32, 64, 1),
                                     "vitb8_K96_D768_nv4": (384, 1280, 96, 768, 4), "vitb8_K64_D768_nv4": (384, 1280, 64, 768, 4)}.items():
        feat = torch.randn((1, 256, Hf, Wf), device=dev, generator=g)
        fm = ops.featmap_pack(feat, torch.float16)
        del feat
        Kc = syn.kitti360_K()
        c2w = np.stack([syn.view_pose_c2w(v) for v in range(nv)])
        w2c = np.linalg.inv(c2w.astype(np.float64)).astype(np.float32)
        imgs = syn.make_images(2, nv)
        sc = ops.Scene(feat=fm[0], K_f=torch.from_numpy(Kc[None]).to(dev), w2c_f=torch.from_numpy(w2c[:1]).to(dev),
                       rgb=torch.from_numpy(imgs).to(dev), K_c=torch.from_numpy(np.broadcast_to(Kc, (nv, 3, 3)).copy()).to(dev),
                       w2c_c=torch.from_numpy(w2c).to(dev))
        mlp = ops.Mlp(*syn.make_mlp(0, d_out=D + 1), device=dev, precision=ops.F16)
        sc = sc.project(mlp)
        R = syn.IMG_H * syn.IMG_W
        view = torch.from_numpy(syn.view_pose_c2w(1).astype(np.float32)).to(dev)[None]
        rays = ops.gen_rays(view, torch.from_numpy(Kc[None].astype(np.float32)).to(dev), syn.IMG_H, syn.IMG_W, syn.Z_NEAR, syn.Z_FAR)
        lin = torch.linspace(0, 1 - 1.0 / K, K, device=dev)
        u = torch.rand((R, K), device=dev, generator=g)
        z = torch.sort(ops.sample_coarse(rays, u, lin, True), dim=1).values.contiguous()
        ro = None

        def step():
            nonlocal ro
            ro = ops.render_pass(sc, mlp, rays, z, per_sample=False, out=ro)

        ms = timed(step, n=8)
        flop = 2 * (295 * 128 + 128 * (D + 1))
        out["render_" + name] = {"ms": ms, "Msamples_s": R * K / ms / 1e3, "algo_tflops": R * K * flop / ms / 1e9}
        del sc, fm, mlp, rays, u, z, ro
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
