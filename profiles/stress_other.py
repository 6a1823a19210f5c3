"""Stress of the other tcgen05 kernels' issue protocol: the fused SSC head, the expansion, render passes (64-d and 768-d),
the x3 tile kernel -- thousands of launches each, results compared with the first launch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedino_b200 import ops, synthetic as syn
dev = "cuda"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
g = torch.Generator(device=dev).manual_seed(1)
N = 1 << 19
f = torch.randn((N, 64), device=dev, generator=g) * 0.5
head = ops.SscHead(syn.make_expand(3), syn.make_ssc_head(21), device=dev)
o = dict(seg=torch.empty((N,), dtype=torch.uint8, device=dev))
ops.ssc_head(head, f, want_scores=False, out=o); ref = o["seg"].clone()
bad = 0
for i in range(n):
    ops.ssc_head(head, f, want_scores=False, out=o)
    if i % 500 == 499:
        torch.cuda.synchronize(); bad += int(not torch.equal(o["seg"], ref))
print("ssc_head", n, "launches, mismatches", bad, flush=True)
mlp_e = ops.Mlp(*syn.make_expand(3), device=dev)
e0 = ops.expand_dim(mlp_e, f[:1 << 17], precision=ops.F16).clone()
for i in range(n):
    e1 = ops.expand_dim(mlp_e, f[:1 << 17], precision=ops.F16)
torch.cuda.synchronize(); print("expand_tc", n, "launches, equal:", bool(torch.equal(e0, e1)), flush=True)
Kc = syn.kitti360_K()
for D, K in ((64, 32), (768, 64)):
    nv = 2
    fm = ops.featmap_pack(torch.randn((1, 256, 192, 640), device=dev, generator=g), torch.float16)
    c2w = np.stack([syn.view_pose_c2w(v) for v in range(nv)]); w2c = np.linalg.inv(c2w.astype(np.float64)).astype(np.float32)
    sc = ops.Scene(feat=fm[0], K_f=torch.from_numpy(Kc[None]).to(dev), w2c_f=torch.from_numpy(w2c[:1]).to(dev),
                   rgb=torch.from_numpy(syn.make_images(2, nv)).to(dev), K_c=torch.from_numpy(np.broadcast_to(Kc, (nv, 3, 3)).copy()).to(dev),
                   w2c_c=torch.from_numpy(w2c).to(dev))
    mlp = ops.Mlp(*syn.make_mlp(0, d_out=D + 1), device=dev, precision=ops.F16)
    sc = sc.project(mlp)
    view = torch.from_numpy(syn.view_pose_c2w(1).astype(np.float32)).to(dev)[None]
    rays = ops.gen_rays(view, torch.from_numpy(Kc[None].astype(np.float32)).to(dev), syn.IMG_H, syn.IMG_W, syn.Z_NEAR, syn.Z_FAR)[::4].contiguous()
    lin = torch.linspace(0, 1 - 1.0 / K, K, device=dev)
    z = torch.sort(ops.sample_coarse(rays, torch.rand((rays.shape[0], K), device=dev, generator=g), lin, True), dim=1).values.contiguous()
    ro = ops.render_pass(sc, mlp, rays, z, per_sample=False); d0 = ro["depth"].clone()
    for i in range(n // 2):
        ro = ops.render_pass(sc, mlp, rays, z, per_sample=False, out=ro)
    torch.cuda.synchronize(); print(f"render D={D}", n // 2, "launches, equal:", bool(torch.equal(d0, ro["depth"])), flush=True)
fm32 = ops.featmap_pack(torch.randn((1, 256, 192, 640), device=dev, generator=g), torch.float32)
sc3 = ops.Scene(feat=fm32[0], K_f=torch.from_numpy(Kc[None]).to(dev), w2c_f=torch.eye(4, device=dev)[None])
mlp3 = ops.Mlp(*syn.make_mlp(0), device=dev)
sc3 = sc3.project_x3(mlp3)
dp = torch.from_numpy(syn.ssc_voxel_grid()[::2].copy()).to(dev)
q = ops.query_points(sc3, mlp3, dp, want_rgb=False, precision=ops.F32TC); s0 = q["sigma"].clone()
oc = dict(q); oc["invalid_features"] = oc["invalid_features"].view(torch.uint8)
for i in range(n):
    ops.query_points(sc3, mlp3, dp, want_rgb=False, precision=ops.F32TC, out=oc)
torch.cuda.synchronize(); print("x3 query", n, "launches, equal:", bool(torch.equal(s0, oc["sigma"])), flush=True)
