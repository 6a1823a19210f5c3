"""Binned tile kernel alone (sort reused), SSC grid, ViT-B/8 map: kernel us from the library's own events."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedino_b200 import _abi, ops, synthetic as syn  # noqa: E402
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
fm = ops.featmap_pack(torch.randn((1, 256, 384, 1280), device=dev, generator=g), torch.float16)
sc = ops.Scene(feat=fm[0], K_f=torch.from_numpy(syn.kitti360_K()[None]).to(dev), w2c_f=torch.eye(4, device=dev)[None])
mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F16)
sc = sc.project(mlp)
dp = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
b = ops.query_points_binned(sc, mlp, dp)
ob = dict(b); ob["invalid_features"] = ob["invalid_features"].view(torch.uint8)
ops.query_points_binned(sc, mlp, dp, out=ob)
ms = []
for _ in range(14):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); e.record()
    _abi.check(_abi.lib().sd_profile_next_kernel(a.cuda_event, e.cuda_event), "p")
    ops.query_points_binned(sc, mlp, dp, out=ob, reuse_sorted=True)
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(e))
print(f"{os.path.basename(os.environ.get('SD_B200_LIB', 'default'))}: binned kernel {np.median(ms[4:]) * 1000:.1f} us  (min {min(ms[4:]) * 1000:.1f})")
