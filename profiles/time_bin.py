"""Tile-kernel timings on the SSC grid (ViT-B/8 map): caller-order rows (scattered STG) against binned rows (TMA tile
stores), sort reused (kernel alone) and full query.    python profiles/time_bin.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedino_b200 import ops, synthetic as syn  # noqa: E402

dev = "cuda"
Hf, Wf = (192, 640) if "small" in sys.argv else (384, 1280)
g = torch.Generator(device=dev).manual_seed(1)
fm = ops.featmap_pack(torch.randn((1, 256, Hf, Wf), device=dev, generator=g), torch.float16)
K = syn.kitti360_K()
sc = ops.Scene(feat=fm[0], K_f=torch.from_numpy(K[None]).to(dev), w2c_f=torch.eye(4, device=dev)[None])
mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F16)
sc = sc.project(mlp)
dp = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
N = dp.shape[0]


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def kernel_ms(fn, n=10):
    """the tile kernel's own duration: CUDA events the library records around its launch (sd_profile_next_kernel)"""
    from scenedino_b200 import _abi
    ms = []
    for _ in range(n + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b.record()
        _abi.check(_abi.lib().sd_profile_next_kernel(a.cuda_event, b.cuda_event), "sd_profile_next_kernel")
        fn()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms[2:]))


q = ops.query_points(sc, mlp, dp, want_rgb=False)
oc = {k: v for k, v in q.items()}
oc["invalid_features"] = oc["invalid_features"].view(torch.uint8)
ops.query_points(sc, mlp, dp, want_rgb=False, out=oc)
b = ops.query_points_binned(sc, mlp, dp)
ob = {k: v for k, v in b.items()}
ob["invalid_features"] = ob["invalid_features"].view(torch.uint8)
ops.query_points_binned(sc, mlp, dp, out=ob)
ok = bool(torch.equal(ob["dino_binned"], oc["dino"][ob["perm"].long()]) and torch.equal(ob["sigma"], oc["sigma"]))
res = {"map": [Hf, Wf], "binned_equals_caller_order": ok,
       "caller_full_ms": timed(lambda: ops.query_points(sc, mlp, dp, want_rgb=False, out=oc)),
       "caller_kernel_ms": timed(lambda: ops.query_points_sorted(sc, mlp, dp, oc)),
       "binned_full_ms": timed(lambda: ops.query_points_binned(sc, mlp, dp, out=ob)),
       "binned_kernel_ms": timed(lambda: ops.query_points_binned(sc, mlp, dp, out=ob, reuse_sorted=True))}
res["caller_kernel_ms"] = kernel_ms(lambda: ops.query_points_sorted(sc, mlp, dp, oc))
res["binned_kernel_ms"] = kernel_ms(lambda: ops.query_points_binned(sc, mlp, dp, out=ob, reuse_sorted=True))
alg = N * (12 + 4 + 256) + 379365 * 256
res["binned_kernel_hbm_frac"] = alg / (res["binned_kernel_ms"] * 1e-3) / 6552.6e9
res["caller_kernel_hbm_frac"] = alg / (res["caller_kernel_ms"] * 1e-3) / 6552.6e9
print(json.dumps(res))
