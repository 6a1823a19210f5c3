"""Absolute clock64 timeline (CTA 0) of consecutive tiles through all roles of field_bin_kernel, binned output.
    SD_TC_DEBUG=8192 python profiles/trace_bin4.py"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from scenedino_b200 import _abi, ops, synthetic as syn  # noqa: E402

dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(1)
feat = ops.featmap_pack(torch.randn((1, 256, 384, 1280), device=dev, generator=g), torch.float16)
sc = ops.Scene(feat=feat[0], K_f=torch.from_numpy(syn.kitti360_K()[None]).to(dev), w2c_f=torch.eye(4, device=dev)[None])
mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F16)
scp = sc.project(mlp)
dp = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
b = ops.query_points_binned(scp, mlp, dp)
ob = dict(b); ob['invalid_features'] = ob['invalid_features'].view(torch.uint8)
ops.query_points_binned(scp, mlp, dp, out=ob)
for _ in range(3):
    ops.query_points_binned(scp, mlp, dp, out=ob, reuse_sorted=True)
torch.cuda.synchronize()
raw = ctypes.CDLL(_abi.LIB_PATH)
buf = (ctypes.c_longlong * (8 * 64 * 8))()
raw.sd_debug_read_trace_bin(buf)
a = np.array(buf[:]).reshape(8, 64, 8)
t0 = a[1, 10, 0]
f = lambda v: f"{int(v - t0):6d}"
print("tile m | publish: begin gotREC end | tma: start gotB issued | pt: start code EMPTY_C sts EMPTY_A done | mma1: start A B chunks commit | epi1: start end | mma2 issue | epi2: start end")
for j in range(10, 26):
    print(f"{j:3d} {a[6, j, 7]:2d} | {f(a[6,j,3])} {f(a[6,j,4])} {f(a[6,j,5])} | {f(a[6,j,0])} {f(a[6,j,1])} {f(a[6,j,2])} | {f(a[2,j,0])} {f(a[2,j,4])} {f(a[2,j,5])} {f(a[2,j,1])} {f(a[2,j,2])} {f(a[2,j,3])} |"
          f" {f(a[1,j,0])} {f(a[1,j,1])} {f(a[1,j,2])} {f(a[1,j,4])} {f(a[1,j,6])} | {f(a[0,j,2])} {f(a[0,j,3])} | {f(a[1,j+1,5])} | {f(a[0,j,0])} {f(a[0,j,1])}")
