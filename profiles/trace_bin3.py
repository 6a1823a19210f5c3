"""Whole-run tile timeline of CTAs 0..3 of field_bin_kernel (binned output): period against chunk count, ramp and tail.
    SD_TC_DEBUG=8192 python profiles/trace_bin3.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from scenedino_b200 import _abi, ops, synthetic as syn  # noqa: E402

dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(1)
feat = ops.featmap_pack(torch.randn((1, 256, 384, 1280), device=dev, generator=g), torch.float16)
K = syn.kitti360_K()[None]
sc = ops.Scene(feat=feat[0], K_f=torch.from_numpy(K).to(dev), w2c_f=torch.eye(4, device=dev)[None])
mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F16)
scp = sc.project(mlp)
dp = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
b = ops.query_points_binned(scp, mlp, dp)
ob = dict(b); ob['invalid_features'] = ob['invalid_features'].view(torch.uint8)
for _ in range(3):
    ops.query_points_binned(scp, mlp, dp, out=ob)
torch.cuda.synchronize()
raw = ctypes.CDLL(_abi.LIB_PATH)
buf = (ctypes.c_longlong * (4 * 2 * 512))()
raw.sd_debug_read_tiles_bin(buf)
a = np.array(buf[:]).reshape(4, 512, 2)
cb = (ctypes.c_ulonglong * 512)()
raw.sd_debug_read_cta_ns(cb)
c = np.array(cb[:]).reshape(256, 2)[:148].astype(np.int64)
t0c = c[:, 0].min()
print('CTA start spread %.1f us; end min/median/max %.1f %.1f %.1f us after first start' % (
    (c[:, 0].max() - t0c) / 1e3, (c[:, 1].min() - t0c) / 1e3, np.median(c[:, 1] - t0c) / 1e3, (c[:, 1].max() - t0c) / 1e3))
for cta in range(4):
    t, m = a[cta, :500, 0], a[cta, :500, 1]
    n = int((t > 0).sum())
    print(f'CTA {cta}: entry -> first tile start {a[cta,0,0]-a[cta,510,0]} cycles; last tile start -> exit {a[cta,511,0]-t[n-1]} cycles; entry -> exit {a[cta,511,0]-a[cta,510,0]}')
    t, m = t[:n], m[:n]
    d = np.diff(t)
    print(f"CTA {cta}: {n} tiles, total {t[-1] - t[0]} cycles, chunks total {m.sum()}")
    for lo in range(0, n - 1, 16):
        hi = min(n - 1, lo + 16)
        print(f"   tiles {lo:3d}-{hi:3d}: mean period {d[lo:hi].mean():7.0f}  chunks/tile {m[lo:hi].mean():5.2f}  max chunks {m[lo:hi].max()}")
    one = d[(m[:-1] == 1)]
    two = d[(m[:-1] == 2)]
    print(f"   single-chunk tiles: period mean {one.mean():.0f} median {np.median(one):.0f};  two-chunk: mean {two.mean() if len(two) else 0:.0f}")
