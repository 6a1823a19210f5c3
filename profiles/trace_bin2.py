"""Per-role clock64 averages of field_bin_kernel (CTA 0, tiles TB_T0..TB_T0+63), caller-order against binned output.
    python profiles/trace_bin2.py"""
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
q = ops.query_points(scp, mlp, dp, want_rgb=False)
oc = dict(q); oc['invalid_features'] = oc['invalid_features'].view(torch.uint8)
ops.query_points(scp, mlp, dp, want_rgb=False, out=oc)
b = ops.query_points_binned(scp, mlp, dp)
ob = dict(b); ob['invalid_features'] = ob['invalid_features'].view(torch.uint8)
ops.query_points_binned(scp, mlp, dp, out=ob)
raw = ctypes.CDLL(_abi.LIB_PATH)
calls = {'caller': lambda: ops.query_points_sorted(scp, mlp, dp, oc),
         'binned': lambda: ops.query_points_binned(scp, mlp, dp, out=ob, reuse_sorted=True)}


def kernel_us(fn, n=10):
    ms = []
    for _ in range(n + 2):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); e.record()
        _abi.check(_abi.lib().sd_profile_next_kernel(a.cuda_event, e.cuda_event), "p")
        fn(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(e))
    return float(np.median(ms[2:])) * 1000


flag = int(os.environ.get('SD_TC_DEBUG', '0'))       # read once per process by the library: one process per flag
for name, fn in calls.items():
    if not (flag & 8192):
        print(f"{name}: SD_TC_DEBUG={flag} (1: no dino stores, 2: no sigma stores) kernel {kernel_us(fn):.1f} us")
        continue
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (8 * 64 * 8))()
    raw.sd_debug_read_trace_bin(buf)
    a = np.array(buf[:]).reshape(8, 64, 8).astype(np.float64)
    s = slice(4, 60)
    d = lambda r, e1, e0: (a[r, s, e1] - a[r, s, e0]).mean()
    print(f"== {name}: period {np.diff(a[1, s, 0]).mean():.0f} cycles/tile, chunks/tile {a[6, s, 7].mean():.2f}")
    print(f"   epi2 busy {d(0,1,0):.0f}   epi1 busy {d(0,3,2):.0f}   epi1 start after mma commit {(a[0,s,2]-a[1,s,6]).mean():.0f}")
    print(f"   mma: wait A {d(1,1,0):.0f}  wait B {d(1,2,1):.0f}  chunks {d(1,4,2):.0f}  code+commit {d(1,6,4):.0f}  total {d(1,6,0):.0f}")
    print(f"   pt0: code {d(2,4,0):.0f}  wait EMPTY_C {d(2,5,4):.0f}  code sts {d(2,1,5):.0f}  wait EMPTY_A {d(2,2,1):.0f}  weights {d(2,3,2):.0f}  total {d(2,6,0):.0f}")
    print(f"   tma: wait EMPTY_B {d(6,1,0):.0f}  issue {d(6,2,1):.0f}")
    print(f"   epi2 start after layer-2 issue of the tile {(a[0,s,0][1:]-a[1,s,5][1:]).mean():.0f} (mma ev5 is stamped at index j+1)")
    print(f"   epi2 start after epi1 end {(a[0,s,0]-a[0,s,3]).mean():.0f}")
