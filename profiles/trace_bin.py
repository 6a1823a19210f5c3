"""clock64 timeline of the four roles of field_bin_kernel (CTA 0), SSC grid.  python profiles/trace_bin.py"""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, '.')
from scenedino_b200 import ops, synthetic as syn, _abi
C_, Hf, Wf = 256, 384, 1280
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(1)
feat = ops.featmap_pack(torch.randn((1, C_, Hf, Wf), device=dev, generator=g), torch.float16)
K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
sc = ops.Scene(feat=feat[0], K_f=torch.from_numpy(K).to(dev), w2c_f=torch.from_numpy(w2c).to(dev))
mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F16)
scp = sc.project(mlp)
dp = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
out = None
q = ops.query_points(scp, mlp, dp, want_rgb=False)
out = dict(q); out['invalid_features'] = out['invalid_features'].view(torch.uint8)
os.environ['SD_TC_DEBUG'] = '8192'
for _ in range(2): ops.query_points(scp, mlp, dp, want_rgb=False, out=out)
torch.cuda.synchronize()
raw = ctypes.CDLL(_abi.LIB_PATH)
buf = (ctypes.c_longlong * (8 * 64 * 8))()
raw.sd_debug_read_trace_bin(buf)
a = np.array(buf[:]).reshape(8, 64, 8)
t0 = a[a > 100000].min()
names = ['epi', 'mma', 'pt0', 'pt1', 'pt2', 'pt3', 'tma']
for j in list(range(0, 2)) + list(range(34, 40)):
    for r in range(7):
        v = a[r, j]
        print(f"tile {j:2d} {names[r]}", ' '.join(f"{(x - t0) if x > 100000 else x:8d}" for x in v))
cb = (ctypes.c_ulonglong * 512)()
raw.sd_debug_read_cta_ns(cb)
c = np.array(cb[:]).reshape(256, 2)[:148].astype(np.int64)
t0c = c[:, 0].min()
dur = (c[:, 1] - t0c) / 1000.0
print('CTA end times (us after first start): min %.1f median %.1f max %.1f; starts spread %.1f us' % (dur.min(), np.median(dur), dur.max(), (c[:,0].max()-t0c)/1000.0))
np.save('gpurun_out/cta_dur.npy', dur)
print('slowest CTAs:', np.argsort(-dur)[:8], np.sort(-dur)[:8] * -1)
os.environ['SD_TC_DEBUG'] = '0'
def tm(flag, n=20):
    os.environ['SD_TC_DEBUG'] = str(flag)
    for _ in range(3): ops.query_points(scp, mlp, dp, want_rgb=False, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): ops.query_points(scp, mlp, dp, want_rgb=False, out=out)
    e1.record(); torch.cuda.synchronize()
    os.environ['SD_TC_DEBUG'] = '0'
    return e0.elapsed_time(e1) / n * 1000
print('step us: full', tm(0), ' no dino stores', tm(1), ' no sigma/mask stores', tm(2), ' neither', tm(3))
# per-tile period in steady state, from the MMA role's tile start stamps
st = a[1, 8:60, 0]
print('mean cycles per tile (mma tile starts):', np.diff(st).mean())
print('chunks per tile (tma):', a[6, 8:60, 7].mean())
