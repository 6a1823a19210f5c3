"""clock64 timeline of field_tc_kernel in render mode (CTA 0): python profiles/trace_render.py [proj]"""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, '.')
from scenedino_b200 import ops, synthetic as syn, _abi
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(1)
feat = ops.featmap_pack(torch.randn((1, 256, 384, 1280), device=dev, generator=g), torch.float16)
K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
imgs = syn.make_images(2, 1)
Kf = torch.from_numpy(K).to(dev); Wf = torch.from_numpy(w2c).to(dev)
sc = ops.Scene(feat=feat[0], K_f=Kf, w2c_f=Wf, rgb=torch.from_numpy(imgs).to(dev), K_c=Kf, w2c_c=Wf)
mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F16)
if 'proj' in sys.argv:
    sc = sc.project(mlp)
rays = torch.from_numpy(syn.image_rays(syn.view_pose_c2w(1), K[0])).to(dev)
R, Kc = rays.shape[0], 64
lin = torch.linspace(0, 1 - 1.0 / Kc, Kc, device=dev)
u = torch.rand((R, Kc), device=dev, generator=g)
z = ops.sample_coarse(rays, u, lin, True)
def run():
    return ops.render_pass(sc, mlp, rays, z, per_sample=False)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f'render {ms:.3f} ms -> {R*Kc/ms/1e6:.2f} Gsamples/s')
os.environ['SD_TC_DEBUG'] = '8192'
for _ in range(2): run()
torch.cuda.synchronize()
raw = ctypes.CDLL(_abi.LIB_PATH)
buf = (ctypes.c_longlong * (4 * 64 * 8))()
raw.sd_debug_read_trace(buf)
a = np.array(buf[:]).reshape(4, 64, 8)
t0 = a[a > 100000].min()
names = ['epi', 'mma', 'pt ', 'ga ']
for j in range(30, 34):
    for r in range(4):
        print(f"tile {j:2d} {names[r]}", ' '.join(f"{(x - t0) if x > 100000 else x:8d}" for x in a[r, j]))
