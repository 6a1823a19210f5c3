"""Minimal driver for ncu: projected-map SSC voxel query, 4 launches.  python profiles/run_bin.py"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from scenedino_b200 import ops, synthetic as syn
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(1)
feat = ops.featmap_pack(torch.randn((1, 256, 384, 1280), device=dev, generator=g), torch.float16)
K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
sc = ops.Scene(feat=feat[0], K_f=torch.from_numpy(K).to(dev), w2c_f=torch.from_numpy(w2c).to(dev))
mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F16)
scp = sc.project(mlp)
dp = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
q = ops.query_points(scp, mlp, dp, want_rgb=False)
out = dict(q); out['invalid_features'] = out['invalid_features'].view(torch.uint8)
for _ in range(3): ops.query_points(scp, mlp, dp, want_rgb=False, out=out)
torch.cuda.synchronize()
print('ok', float(q['sigma'].sum()))
