"""Runs the binned / caller-order query with cuda:1 as the current device (needs a 2-GPU box)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedino_b200 import ops, synthetic as syn
torch.cuda.set_device(1)
dev = "cuda:1"
g = torch.Generator(device=dev).manual_seed(1)
fm = ops.featmap_pack(torch.randn((1, 256, 384, 1280), device=dev, generator=g), torch.float16)
sc = ops.Scene(feat=fm[0], K_f=torch.from_numpy(syn.kitti360_K()[None]).to(dev), w2c_f=torch.eye(4, device=dev)[None])
mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F16)
sc = sc.project(mlp)
dp = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
q = ops.query_points(sc, mlp, dp, want_rgb=False); torch.cuda.synchronize(); print("caller ok", flush=True)
b = ops.query_points_binned(sc, mlp, dp); torch.cuda.synchronize(); print("binned ok", flush=True)
print(torch.equal(b["dino_binned"], q["dino"][b["perm"].long()]))
ob = dict(b); ob["invalid_features"] = ob["invalid_features"].view(torch.uint8)
gr = ops.QueryGraph(sc, mlp, dp, ob, binned_out=True); gr.replay(); torch.cuda.synchronize(); print("graph ok", flush=True)
