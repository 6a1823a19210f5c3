"""Host issue time of the pieces of one SSC frame through BTSNet (no device sync inside the timed pieces)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import scenedino_b200 as sd
from scenedino_b200 import synthetic as syn
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
holder = {"map": torch.randn((1, 256, 384, 1280), device=dev, generator=g)}
net = bench.build_net(sd, torch, holder, dev, "fp16")
Kt = torch.from_numpy(syn.kitti360_K()[None]).to(dev)[None]
eye = torch.eye(4, device=dev)[None, None]
img = torch.zeros(1, 1, 3, 8, 8, device=dev)
xyz = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)[None]
net.static_query, net.materialize_dino_full, net.one_hot_seg = True, False, False
def frame():
    net.encode(img, Kt, eye, ids_encoder=[0], ids_render=[0], images_alt=img); net.set_scale(0)
    with torch.no_grad():
        return net(xyz, predict_segmentation=True)
for _ in range(3): frame()
torch.cuda.synchronize()
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(20): frame()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
