"""One SSC frame through BTSNet (encode of a resident map + forward(predict_segmentation=True) on the static grid): host issue
time and device time per frame, with the pose inverse of encode() replayed from a CUDA graph and eager."""
import os, sys, time, json
import torch
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

for mode in (True, False, True):
    net.graph_pose_inverse = mode
    for _ in range(5): frame()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(50): frame()
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"graph_pose_inverse": mode, "host_issue_ms": round(1e3 * (t1 - t0) / 50, 4),
                      "device_ms": round(e0.elapsed_time(e1) / 50, 4),
                      "captured": [type(v).__name__ for v in net._pose_inverse._by_shape.values()]}))
