"""SD_MLP_F32_TC on the SSC grid (ViT-B/8 map): projection, full query, tile kernel alone; max error against the fp32 kernel."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedino_b200 import _abi, ops, synthetic as syn
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
fm = ops.featmap_pack(torch.randn((1, 256, 384, 1280), device=dev, generator=g), torch.float32)
sc = ops.Scene(feat=fm[0], K_f=torch.from_numpy(syn.kitti360_K()[None]).to(dev), w2c_f=torch.eye(4, device=dev)[None])
mlp = ops.Mlp(*syn.make_mlp(0), device=dev)
dp = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
N = dp.shape[0]

def timed(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    return float(np.median(ms))

st = {}
def proj(): st["s"] = sc.project_x3(mlp)
res = {"project_x3_ms": timed(proj, 4)}
scp = st["s"]
q = ops.query_points(scp, mlp, dp, want_rgb=False, precision=ops.F32TC)
oc = dict(q); oc["invalid_features"] = oc["invalid_features"].view(torch.uint8)
ops.query_points(scp, mlp, dp, want_rgb=False, precision=ops.F32TC, out=oc)
res["full_ms"] = timed(lambda: ops.query_points(scp, mlp, dp, want_rgb=False, precision=ops.F32TC, out=oc))
km = []
for _ in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); b.record()
    _abi.check(_abi.lib().sd_profile_next_kernel(a.cuda_event, b.cuda_event), "p")
    ops.query_points(scp, mlp, dp, want_rgb=False, precision=ops.F32TC, out=oc); torch.cuda.synchronize()
    km.append(a.elapsed_time(b))
res["kernel_ms"] = float(np.median(km[2:]))
res["gvoxel_s_full"] = N / res["full_ms"] / 1e6
sub = slice(0, 262144)
q32 = ops.query_points(sc, mlp, dp[sub].contiguous(), want_rgb=False, precision=ops.FP32)
for k in ("sigma", "dino"):
    a, b = q[k][sub].double(), q32[k].double()
    res[f"max_rel_{k}_vs_fp32_kernel"] = float(((a - b).abs() / torch.clamp(b.abs(), min=b.pow(2).mean().sqrt())).max())
print(json.dumps(res))
