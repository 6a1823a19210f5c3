"""Development check of the projected-map tile kernel (field_proj.cu + field_bin.cu): parity against the gather
kernel and the oracle on the SSC grid, and timing.  Run on the GPU box: python profiles/dev_bin.py [small]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from scenedino_b200 import ops, synthetic as syn
from oracle import oracle as O
from helpers import rel_err

small = 'small' in sys.argv
C_, Hf, Wf = (256, 192, 640) if small else (256, 384, 1280)
dev = 'cuda'
feat = syn.make_feature_map(1, C_, Hf, Wf)
K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
mlp_w = syn.make_mlp(0, bias_scale=0.05)
pts = syn.ssc_voxel_grid()
sc = ops.Scene.from_arrays(feat, K, w2c, device=dev, feat_dtype=torch.float16)
mlp = ops.Mlp(*mlp_w, device=dev, precision=ops.F16)
dp = torch.from_numpy(pts).to(dev)
t0 = time.time()
scp = sc.project(mlp)
torch.cuda.synchronize()
print('project ok', time.time() - t0, flush=True)
# --- P against a torch matmul of the same fp16 operands
Pm = scp.proj[50176:].view(torch.float16).view(Hf * Wf, 128).float()
Wf16 = torch.from_numpy(mlp_w[0][:, :256]).to(dev).half().float()
ref = sc.feat.view(Hf * Wf, 256).float() @ Wf16.T
err = (Pm - ref).abs().max().item()
print('P max abs err vs fp32 matmul of fp16 operands', err, 'ref absmax', ref.abs().max().item(), flush=True)
assert err < 2e-2
q_old = ops.query_points(sc, mlp, dp, want_rgb=False)
torch.cuda.synchronize()
print('old ok', flush=True)
q_new = ops.query_points(scp, mlp, dp, want_rgb=False)
torch.cuda.synchronize()
print('new ok', flush=True)
assert torch.equal(q_new['invalid_features'], q_old['invalid_features'])
for k in ('sigma', 'dino'):
    a, b = q_new[k].cpu().numpy(), q_old[k].cpu().numpy()
    e = rel_err(a, b)
    print(k, 'new vs old: max rel', e.max(), 'mean', e.mean(), 'finite', np.isfinite(a).all(), flush=True)
sub = np.arange(0, len(pts), 257)
o = O.query_points(O.Scene(feat=feat, K_f=K, w2c_f=w2c), O.Mlp(*mlp_w), pts[sub], want_rgb=False)
for k in ('sigma', 'dino'):
    for name, q in (('new', q_new), ('old', q_old)):
        e = rel_err(q[k].cpu().numpy()[sub], o[k])
        print(k, name, 'vs oracle: max rel', e.max(), flush=True)
out_new = {k: v for k, v in q_new.items()}; out_new['invalid_features'] = out_new['invalid_features'].view(torch.uint8)
out_old = {k: v for k, v in q_old.items()}; out_old['invalid_features'] = out_old['invalid_features'].view(torch.uint8)
def tm(scene, out, n=20):
    for _ in range(3): ops.query_points(scene, mlp, dp, want_rgb=False, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): ops.query_points(scene, mlp, dp, want_rgb=False, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_old, t_new = tm(sc, out_old), tm(scp, out_new)
print(f'old {t_old:.3f} ms  new {t_new:.3f} ms  -> {len(pts)/t_new/1e6:.2f} Gvoxel/s', flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): sc.project(mlp)
e1.record(); torch.cuda.synchronize()
print('project ms', e0.elapsed_time(e1) / 5)
