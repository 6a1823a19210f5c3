"""Debug build helper (-DSD_DEBUG_WAIT): which barrier waits time out in field_bin_kernel."""
import sys, ctypes
import numpy as np, torch
sys.path.insert(0, '.')
from scenedino_b200 import ops, synthetic as syn, _abi
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(1)
feat = ops.featmap_pack(torch.randn((1, 256, 384, 1280), device=dev, generator=g), torch.float16)
K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
sc = ops.Scene(feat=feat[0], K_f=torch.from_numpy(K).to(dev), w2c_f=torch.from_numpy(w2c).to(dev))
mlp = ops.Mlp(*syn.make_mlp(0), device=dev, precision=ops.F16)
scp = sc.project(mlp)
dp = torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
q = ops.query_points(scp, mlp, dp, want_rgb=False)
torch.cuda.synchronize()
raw = ctypes.CDLL(_abi.LIB_PATH)
buf = (ctypes.c_uint * 260)()
raw.sd_debug_read_timeout(buf)
n = buf[0]
print('timeouts', n)
base = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for k in range(min(n, 40)):
    bar, par, tid, blk = buf[4 + 4 * k], buf[5 + 4 * k], buf[6 + 4 * k], buf[7 + 4 * k]
    print('bar addr', bar, 'idx', (bar - 1024 - base) // 8 if base else '?', 'parity', par, 'warp', tid // 32, 'lane', tid % 32, 'block', blk)
