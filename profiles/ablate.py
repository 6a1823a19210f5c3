import os, sys, torch, numpy as np
sys.path.insert(0, '.')
from scenedino_b200 import ops, synthetic as syn
dev='cuda'
g=torch.Generator(device=dev).manual_seed(1)
feat=ops.featmap_pack(torch.randn((1,256,384,1280),device=dev,generator=g), torch.float16)
K=syn.kitti360_K()[None]; w2c=np.eye(4,dtype=np.float32)[None]
scene=ops.Scene(feat=feat[0],K_f=torch.from_numpy(K).to(dev),w2c_f=torch.from_numpy(w2c).to(dev))
mlp=ops.Mlp(*syn.make_mlp(0),device=dev,precision=ops.F16)
pts=torch.from_numpy(syn.ssc_voxel_grid()).to(dev)
N=len(pts)
out=dict(sigma=torch.empty(N,device=dev),dino=torch.empty((N,64),device=dev),invalid_features=torch.empty(N,dtype=torch.uint8,device=dev))
def t(binned, flag):
    """GPU time per query: 10 calls captured in a CUDA graph (no Python/ctypes launch gaps)."""
    os.environ['SD_TC_DEBUG']=str(flag)
    for _ in range(2): ops.query_points(scene,mlp,pts,want_rgb=False,out=out,binned=binned)
    torch.cuda.synchronize()
    st=torch.cuda.Stream()
    with torch.cuda.stream(st):
        gr=torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(10): ops.query_points(scene,mlp,pts,want_rgb=False,out=out,binned=binned)
        gr.replay(); torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(st); gr.replay(); e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/10
import ctypes
from scenedino_b200 import _abi
raw=ctypes.CDLL(_abi.LIB_PATH)
def trace(flag, label, binned=False):
    os.environ['SD_TC_DEBUG']=str(flag+8192)
    for _ in range(2): ops.query_points(scene,mlp,pts,want_rgb=False,out=out,binned=binned)
    torch.cuda.synchronize()
    buf=(ctypes.c_longlong*(4*64*8))()
    raw.sd_debug_read_trace(buf)
    a=np.array(buf[:]).reshape(4,64,8)
    t0=a[a>0].min()
    a=np.where(a>0,a-t0,-1)
    print('=== trace', label)
    names=['epi','mma','pt ','ga ']
    for j in range(20,28):
        for r in range(4):
            print(f"tile {j:2d} {names[r]}", ' '.join(f"{v:7d}" for v in a[r,j]))
trace(0,'full unbinned')
trace(0,'full binned', True)
print('full unbinned', t(False,0)*1000, 'us;  binned', t(True,0)*1000,'us')
