"""Where the e2e step's time goes at N ranks: host issue vs device, with and without the device->host leg.

python profiles/host_e2e.py                       (one GPU; also prints a cProfile of the step)
torchrun --nproc-per-node N profiles/host_e2e.py  (N concurrent ranks, max over ranks)
"""
import os, sys, time, json, cProfile, pstats
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import scenedino_b200 as sd
from scenedino_b200 import ops, synthetic as syn

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = f"cuda:{lr}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
g = torch.Generator(device=dev).manual_seed(1)
holder = {"map": torch.randn((1, 256, 384, 1280), device=dev, generator=g)}
net = bench.build_net(sd, torch, holder, dev, "fp16")
Kt = torch.from_numpy(syn.kitti360_K()[None]).to(dev)[None]
eye = torch.eye(4, device=dev)[None, None]
img = torch.zeros(1, 1, 3, 8, 8, device=dev)
net.encode(img, Kt, eye, ids_encoder=[0], ids_render=[0], images_alt=img); net.set_scale(0)
N = 256 * 256 * 32
NB = 2
pts = [torch.empty((1, N, 3), device=dev) for _ in range(NB)]
res = [torch.empty(N * 5, dtype=torch.uint8, device=dev) for _ in range(NB)]
host = [torch.empty(N * 5, dtype=torch.uint8).pin_memory() for _ in range(NB)]
ev_k = [torch.cuda.Event() for _ in range(NB)]; ev_out = [torch.cuda.Event() for _ in range(NB)]
d2h = torch.cuda.Stream(); main_s = torch.cuda.current_stream()
T = syn.velo_to_cam()
state = {"i": 0}

def step(query=True, back=True):
    b = state["i"] % NB; state["i"] += 1
    if query:
        ops.gen_voxel_grid(T, dims=(256, 256, 32), out=pts[b][0])
        main_s.wait_event(ev_out[b])
        with torch.no_grad():
            _, invalid, sigma, _, _ = net(pts[b], only_density=True)
        res[b][:N * 4].view(torch.float32).copy_(sigma.reshape(-1))
        res[b][N * 4:].copy_(invalid.reshape(-1))
    ev_k[b].record()
    if back:
        with torch.cuda.stream(d2h):
            d2h.wait_event(ev_k[b])
            host[b].copy_(res[b], non_blocking=True)
            ev_out[b].record()

def fence():
    main_s.wait_stream(d2h)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def mx(v):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return t.item()

def run(name, steps=200, **kw):
    for _ in range(5): step(**kw)
    fence()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(steps): step(**kw)
    t1 = time.perf_counter()
    main_s.wait_stream(d2h); e1.record(); fence()
    out = {"variant": name, "world": world, "host_issue_ms": round(mx(1e3 * (t1 - t0) / steps), 4),
           "device_ms": round(mx(e0.elapsed_time(e1) / steps), 4)}
    if rank == 0:
        print(json.dumps(out), flush=True)

run("query + read-back")
run("query only", back=False)
run("read-back only", query=False)
run("query + read-back (again)")
if world == 1:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(50): step()
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(16)
if world > 1:
    dist.destroy_process_group()
