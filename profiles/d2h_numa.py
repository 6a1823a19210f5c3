"""Concurrent pinned device->host bandwidth per rank, with and without binding the rank to its GPU's NUMA node.

torchrun --nproc-per-node N profiles/d2h_numa.py   (the e2e scaling question: where do N concurrent 10.5 MB reads saturate?)
"""
import os, sys, time, json, subprocess
import torch, torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)

def bus_id():
    p = torch.cuda.get_device_properties(lr)
    return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"

def numa_of(bid):
    try:
        return int(open(f"/sys/bus/pci/devices/{bid}/numa_node").read())
    except Exception as e:
        return None

def cpus_of(node):
    try:
        s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    except Exception:
        return None
    out = set()
    for part in s.split(","):
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def bw(nbytes, iters, host):
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s):
            host.copy_(src, non_blocking=True)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record()
        for _ in range(iters):
            host.copy_(src, non_blocking=True)
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return nbytes * iters * world / (t.item() * 1e-3) / 1e9   # aggregate GB/s by the slowest rank

bid = bus_id(); node = numa_of(bid)
info = {"rank": rank, "bus": bid, "numa": node, "aff0": len(os.sched_getaffinity(0)),
        "nodes": sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))}
N1 = 256 * 256 * 32 * 5
res = {}
h = torch.empty(N1, dtype=torch.uint8).pin_memory()
res["default_10MB"] = bw(N1, 200, h)
hb = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
res["default_64MB"] = bw(64 << 20, 40, hb)
del h, hb
cp = cpus_of(node) if node is not None and node >= 0 else None
if cp:
    try:
        os.sched_setaffinity(0, cp & os.sched_getaffinity(0) or cp)
        info["bound"] = len(os.sched_getaffinity(0))
    except Exception as e:
        info["bound"] = repr(e)
    h = torch.empty(N1, dtype=torch.uint8).pin_memory()
    h.zero_()
    res["bound_10MB"] = bw(N1, 200, h)
    hb = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); hb.zero_()
    res["bound_64MB"] = bw(64 << 20, 40, hb)
# cudaHostAlloc through torch's caching host allocator vs a write-combined/portable mapping is not reachable from torch; the
# remaining knob is splitting one read over two streams
h = torch.empty(N1, dtype=torch.uint8).pin_memory()
src = torch.empty(N1, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(); half = N1 // 2
barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
e0.record(); s1.wait_event(e0); s2.wait_event(e0)
for _ in range(200):
    with torch.cuda.stream(s1): h[:half].copy_(src[:half], non_blocking=True)
    with torch.cuda.stream(s2): h[half:].copy_(src[half:], non_blocking=True)
e1.record(s1); e2.record(s2)
barrier()
ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
t = torch.tensor([ms], device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
res["two_streams_10MB"] = N1 * 200 * world / (t.item() * 1e-3) / 1e9
for r in range(world):
    if r == rank:
        print(json.dumps({**info, **({k: round(v, 1) for k, v in res.items()} if rank == 0 else {})}), flush=True)
    if world > 1: dist.barrier()
if rank == 0:
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[:3000])
        print(subprocess.run(["lscpu"], capture_output=True, text=True, timeout=20).stdout[:1500])
    except Exception as e:
        print("topo unavailable", e)
if world > 1:
    dist.destroy_process_group()
