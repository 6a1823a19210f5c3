#!/usr/bin/env python
"""Benchmark of the feature-field query-and-render hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp16|fp32] [--impl reference]

Headline workload (BASELINE.json configs[1]): the SSCBench voxel-grid query -- 256 x 256 x 32 voxels
@ 0.2 m projected into one 192 x 640 view whose DINO ViT-B/8 feature map is 256 x 384 x 1280, MLP head
295 -> 128 -> 65.  One "step" = one pass of the hot path over the whole grid (2 097 152 voxels):
sd_query_points -> sigma [N] + 64-d features [N,64] + frustum mask [N].  Synthetic data: seeded random
feature map and random-init (kaiming) head weights.  The same line also carries a full-image render
(122 880 rays x 64 samples) as ``render``.

N > 1 (torchrun, one rank per GPU): every rank queries one full grid against its own replica of the
map (weak scaling: voxel slabs / frames are independent), then all ranks all-gather the density grid
and the frustum mask over NCCL; time = max over ranks.

``--impl reference``: the CPU restatement of the reference algorithm (oracle/, the reference is pure
Python and cannot travel to the GPU box) on all host cores, rank 0 only, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from scenedino_b200 import synthetic as syn  # noqa: E402

GRID = (256, 256, 32)
C_FEAT, HF, WF = 256, 384, 1280          # DINO ViT-B/8 + DPT head at 192 x 640
D_IN, D_HID, D_OUT = 295, 128, 65
FLOP_PER_POINT = 2 * (D_IN * D_HID + D_HID * D_OUT)   # 92 160
RENDER_R, RENDER_K = syn.IMG_H * syn.IMG_W, 64


def measured_traffic(kernel):
    """dram bytes per launch of `kernel` from the latest committed ncu summary (profiles/*_traffic.json), or None."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        try:
            d = json.load(open(f))
        except (OSError, ValueError):
            continue
        if kernel in d:
            best = d[kernel]
    return best


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sus=1400.0, src="fallback (B200_PROFILING.md)")


# ---- nvidia-smi clock sampler ---------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- workload -------------------------------------------------------------------------------------
def unique_texels(pts, K, Hf, Wf):
    """Number of distinct feature-map texels the 4-tap footprints of all points touch (for the
    compulsory-bytes figure of SURVEY.md 8d); fp32 numpy restatement of the projection."""
    p = pts.astype(np.float32)
    q = (K.astype(np.float32) @ p.T).T
    zc = np.maximum(q[:, 2], np.float32(1e-3))
    x = np.clip(q[:, 0] / zc, -2, 2); y = np.clip(q[:, 1] / zc, -2, 2)
    ix = np.clip((x + 1) * np.float32(Wf * 0.5) - 0.5, 0, Wf - 1); iy = np.clip((y + 1) * np.float32(Hf * 0.5) - 0.5, 0, Hf - 1)
    x0 = np.floor(ix).astype(np.int64); y0 = np.floor(iy).astype(np.int64)
    x1 = np.minimum(x0 + 1, Wf - 1); y1 = np.minimum(y0 + 1, Hf - 1)
    ids = np.concatenate([y0 * Wf + x0, y0 * Wf + x1, y1 * Wf + x0, y1 * Wf + x1])
    return int(np.unique(ids).size)


def cpu_reference_voxels(feat_nchw, mlp_w, pts, budget_s=12.0):
    """Times the CPU oracle (OpenMP, all host cores) on a strided sample of the grid."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    O.set_num_threads(cores)
    K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
    sc = O.Scene(feat=feat_nchw, K_f=K, w2c_f=w2c)
    mlp = O.Mlp(*mlp_w)
    n = 16384
    sub = pts[:: max(1, len(pts) // n)][:n]
    O.query_points(sc, mlp, sub[:1024], want_rgb=False)
    t0 = time.perf_counter(); O.query_points(sc, mlp, sub, want_rgb=False); dt = time.perf_counter() - t0
    n2 = int(min(len(pts), max(n, n * budget_s / max(dt, 1e-6))))
    n2 = max(n, (n2 // 4096) * 4096)
    sub = pts[:: max(1, len(pts) // n2)][:n2]
    t0 = time.perf_counter(); O.query_points(sc, mlp, sub, want_rgb=False); dt = time.perf_counter() - t0
    return dict(value=len(sub) / dt, unit="voxels/s", cores=O.num_threads(), kind="port",
                sample=f"{len(sub)} voxels strided over the 256x256x32 grid, oracle/sd_oracle.c (OpenMP), {dt:.2f} s")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    feat = syn.make_feature_map(1, C_FEAT, HF, WF)
    mlp_w = syn.make_mlp(0)
    pts = syn.ssc_voxel_grid(GRID)
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_voxels(feat, mlp_w, pts, budget_s=max(2.0, 40.0 / (args.warmup + args.steps)))
        if i >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals))
    last["value"] = v
    line = {"impl": "reference", "metric": "ssc_voxel_query_throughput", "value": v, "unit": "voxels/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": GRID[0] * GRID[1] * GRID[2] / v * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config("fp16"), "cpu_baseline": last,
            "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(precision):
    return {"workload": "SSC voxel-grid query 256x256x32 @0.2m (51.2 m), DINO ViT-B/8 map 256x384x1280, "
                        "MLP 295->128->65, outputs sigma+64-d features+mask",
            "path": ("texel sort + projected-map tile kernel (tcgen05 interpolation from TMA tiles)" if precision == "fp16"
                     else "fp32 CUDA-core parity path"),
            "voxels_per_step": GRID[0] * GRID[1] * GRID[2], "feature_map": [C_FEAT, HF, WF],
            "mlp": [D_IN, D_HID, D_OUT],
            "l2": "working set per step (map + 25 MB points + 545 MB outputs) exceeds the 126 MB L2; no explicit flush",
            "launch": "one CUDA-graph replay per step (memset + 4 sort kernels + field kernel)" if precision == "fp16" else "direct calls"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time direct library calls instead of CUDA-graph replays")
    args = ap.parse_args()
    # a run that takes absurdly long dumps its Python stacks and exits instead of hanging its caller (seconds;
    # SD_BENCH_WATCHDOG=0 switches it off)
    wd = float(os.environ.get("SD_BENCH_WATCHDOG", "900"))
    if wd > 0:
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from scenedino_b200 import _abi, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; rank 0's stdout must carry
        # exactly one JSON line, so route fd 1 to stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    pk = peaks()
    prec = ops.F16 if args.precision == "fp16" else ops.FP32
    fdt = torch.float16 if args.precision == "fp16" else torch.float32

    # ---- scene: seeded random map (encoder stand-in), camera, head -----------------------------------
    g = torch.Generator(device=dev).manual_seed(1)
    feat_nchw = torch.randn((1, C_FEAT, HF, WF), device=dev, generator=g)
    K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
    imgs = syn.make_images(2, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    feat = ops.featmap_pack(feat_nchw, fdt)            # first call also pays the module load; time the second
    torch.cuda.synchronize()
    e0.record()
    feat = ops.featmap_pack(feat_nchw, fdt)
    e1.record(); torch.cuda.synchronize()
    pack_ms = e0.elapsed_time(e1)
    scene = ops.Scene(feat=feat[0], K_f=torch.from_numpy(K).to(dev), w2c_f=torch.from_numpy(w2c).to(dev))
    scene_rgb = ops.Scene(feat=feat[0], K_f=scene.K_f, w2c_f=scene.w2c_f, rgb=torch.from_numpy(imgs).to(dev),
                          K_c=scene.K_f, w2c_c=scene.w2c_f)
    mlp_w = syn.make_mlp(0)
    mlp = ops.Mlp(*mlp_w, device=dev, precision=prec)
    project_ms = None
    if args.precision == "fp16":
        # once per encode (like the pack above): the map pushed through the feature columns of the head's first layer;
        # the voxel query then interpolates 128 hidden pre-activations on the tensor cores (field_proj.cu, field_bin.cu)
        for _ in range(2):                 # (the second call lets the caching allocator settle: no cudaMalloc in the timed one)
            scene = scene.project(mlp)
        torch.cuda.synchronize()
        e0.record()
        scene = scene.project(mlp)
        e1.record(); torch.cuda.synchronize()
        project_ms = e0.elapsed_time(e1)
        import dataclasses
        scene_rgb = dataclasses.replace(scene_rgb, proj=scene.proj)     # the render runs on the projected map too
    pts_np = syn.ssc_voxel_grid(GRID)
    N = len(pts_np)
    pts_host = torch.from_numpy(pts_np).pin_memory()
    pts = pts_host.to(dev)
    # Output buffers, double-buffered so that the all-gather of step i (communication stream) overlaps the kernel
    # of step i+1.  The density grid (fp32) and the frustum mask (u8) of a step live in ONE byte buffer, so a
    # step costs one collective of 5 B/voxel; the 64-d features stay sharded (they feed a per-voxel head).
    NB = 2
    small = [torch.empty(N * 5, dtype=torch.uint8, device=dev) for _ in range(NB)]
    outs = [dict(sigma=small[i][:N * 4].view(torch.float32), invalid_features=small[i][N * 4:],
                 dino=torch.empty((N, D_OUT - 1), device=dev)) for i in range(NB)]
    out = outs[0]
    gathered = [torch.empty((world, N * 5), dtype=torch.uint8, device=dev) for _ in range(NB)] if world > 1 else None
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    # All-gather of the small outputs.  Preferred: every rank WRITES its shard into its peers' buffers over NVLink with
    # the copy engines (symmetric memory: peer pointers + a signal-pad barrier) -- no SM is taken from the persistent
    # field kernel, which an NCCL all-gather kernel does (measured at N = 4: 87 % of ideal with NCCL).  If the symmetric
    # rendezvous is not available on the box, NCCL's all-gather is used; the JSON line says which.
    peer, gather_how = None, "nccl all_gather_into_tensor"
    if world > 1 and os.environ.get("SD_BENCH_ALLGATHER", "p2p") == "p2p":
        try:
            from scenedino_b200.sharding import PeerGather
            peer = PeerGather(N * 5, dev, n_buffers=NB)
            gathered = peer.gathered
            gather_how = "peer writes over NVLink by the copy engines (symmetric memory) + signal barrier"
        except Exception as exc:      # noqa: BLE001 -- any failure of the optional transport falls back to NCCL
            print(f"[bench rank {rank}] symmetric memory unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
            peer = None
    done_k = [torch.cuda.Event() for _ in range(NB)]      # kernel of the step using buffer b finished
    done_c = [torch.cuda.Event() for _ in range(NB)]      # all-gather reading buffer b finished
    state = {"i": 0}

    # CUDA events recorded by the library around the dominant kernel (sd_profile_next_kernel): a few direct calls before
    # the timed region give the kernel's own duration for the roofline
    k_events = []
    # the timed steps replay a CUDA graph of the query (one launch instead of seven per step), one graph per output buffer
    graphs = [ops.QueryGraph(scene, mlp, pts, outs[i]) for i in range(NB)] if args.precision == "fp16" and not args.no_graph else None

    def step(timed=False):
        b = state["i"] % NB
        state["i"] += 1
        if world > 1:
            torch.cuda.current_stream().wait_event(done_c[b])      # buffer b is free again
        if timed:
            ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ka.record(); kb.record()                                # materialise the handles
            _abi.check(_abi.lib().sd_profile_next_kernel(ka.cuda_event, kb.cuda_event), "sd_profile_next_kernel")
            k_events.append((ka, kb))
        if graphs is not None and not timed:
            graphs[b].replay()
        else:
            ops.query_points(scene, mlp, pts, want_rgb=False, out=outs[b])
        if world > 1:
            done_k[b].record()
            with torch.cuda.stream(comm):
                comm.wait_event(done_k[b])
                if peer is not None:
                    peer.gather(b, small[b])
                else:
                    dist.all_gather_into_tensor(gathered[b], small[b])
                done_c[b].record()

    def fence():
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)          # the last all-gather is inside the timed region
            dist.barrier()
        torch.cuda.synchronize()

    def note(msg):
        if os.environ.get("SD_BENCH_VERBOSE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    note("setup done")
    for _ in range(args.warmup):
        step()
    fence()
    for _ in range(5):                 # direct calls with the kernel-timing hook (not part of the timed region)
        step(timed=True)
    fence()
    kernel_ms = float(np.mean([a.elapsed_time(b_) for a, b_ in k_events]))
    note("warm-up done")
    n0 = _abi.launch_count()
    replays0 = state["i"]
    with Clocks(local) as clk:
        fence()
        e0.record()
        for _ in range(args.steps):
            step()
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)
        e1.record()
        fence()
        ms = e0.elapsed_time(e1)
        launches = _abi.launch_count() - n0
        if graphs is not None:         # kernels inside the replayed graphs (the library counts launches at capture time only)
            launches += (state["i"] - replays0) * graphs[0].launches
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        # keep the sampler alive long enough to see the load on very short runs.  The number of extra steps comes from
        # the agreed (max-over-ranks) time, so that every rank issues the same number of all-gathers: a time-based loop
        # here made ranks disagree and hang now and then.
        n_extra = int(max(0.0, 0.6 - ms / 1e3) / max(ms / args.steps / 1e3, 1e-6))
        for _ in range(n_extra):
            step()
        fence()
    note("timed region done")
    ms_step = ms / args.steps
    value = world * N / (ms_step * 1e-3)

    # ---- the same query with the texel sort reused (fixed grid and cameras, new feature map every frame: what the SSC
    #      evaluation loop does).  Reported beside `value`, which always includes the sort. -----------------------------
    sorted_reuse = None
    if args.precision == "fp16":
        ops.query_points(scene, mlp, pts, want_rgb=False, out=outs[0])
        for _ in range(3):
            ops.query_points_sorted(scene, mlp, pts, outs[0])
        fence()
        e0.record()
        for _ in range(args.steps):
            ops.query_points_sorted(scene, mlp, pts, outs[0])
        e1.record(); fence()
        sr_ms = e0.elapsed_time(e1) / args.steps
        sorted_reuse = {"value": world * N / (sr_ms * 1e-3), "unit": "voxels/s", "ms_per_step": sr_ms,
                        "note": "sd_query_points_sorted: tile kernel only, the sort of the (unchanged) points is reused"}

    # ---- end to end: pinned host points -> device -> query -> density grid + mask back to the host ------
    # Every step moves ITS OWN inputs host->device and its results device->host; the three legs run on three
    # streams with double buffers, so the copy of step i+1 overlaps the kernel of step i (PCIe is full duplex).
    h2d, d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    pts_in = [torch.empty_like(pts) for _ in range(NB)]
    res_host = [torch.empty(N * 5, dtype=torch.uint8).pin_memory() for _ in range(NB)]
    ev_in = [torch.cuda.Event() for _ in range(NB)]
    ev_k = [torch.cuda.Event() for _ in range(NB)]
    ev_out = [torch.cuda.Event() for _ in range(NB)]
    e2e_state = {"i": 0}

    def e2e_step():
        b = e2e_state["i"] % NB
        e2e_state["i"] += 1
        with torch.cuda.stream(h2d):
            h2d.wait_event(ev_k[b])                       # the kernel that last read pts_in[b] is done
            pts_in[b].copy_(pts_host, non_blocking=True)
            ev_in[b].record()
        main.wait_event(ev_in[b])
        main.wait_event(ev_out[b])                        # the read-back of the step that last used outs[b] is done
        ops.query_points(scene, mlp, pts_in[b], want_rgb=False, out=outs[b])
        ev_k[b].record()
        with torch.cuda.stream(d2h):
            d2h.wait_event(ev_k[b])
            res_host[b].copy_(small[b], non_blocking=True)
            ev_out[b].record()

    def e2e_fence():
        main.wait_stream(h2d); main.wait_stream(d2h)
        fence()

    for _ in range(3):
        e2e_step()
    e2e_fence()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    main.wait_stream(d2h)
    e1.record()
    e2e_fence()
    wall_ms = (time.perf_counter() - t0) * 1e3
    note("end-to-end done")
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * N / (e2e_ms / args.steps * 1e-3)

    line = None
    if rank == 0:
        ntex = unique_texels(pts_np, K[0], HF, WF)
        # Algorithmic bytes of the dominant kernel per launch: per voxel its 32-byte record in,
        # density + 64 features out (the frustum mask is written by the sort); per touched texel its row of the map
        # the kernel reads -- fp16, 128 projected channels on the tensor-core path (fp32, 256 channels on the fp32 path).
        if args.precision == "fp16":
            algo_bytes = N * (32 + 4 + 4 * (D_OUT - 1)) + ntex * 128 * 2
            dom = "field_bin_kernel"
        else:
            algo_bytes = N * (12 + 4 + 4 * (D_OUT - 1) + 1) + ntex * C_FEAT * 4
            dom = "field_simt_kernel"
        t_kernel = kernel_ms * 1e-3 if args.precision == "fp16" else ms_step * 1e-3
        hbm_ach = algo_bytes / t_kernel / 1e9
        tc_ach = N * FLOP_PER_POINT / t_kernel / 1e12
        traffic = measured_traffic(dom)
        roof_hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": pk["hbm"], "unit": "GB/s", "frac": hbm_ach / pk["hbm"],
                    "traffic": traffic, "algorithmic_bytes": algo_bytes, "unique_texels": ntex, "peak_source": pk["src"]}
        roof_tc = {"bound": "tensor", "achieved": tc_ach, "peak": pk["tc_burst"], "unit": "TFLOP/s",
                   "frac": tc_ach / pk["tc_burst"], "traffic": traffic, "algorithmic_flops": N * FLOP_PER_POINT,
                   "peak_source": pk["src"] + " (burst: kernel timed alone)"}
        # The tile kernel is bound by memory-side work (records in, 260 B/voxel out, map tiles through L2): its tensor
        # work is ~1/3 of the reference's 92 160 FLOP/voxel because the map is pre-projected once per encode, so the HBM
        # roof is the honest one; the tensor fraction (reference FLOPs over the same time) is reported beside it.
        primary = roof_hbm
        roof_hbm["kernel"] = roof_tc["kernel"] = dom
        roof_hbm["kernel_ms"] = roof_tc["kernel_ms"] = t_kernel * 1e3
        roof_tc["note"] = "reference-algorithm FLOPs (2*(295*128+128*65) per voxel) over the kernel time"
        line = {"metric": "ssc_voxel_query_throughput", "value": value, "unit": "voxels/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "f32",
                "data": "synthetic", "config": workload_config(args.precision),
                "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": N * 12,
                        "d2h_bytes_per_step": N * 5, "ms_per_step": e2e_ms / args.steps,
                        "note": "every step: pinned xyz host->device, query, density grid + frustum mask device->host "
                                "(the 64-d features stay on the device for expand_dim / the SSC head, as in the reference); "
                                "copies and kernels of consecutive steps overlap on three streams"},
                "gpu_launches": int(launches), "clocks": clk.summary(),
                "all_gather": gather_how if world > 1 else None,
                "roofline": primary, "roofline_hbm": roof_hbm, "roofline_tensor": roof_tc,
                "featmap_pack_ms": pack_ms, "featmap_project_ms": project_ms, "sorted_reuse": sorted_reuse}

    # ---- full-image render (122 880 rays x 64 samples), reported in the same line ---------------------
    if not args.no_render:
        # the view is 100 bytes (pose + intrinsics): its rays are generated on the device inside every step (sd_gen_rays)
        view_c2w = torch.from_numpy(syn.view_pose_c2w(1).astype(np.float32)).to(dev)[None]
        view_K = torch.from_numpy(np.asarray(K[0], np.float32)).to(dev)[None]
        rays = torch.empty((RENDER_R, 11), device=dev)
        lin = torch.linspace(0, 1 - 1.0 / RENDER_K, RENDER_K, device=dev)
        u = torch.rand((RENDER_R, RENDER_K), device=dev, generator=g)

        def render_step():
            ops.gen_rays(view_c2w, view_K, syn.IMG_H, syn.IMG_W, syn.Z_NEAR, syn.Z_FAR, out=rays)
            z = ops.sample_coarse(rays, u, lin, True)
            return ops.render_pass(scene_rgb, mlp, rays, z, per_sample=False)

        for _ in range(3):
            render_step()
        fence()
        n_r = max(3, args.steps // 2)
        e0.record()
        for _ in range(n_r):
            render_step()
        e1.record(); fence()
        r_ms = e0.elapsed_time(e1) / n_r
        if line is not None:
            line["render"] = {"workload": "full 192x640 image from a stereo-offset view (rays generated on the device from pose + "
                                          "intrinsics every step), 64 coarse samples/ray, "
                                          "per-ray outputs depth+64-d+rgb"
                                          + (" (projected map: 128-channel gather)" if args.precision == "fp16" else ""),
                              "msamples_per_s": RENDER_R * RENDER_K / (r_ms * 1e-3) / 1e6, "ms": r_ms,
                              "tensor_frac": RENDER_R * RENDER_K * FLOP_PER_POINT / (r_ms * 1e-3) / 1e12 / pk["tc_burst"]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_voxels(feat_nchw.cpu().numpy(), mlp_w, pts_np)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
