#!/usr/bin/env python
"""Benchmark of the feature-field query-and-render hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp16|fp32] [--impl reference]

Headline workload (BASELINE.json configs[1]): the SSCBench voxel-grid query -- 256 x 256 x 32 voxels @ 0.2 m projected
into one 192 x 640 view whose DINO ViT-B/8 feature map is 256 x 384 x 1280, MLP head 295 -> 128 -> 65.  One "step" = one
pass of the hot path over the whole grid (2 097 152 voxels): sd_query_points -> sigma [N] + 64-d features [N,64] + frustum
mask [N].  Synthetic data: seeded random feature map and random-init (kaiming) head weights.

The one JSON line carries, besides the contract's keys (value, e2e, roofline, cpu_baseline, clocks, gpu_launches):
  e2e            the same query through the reference-shaped API, ``BTSNet.forward(xyz)``, with the points coming from
                 pinned host memory and sigma + mask going back to it every step
  per_frame      what ONE FRAME of the SSC evaluation costs (sscbench/evaluate_model_sscbench.py:690-756): a NEW feature
                 map every step -> pack + project (BTSNet.encode) -> query of the (fixed) grid -> expansion + SSC head ->
                 sigma + label back on the host; through ``BTSNet.forward(predict_segmentation=True)``
  renders        BASELINE configs 1, 3 and 4 through ``NeRFRenderer`` (sd_render_rays): Msamples/s, tensor roofline, e2e
  fp32           the rel-1e-4 mode (CUDA-core parity path) on a slab of the grid
  fp32_tc        the rel-1e-4 mode on the tensor cores (SD_MLP_F32_TC: fp16 hi/lo operand pairs, three products) on the whole grid
  strong_scaling (N > 1) ONE grid split into x-slabs across the ranks + all-gather of sigma and mask

N > 1 (torchrun, one rank per GPU): ``value`` is weak scaling -- every rank queries one full grid against its own replica
of the map (frames are independent), then all ranks all-gather the density grid and the frustum mask; time = max over
ranks.

``--impl reference``: the reference's own PyTorch CPU path (baseline/_ref, the unmodified reference tree; the C/OpenMP
oracle port when that tree is absent) on all host cores, rank 0 only, on a bounded sample.
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
HF2, WF2 = 192, 640                      # DINOv2 ViT-B/14 variant
D_IN, D_HID, D_OUT = 295, 128, 65
FLOP_PER_POINT = 2 * (D_IN * D_HID + D_HID * D_OUT)        # 92 160
FLOP_PER_POINT_768 = 2 * (D_IN * D_HID + D_HID * 769)      # 272 384
FLOP_EXPAND = 2 * (64 * 128 + 128 * 768)                   # 212 992 per vector
FLOP_SSC_HEAD = 2 * (768 * 64 + 768 * 768 + 768 * 64)      # 1 376 256 per voxel ("1.38 MFLOP")


def measured_traffic(kernel):
    """dram bytes per launch of `kernel` from the latest committed ncu summary (profiles/*_traffic.json) and the commit
    it was profiled at -- a constant read from a file, NOT a measurement of this run -- or None."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        try:
            d = json.load(open(f))
        except (OSError, ValueError):
            continue
        if kernel in d:
            best = (d[kernel], d.get("commit", "unknown"), os.path.basename(f))
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


def cpu_port_voxels(feat_nchw, mlp_w, pts, budget_s=12.0):
    """Times the CPU oracle (C / OpenMP restatement, all host cores) on a strided sample of the grid."""
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


def cpu_reference_voxels(feat_nchw, mlp_w, pts, budget_s=12.0, net_cache={}):
    """The reference's own CPU path where its tree travelled with the snapshot (kind "reference": the unmodified PyTorch code,
    baseline/ref_bench.py), else the oracle port (kind "port")."""
    try:
        from baseline import ref_bench
        if ref_bench.available():
            if "net" not in net_cache:
                net_cache["net"] = ref_bench.build_net(feat_nchw, mlp_w, syn.kitti360_K())
            return ref_bench.time_voxel_query(net_cache["net"], pts, budget_s)
    except Exception as exc:          # noqa: BLE001 -- the baseline must not take the benchmark down
        print(f"[bench] reference CPU path unavailable ({type(exc).__name__}: {exc}); timing the oracle port", file=sys.stderr)
    return cpu_port_voxels(feat_nchw, mlp_w, pts, budget_s)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    feat = syn.make_feature_map(1, C_FEAT, HF, WF)
    mlp_w = syn.make_mlp(0)
    pts = syn.ssc_voxel_grid(GRID)
    vals, last = [], None
    per = max(2.0, 40.0 / (args.warmup + args.steps))
    for i in range(args.warmup + args.steps):
        last = cpu_reference_voxels(feat, mlp_w, pts, budget_s=per)
        if i >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals))
    last["value"] = v
    n_sample = int(last["sample"].split()[0])
    line = {"impl": "reference", "metric": "ssc_voxel_query_throughput", "value": v, "unit": "voxels/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_sample / v * 1e3,
            "step_unit": f"one step = {n_sample} voxels (a bounded strided sample of the grid), not the whole grid",
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config("fp16", binned=args.layout == "binned"), "cpu_baseline": last,   # (the native arm's config)
            "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(precision, binned=False):
    return {"output_layout": ("sigma + frustum mask in the caller's order; 64-d feature rows in texel-bin order + perm[] (row r belongs "
                              "to point perm[r]) -- the layout the fused SSC head (sd_ssc_head) consumes; `other_layout` has the "
                              "caller-order variant" if binned else "everything in the caller's order"),
            "workload": "SSC voxel-grid query 256x256x32 @0.2m (51.2 m), DINO ViT-B/8 map 256x384x1280, "
                        "MLP 295->128->65, outputs sigma+64-d features+mask",
            "path": ("texel sort + projected-map tile kernel (tcgen05 interpolation from TMA tiles)" if precision == "fp16"
                     else "fp32 CUDA-core parity path"),
            "voxels_per_step": GRID[0] * GRID[1] * GRID[2], "feature_map": [C_FEAT, HF, WF],
            "mlp": [D_IN, D_HID, D_OUT],
            "l2": "working set per step (map + 25 MB points + 545 MB outputs) exceeds the 126 MB L2; no explicit flush",
            "launch": "one CUDA-graph replay per step (memset + 3 sort kernels + field kernel)" if precision == "fp16" else "direct calls"}


def build_net(sd, torch, feat_holder, dev, precision, d_out=D_OUT, with_head=True, seed=0):
    """scenedino_b200.BTSNet the way a caller builds it: encoder (a module that hands out the current feature map: the ViT
    is out of scope), positional code, ResnetFC head, SemanticHead; weights by state-dict key."""
    class Encoder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.latent_size, self.extra_outs = C_FEAT, 0
            self.dim_reduction = sd.MlpDimReduction(768, 64, 128)

        def forward(self, x, ground_truth=False):
            return [feat_holder["map"]]

        def expand_dim(self, f):
            return self.dim_reduction.transform_expand(f)

    conf = {"predict_dino": True, "dino_dims": d_out - 1, "inv_z": True, "learn_empty": False, "code_mode": "z",
            "sd_precision": precision}
    code = sd.PositionalEncoding.from_conf({"num_freqs": 6, "freq_factor": 1.5, "include_input": True}, d_in=3)
    head = sd.make_head({"type": "resnet", "name": "normal_head", "args": {"n_blocks": 0, "d_hidden": 128}}, C_FEAT + code.d_out, d_out)
    down = sd.SemanticHead.from_conf({"n_classes": 27, "gt_classes": 19, "input_dim": 768, "code_dim": 64}) if with_head else None
    net = sd.BTSNet(conf, Encoder(), code, {"normal_head": head}, None, downstream_head=down)
    net.encode_loss_features = False
    mlp_w, ex = syn.make_mlp(seed, d_out=d_out), syn.make_expand(3)
    state = {"heads.normal_head.lin_in.weight": mlp_w[0], "heads.normal_head.lin_in.bias": mlp_w[1],
             "heads.normal_head.lin_out.weight": mlp_w[2], "heads.normal_head.lin_out.bias": mlp_w[3],
             "encoder.dim_reduction.linear_in.weight": ex[0], "encoder.dim_reduction.linear_in.bias": ex[1],
             "encoder.dim_reduction.linear_out.weight": ex[2], "encoder.dim_reduction.linear_out.bias": ex[3]}
    if with_head:
        hw = syn.make_ssc_head(21)
        state.update({"downstream_head.stego_head.linear_path.0.weight": hw["wl"].reshape(64, 768, 1, 1),
                      "downstream_head.stego_head.linear_path.0.bias": hw["bl"],
                      "downstream_head.stego_head.nonlinear_path.0.weight": hw["wn1"].reshape(768, 768, 1, 1),
                      "downstream_head.stego_head.nonlinear_path.0.bias": hw["bn1"],
                      "downstream_head.stego_head.nonlinear_path.2.weight": hw["wn2"].reshape(64, 768, 1, 1),
                      "downstream_head.stego_head.nonlinear_path.2.bias": hw["bn2"],
                      "downstream_head.stego_cluster_head.cluster_centers": hw["centres"],
                      "downstream_head.stego_cluster_head.pseudo_assignment": hw["lut"]})
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=False)
    return net.to(dev).eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip per_frame / fp32 / strong-scaling objects")
    ap.add_argument("--no-graph", action="store_true", help="time direct library calls instead of CUDA-graph replays")
    ap.add_argument("--e2e-input", default="device-grid", choices=["device-grid", "upload"],
                    help="e2e step input: frame spec (128 B) + voxel grid made on the device, or the 25 MB of points uploaded")
    ap.add_argument("--layout", default="binned", choices=["binned", "caller"],
                    help="fp16 value step: 64-d rows in texel-bin order + perm (sd_query_points_binned, what the fused SSC head "
                         "consumes) or scattered to the caller's order (sd_query_points); the other one is reported beside it")
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
    import scenedino_b200 as sd
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
    f16 = args.precision == "fp16"
    prec = ops.F16 if f16 else ops.FP32
    fdt = torch.float16 if f16 else torch.float32

    # ---- scene: seeded random map (encoder stand-in), camera, head -----------------------------------
    g = torch.Generator(device=dev).manual_seed(1)
    feat_nchw = torch.randn((1, C_FEAT, HF, WF), device=dev, generator=g)
    K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_ms(fn, n=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    feat = ops.featmap_pack(feat_nchw, fdt)            # first call also pays the module load
    pack_ms = timed_ms(lambda: ops.featmap_pack(feat_nchw, fdt), n=3, warm=1)
    scene = ops.Scene(feat=feat[0], K_f=torch.from_numpy(K).to(dev), w2c_f=torch.from_numpy(w2c).to(dev))
    mlp_w = syn.make_mlp(0)
    mlp = ops.Mlp(*mlp_w, device=dev, precision=prec)
    project_ms = None
    if f16:
        # once per encode (like the pack above): the map pushed through the feature columns of the head's first layer;
        # the voxel query then interpolates 128 hidden pre-activations on the tensor cores (field_proj.cu, field_bin.cu)
        scene = scene.project(mlp)
        state_p = {}

        def proj_step():
            state_p["s"] = scene.project(mlp)
        project_ms = timed_ms(proj_step, n=3, warm=2)
        scene = state_p["s"]
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
    binned = f16 and args.layout == "binned"
    perm_bufs = [torch.empty((N,), dtype=torch.int32, device=dev) for _ in range(NB)] if f16 else None
    # the same buffers seen as the outputs of sd_query_points_binned: rows of `dino` in texel-bin order + the permutation
    outs_b = [dict(sigma=outs[i]["sigma"], invalid_features=outs[i]["invalid_features"], dino_binned=outs[i]["dino"],
                   perm=perm_bufs[i]) for i in range(NB)] if f16 else None
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
    def make_graphs(bin_layout):
        if not f16 or args.no_graph:
            return None
        return [ops.QueryGraph(scene, mlp, pts, outs_b[i] if bin_layout else outs[i], binned_out=bin_layout) for i in range(NB)]

    graphs = make_graphs(binned)
    mode = {"binned": binned}

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
        elif mode["binned"]:
            ops.query_points_binned(scene, mlp, pts, out=outs_b[b])
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

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

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
        ms = max_over_ranks(ms)
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
    other_layout = None
    if f16:
        def reuse_call():
            if binned:
                ops.query_points_binned(scene, mlp, pts, out=outs_b[0], reuse_sorted=True)
            else:
                ops.query_points_sorted(scene, mlp, pts, outs[0])
        step(timed=False) if graphs is None else (ops.query_points_binned(scene, mlp, pts, out=outs_b[0]) if binned
                                                  else ops.query_points(scene, mlp, pts, want_rgb=False, out=outs[0]))
        for _ in range(3):
            reuse_call()
        fence()
        e0.record()
        for _ in range(args.steps):
            reuse_call()
        e1.record(); fence()
        sr_ms = e0.elapsed_time(e1) / args.steps
        sorted_reuse = {"value": world * N / (sr_ms * 1e-3), "unit": "voxels/s", "ms_per_step": sr_ms,
                        "note": "tile kernel only, the sort of the (unchanged) points is reused "
                                "(sd_query_points_binned(reuse_sorted) / sd_query_points_sorted)"}
        # ---- the other output layout, same timing recipe (graph replays for the step, library events for the kernel) --
        if world == 1 and not args.no_extras:
            del graphs
            mode["binned"] = not binned
            graphs = make_graphs(not binned)
            k_events.clear()
            for _ in range(3):
                step()
            for _ in range(5):
                step(timed=True)
            fence()
            ok_ms = float(np.mean([a.elapsed_time(b_) for a, b_ in k_events]))
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record(); fence()
            o_ms = e0.elapsed_time(e1) / args.steps
            other_layout = {"layout": "binned" if not binned else "caller", "ms_per_step": o_ms, "value": N / (o_ms * 1e-3),
                            "unit": "voxels/s", "kernel_ms": ok_ms}
            mode["binned"] = binned
    del graphs
    torch.cuda.empty_cache()

    # ---- end to end through the reference-shaped API: pinned host points -> device -> BTSNet.forward(xyz) -> density
    #      grid + mask back to the host.  Every step moves ITS OWN inputs host->device and its results device->host; the
    #      three legs run on three streams with double buffers (PCIe is full duplex). ----------------------------------
    holder = {"map": feat_nchw}
    net = build_net(sd, torch, holder, dev, args.precision)
    Kt = torch.from_numpy(K).to(dev)[None]
    eye = torch.eye(4, device=dev)[None, None]
    dummy_img = torch.zeros(1, 1, 3, 8, 8, device=dev)

    def encode():
        net.encode(dummy_img, Kt, eye, ids_encoder=[0], ids_render=[0], images_alt=dummy_img)
        net.set_scale(0)

    encode()
    h2d, d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main_s = torch.cuda.current_stream()
    pts_in = [torch.empty((1, N, 3), device=dev) for _ in range(NB)]
    res_host = [torch.empty(N * 5, dtype=torch.uint8).pin_memory() for _ in range(NB)]
    res_dev = [torch.empty(N * 5, dtype=torch.uint8, device=dev) for _ in range(NB)]
    ev_in = [torch.cuda.Event() for _ in range(NB)]
    ev_k = [torch.cuda.Event() for _ in range(NB)]
    ev_out = [torch.cuda.Event() for _ in range(NB)]
    e2e_state = {"i": 0}

    # The step's inputs.  "device-grid" (default): what changes from frame to frame in the SSC loop is the camera, not the
    # grid -- the frame's spec (lidar -> camera matrix, origin, voxel size: 128 B) goes host -> device from pinned memory and
    # the 2 097 152 voxel centres are made on the device (sd_gen_voxel_grid, bit-identical to the host construction of
    # sscbench/evaluate_model_sscbench.py:270-278).  "upload": the 25 MB of points travel over PCIe every step (round 1's
    # e2e; eight ranks doing that saturate the host's memory system: 30 % scaling efficiency at N = 8).
    T_v2c = syn.velo_to_cam()
    spec_host = torch.zeros(16, dtype=torch.float64).pin_memory()
    spec_host[:12] = torch.from_numpy(np.ascontiguousarray(T_v2c[:3, :4]).reshape(-1))
    spec_host[12:15] = torch.tensor([0.0, -25.6, -2.0], dtype=torch.float64)
    spec_host[15] = 0.2
    spec_dev = torch.empty(16, dtype=torch.float64, device=dev)
    upload = args.e2e_input == "upload"

    def e2e_step():
        b = e2e_state["i"] % NB
        e2e_state["i"] += 1
        if upload:
            with torch.cuda.stream(h2d):
                h2d.wait_event(ev_k[b])                   # the kernel that last read pts_in[b] is done
                pts_in[b][0].copy_(pts_host, non_blocking=True)
                ev_in[b].record()
            main_s.wait_event(ev_in[b])
        else:
            spec_dev.copy_(spec_host, non_blocking=True)
            ops.gen_voxel_grid(T_v2c, dims=GRID, out=pts_in[b][0])
        main_s.wait_event(ev_out[b])                      # the read-back of the step that last used res_dev[b] is done
        with torch.no_grad():
            _, invalid, sigma, _, _ = net(pts_in[b], only_density=True)
        res_dev[b][:N * 4].view(torch.float32).copy_(sigma.reshape(-1))
        res_dev[b][N * 4:].copy_(invalid.reshape(-1))     # (only_density: invalid = the frustum mask as floats, bts.py:570-572)
        ev_k[b].record()
        with torch.cuda.stream(d2h):
            d2h.wait_event(ev_k[b])
            res_host[b].copy_(res_dev[b], non_blocking=True)
            ev_out[b].record()

    def e2e_fence():
        main_s.wait_stream(h2d); main_s.wait_stream(d2h)
        fence()

    for _ in range(3):
        e2e_step()
    e2e_fence()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    issue_ms = (time.perf_counter() - t0) * 1e3          # host time to issue the steps (no wait on the device)
    main_s.wait_stream(d2h)
    e1.record()
    e2e_fence()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_dev_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_issue_ms = max_over_ranks(issue_ms) / args.steps
    note("end-to-end done")
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
    e2e_value = world * N / (e2e_ms / args.steps * 1e-3)
    del pts_in, res_dev
    torch.cuda.empty_cache()

    line = None
    if rank == 0:
        ntex = unique_texels(pts_np, K[0], HF, WF)
        # Algorithmic bytes of the dominant kernel per launch: per voxel its 12-byte
        # point in, density + 64 features out (the frustum mask is written by the sort); per touched texel its row of the
        # map the kernel reads -- fp16, 128 projected channels on the tensor-core path (fp32, 256 channels on the fp32 path).
        if f16:
            algo_bytes = N * (12 + 4 + 4 * (D_OUT - 1)) + ntex * 128 * 2
            dom = "field_bin_kernel"
        else:
            algo_bytes = N * (12 + 4 + 4 * (D_OUT - 1) + 1) + ntex * C_FEAT * 4
            dom = "field_simt_kernel"
        t_kernel = kernel_ms * 1e-3 if f16 else ms_step * 1e-3
        hbm_ach = algo_bytes / t_kernel / 1e9
        tc_ach = N * FLOP_PER_POINT / t_kernel / 1e12
        tr = measured_traffic(dom + ("_binned" if f16 and binned else ""))
        traffic = tr[0] if tr else None
        roof_hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": pk["hbm"], "unit": "GB/s", "frac": hbm_ach / pk["hbm"],
                    "traffic": traffic,
                    "traffic_source": (f"constant from profiles/{tr[2]} (ncu --set full at commit {tr[1]}), not a measurement of this run"
                                       if tr else None),
                    "algorithmic_bytes": algo_bytes, "unique_texels": ntex, "peak_source": pk["src"],
                    "step_frac": algo_bytes / (ms_step * 1e-3) / 1e9 / pk["hbm"],
                    "step_frac_note": "the same bytes over the whole step (texel sort + tile kernel): what a caller waits for"}
        roof_tc = {"bound": "tensor", "achieved": tc_ach, "peak": pk["tc_burst"], "unit": "TFLOP/s",
                   "frac": tc_ach / pk["tc_burst"], "traffic": traffic, "algorithmic_flops": N * FLOP_PER_POINT,
                   "peak_source": pk["src"] + " (burst: kernel timed alone)"}
        # The tile kernel is bound by memory-side work (records in, 260 B/voxel out, map tiles through L2): its tensor
        # work is ~1/3 of the reference's 92 160 FLOP/voxel because the map is pre-projected once per encode, so the HBM
        # roof is the honest one; the tensor fraction (reference FLOPs over the same time) is reported beside it.
        if f16:
            roof_hbm["layout"] = ("binned: 64-d rows in texel-bin order by TMA tile stores + perm (sd_query_points_binned)" if binned
                                  else "caller order: scattered 16-byte stores (sd_query_points)")
            if other_layout is not None:       # the other output layout of the same kernel, same bytes, same recipe
                other_layout["roofline_frac"] = algo_bytes / (other_layout["kernel_ms"] * 1e-3) / 1e9 / pk["hbm"]
                other_layout["step_frac"] = algo_bytes / (other_layout["ms_per_step"] * 1e-3) / 1e9 / pk["hbm"]
        roof_hbm["kernel"] = roof_tc["kernel"] = dom
        roof_hbm["kernel_ms"] = roof_tc["kernel_ms"] = t_kernel * 1e3
        roof_tc["note"] = "reference-algorithm FLOPs (2*(295*128+128*65) per voxel) over the kernel time"
        line = {"metric": "ssc_voxel_query_throughput", "value": value, "unit": "voxels/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16" if f16 else "f32",
                "data": "synthetic", "config": workload_config(args.precision, binned),
                "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": N * 12 if upload else 128,
                        "input": ("xyz [N,3] fp32 uploaded from pinned host memory every step" if upload else
                                  "frame spec (lidar->camera matrix, origin, voxel size: 128 B) from pinned host memory; the voxel "
                                  "centres are generated on the device (sd_gen_voxel_grid, bit-identical to the host grid)"),
                        "d2h_bytes_per_step": N * 5, "ms_per_step": e2e_ms / args.steps,
                        "device_ms": e2e_dev_ms, "host_issue_ms": e2e_issue_ms,
                        "api": "scenedino_b200.BTSNet.forward(xyz, only_density=True) (models/bts.py:476-595), fresh outputs and "
                               "workspace per call, full texel sort per call",
                        "note": "every step: inputs host->device (see `input`), BTSNet.forward, density grid + frustum mask device->host "
                                "(the 64-d features stay on the device for expand_dim / the SSC head, as in the reference); "
                                "copies and kernels of consecutive steps overlap on three streams"},
                "gpu_launches": int(launches), "clocks": clk.summary(),
                "all_gather": gather_how if world > 1 else None,
                "roofline": roof_hbm, "roofline_hbm": roof_hbm, "roofline_tensor": roof_tc,
                "featmap_pack_ms": pack_ms, "featmap_project_ms": project_ms, "sorted_reuse": sorted_reuse,
                "other_layout": other_layout}

    # ---- one frame of the SSC evaluation: NEW map -> encode (pack + project) -> query of the fixed grid -> expansion +
    #      SSC head -> sigma + label on the host.  Through BTSNet.encode / BTSNet.forward(predict_segmentation=True). -----
    if f16 and not args.no_extras:
        maps = [feat_nchw, torch.randn((1, C_FEAT, HF, WF), device=dev, generator=g)]
        net.static_query, net.materialize_dino_full, net.one_hot_seg = True, False, False
        xyz_dev = pts[None]
        out_host = [torch.empty(N * 5, dtype=torch.uint8).pin_memory() for _ in range(NB)]
        out_dev = [torch.empty(N * 5, dtype=torch.uint8, device=dev) for _ in range(NB)]
        ev_f = [torch.cuda.Event() for _ in range(NB)]
        ev_o = [torch.cuda.Event() for _ in range(NB)]
        fstate = {"i": 0}

        def frame():
            b = fstate["i"] % NB
            holder["map"] = maps[fstate["i"] % 2]
            fstate["i"] += 1
            encode()
            with torch.no_grad():
                _, _, sigma, seg = net(xyz_dev, predict_segmentation=True, prediction_mode="stego_kmeans")
            main_s.wait_event(ev_o[b])                       # the read-back that last used out_dev[b] is done
            out_dev[b][:N * 4].view(torch.float32).copy_(sigma.reshape(-1))
            out_dev[b][N * 4:].copy_(seg.reshape(-1))          # (labels as uint8: one_hot_seg = False; the reference returns them one-hot)
            ev_f[b].record()
            with torch.cuda.stream(d2h):                     # the read-back of frame i overlaps the kernels of frame i + 1
                d2h.wait_event(ev_f[b])
                out_host[b].copy_(out_dev[b], non_blocking=True)
                ev_o[b].record()

        frame(); frame(); fence()
        nl0 = _abi.launch_count()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            frame()
        main_s.wait_stream(d2h)
        e1.record(); fence()
        pf_wall = (time.perf_counter() - t0) * 1e3 / args.steps
        pf_dev = e0.elapsed_time(e1) / args.steps
        pf_ms = max_over_ranks(max(pf_dev, pf_wall))
        # where the host time goes (no device sync inside: pure issue time of the Python layer)
        th = {"encode": 0.0, "forward": 0.0}
        for _ in range(args.steps):
            holder["map"] = maps[0]
            ta = time.perf_counter(); encode(); tb = time.perf_counter()
            with torch.no_grad():
                net(xyz_dev, predict_segmentation=True, prediction_mode="stego_kmeans")
            tc = time.perf_counter()
            th["encode"] += (tb - ta) * 1e3 / args.steps; th["forward"] += (tc - tb) * 1e3 / args.steps
        fence()
        pf_launch = (_abi.launch_count() - nl0) / args.steps
        # breakdown with the functional layer (same kernels), kernel-only
        hd = ops.SscHead(syn.make_expand(3), syn.make_ssc_head(21), device=dev)
        seg_o = dict(seg=torch.empty((N,), dtype=torch.uint8, device=dev))
        ops.query_points_binned(scene, mlp, pts, out=outs_b[0])
        q_ms = timed_ms(lambda: ops.query_points_binned(scene, mlp, pts, out=outs_b[0], reuse_sorted=True))
        h_ms = timed_ms(lambda: ops.ssc_head(hd, outs_b[0]["dino_binned"], want_scores=False, perm=outs_b[0]["perm"], out=seg_o))
        ex_mlp = ops.Mlp(*syn.make_expand(3), device=dev)
        ex_ms = timed_ms(lambda: ops.expand_dim(ex_mlp, outs[0]["dino"][: N // 4], precision=ops.F16), n=3, warm=1) * 4
        if line is not None:
            dev_sum = pack_ms + project_ms + q_ms + h_ms
            line["per_frame"] = {
                "what": "one SSC frame: new 256x384x1280 map -> BTSNet.encode (pack fp32 planar -> fp16 channels-last, project "
                        "through lin_in) -> BTSNet.forward(grid, predict_segmentation=True) on the fixed 2 097 152-voxel grid (texel "
                        "sort kept: static_query = True, the caller vouches for unchanged points and camera) -> fused expansion + STEGO head + cosine argmax + LUT -> sigma fp32 + label u8 to "
                        "pinned host memory",
                "ms": pf_ms, "device_ms": pf_dev, "host_issue_ms": th, "voxels_per_s": world * N / (pf_ms * 1e-3), "launches_per_frame": pf_launch,
                "d2h_bytes": N * 5, "h2d_bytes": 0,
                "note_ms": "max(device time, host wall time) per frame: includes the torch glue (output allocation, packing sigma + "
                           "labels into one buffer) around the library calls; the read-back of frame i overlaps frame i + 1",
                "kernels_ms": {"featmap_pack": pack_ms, "field_project": project_ms, "query_binned_sorted(field_bin)": q_ms,
                               "expand+ssc_head(ssc_head_kernel)": h_ms, "sum": dev_sum},
                "ssc_head": {"ms": h_ms, "executed_tflops": N * 364544 / (h_ms * 1e-3) / 1e12,
                             "reference_algorithm_tflops": N * (FLOP_EXPAND + FLOP_SSC_HEAD) / (h_ms * 1e-3) / 1e12,
                             "tensor_frac_executed": N * 364544 / (h_ms * 1e-3) / 1e12 / pk["tc_burst"],
                             "note": "the folded head executes 0.36 MFLOP/voxel of the reference's 1.59 (expansion + head); "
                                     "fraction of the measured bf16 cuBLAS peak on the executed FLOPs"},
                "unfused_expand_dim_ms": ex_ms,
                "headline_ratio_uses": "the driver's e2e ratio uses `e2e` (plain query through BTSNet.forward), not this object"}
        net.static_query, net.materialize_dino_full, net.one_hot_seg = False, True, True
        net.reset_static_query()
        del maps, hd, seg_o, ex_mlp
        torch.cuda.empty_cache()
        note("per-frame done")

    # ---- the rel-1e-4 mode (fp32 map, CUDA-core head) on one 32-voxel-thick x-slab -------------------------------------
    if f16 and not args.no_extras:
        n32 = N // 8
        feat32 = ops.featmap_pack(feat_nchw, torch.float32)
        sc32 = ops.Scene(feat=feat32[0], K_f=scene.K_f, w2c_f=scene.w2c_f)
        o32 = None

        def q32():
            nonlocal o32
            o32 = ops.query_points(sc32, mlp, pts[:n32], want_rgb=False, precision=ops.FP32, out=o32)
        ms32 = timed_ms(q32, n=2, warm=1)
        if line is not None:
            line["fp32"] = {"what": "rel-1e-4 parity mode: fp32 channels-last map, FFMA head (field_simt_kernel), one x-slab of 262 144 voxels",
                            "voxels_per_s": n32 / (ms32 * 1e-3), "ms": ms32, "ffma_tflops": n32 * FLOP_PER_POINT / (ms32 * 1e-3) / 1e12}
        # ---- the same bar ON the tensor cores: SD_MLP_F32_TC (fp16 hi/lo operand pairs, three kind::f16 products per
        #      contraction, fp32 epilogues: field_bin_x3.cu) on the whole grid; its once-per-encode projection beside it --
        st3 = {}

        def p3():
            st3["s"] = sc32.project_x3(mlp)
        proj3_ms = timed_ms(p3, n=2, warm=1)
        sc3 = st3["s"]
        o3 = None

        def q3():
            nonlocal o3
            o3 = ops.query_points(sc3, mlp, pts, want_rgb=False, precision=ops.F32TC, out=o3)
        ms3 = timed_ms(q3, n=max(3, args.steps // 2), warm=2)
        k3 = []
        for _ in range(4):
            ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ka.record(); kb.record()
            _abi.check(_abi.lib().sd_profile_next_kernel(ka.cuda_event, kb.cuda_event), "sd_profile_next_kernel")
            q3(); torch.cuda.synchronize()
            k3.append(ka.elapsed_time(kb))
        k3_ms = float(np.median(k3))
        a64, b64 = o3["sigma"][:n32].double(), o32["sigma"].double()
        d64, e64 = o3["dino"][:n32].double(), o32["dino"].double()
        err = max(float(((a64 - b64).abs() / torch.clamp(b64.abs(), min=float(b64.pow(2).mean().sqrt()))).max()),
                  float(((d64 - e64).abs() / torch.clamp(e64.abs(), min=float(e64.pow(2).mean().sqrt()))).max()))
        if line is not None:
            # executed tensor work: 3 products per contraction of the tile kernel's own algebra (interpolation K = 64 per
            # chunk ~1.6 chunks per tile, code block K = 48, layer 2 K = 128 x N = 80)
            line["fp32_tc"] = {"what": "rel-1e-4 mode on the tensor cores (SD_MLP_F32_TC): texel sort + field_bin_kernel (x3) on the whole "
                                       "2 097 152-voxel grid, fp32 map projected once per encode into an fp16 (hi, lo) pair of maps",
                               "voxels_per_s": N / (ms3 * 1e-3), "ms": ms3, "kernel_ms": k3_ms, "project_x3_ms": proj3_ms,
                               "max_rel_err_vs_fp32_cuda_core_kernel": err,
                               "speedup_vs_fp32_cuda_core_kernel": (N / ms3) / (n32 / ms32),
                               "roofline": {"bound": "tensor", "achieved": N * FLOP_PER_POINT / (k3_ms * 1e-3) / 1e12,
                                            "peak": pk["tc_burst"], "unit": "TFLOP/s",
                                            "frac": N * FLOP_PER_POINT / (k3_ms * 1e-3) / 1e12 / pk["tc_burst"], "traffic": None,
                                            "note": "reference-algorithm FLOPs (92 160 per voxel) over the tile kernel's time against the "
                                                    "fp16 dense peak; the kernel issues 3 fp16 products per contraction, so 1/3 is its ceiling"}}
        del feat32, sc32, o32, sc3, o3, st3
        torch.cuda.empty_cache()

    # ---- strong scaling: ONE grid split into x-slabs (what north_star's partition names), all-gather of sigma + mask ----
    if world > 1 and f16 and not args.no_extras:
        from scenedino_b200.sharding import shard_slice
        sx = shard_slice(GRID[0], rank, world)
        n_loc = (sx.stop - sx.start) * GRID[1] * GRID[2]
        loc = pts[sx.start * GRID[1] * GRID[2]: sx.stop * GRID[1] * GRID[2]].contiguous()
        small_l = torch.empty(n_loc * 5, dtype=torch.uint8, device=dev)
        out_l = dict(sigma=small_l[:n_loc * 4].view(torch.float32), invalid_features=small_l[n_loc * 4:],
                     dino=torch.empty((n_loc, D_OUT - 1), device=dev))
        gath = torch.empty((world, n_loc * 5), dtype=torch.uint8, device=dev)
        qg = ops.QueryGraph(scene, mlp, loc, out_l)

        def sstep():
            qg.replay()
            dist.all_gather_into_tensor(gath, small_l)

        for _ in range(3):
            sstep()
        fence()
        e0.record()
        for _ in range(args.steps):
            sstep()
        e1.record(); fence()
        s_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        if line is not None:
            line["strong_scaling"] = {"what": f"ONE 256x256x32 grid in {world} x-slabs of {sx.stop - sx.start} (sort + tile kernel per slab, "
                                              "graph replay) + NCCL all-gather of sigma + mask (5 B/voxel); max over ranks",
                                      "ms_per_grid": s_ms, "voxels_per_s": N / (s_ms * 1e-3),
                                      "speedup_vs_one_gpu_step": ms_step / s_ms if world == 1 else None,
                                      "one_gpu_ms_reference": "compare with ms_per_step of the N=1 run"}
        del qg, out_l, gath, small_l, loc
        torch.cuda.empty_cache()

    # ---- renders: BASELINE configs 1, 4, 3 through NeRFRenderer (one sd_render_rays call per scene) --------------------
    if not args.no_render and f16:
        renders = {}
        ray_sampler = sd.ImageRaySampler(z_near=syn.Z_NEAR, z_far=syn.Z_FAR, height=syn.IMG_H, width=syn.IMG_W)

        def render_cfg(tag, what, Hf_, Wf_, views, n_coarse, n_fine, d_out, nv_c, ray_subset=None, expand=False):
            hold = {"map": feat_nchw if (Hf_, Wf_) == (HF, WF) else torch.randn((1, C_FEAT, Hf_, Wf_), device=dev, generator=g)}
            rnet = build_net(sd, torch, hold, dev, "fp16", d_out=d_out, with_head=False, seed=0)
            ren = sd.NeRFRenderer.from_conf({"n_coarse": n_coarse, "n_fine": n_fine, "n_fine_depth": 0, "lindisp": True,
                                             "hard_alpha_cap": n_fine > 0})
            ren.nan_check = False                       # (one host sync per pass; the throughput configuration skips it)
            wrapped = ren.bind_parallel(rnet, gpus=None).eval()
            imgs = torch.from_numpy(syn.make_images(2, nv_c)).to(dev)[None]
            Kc = torch.from_numpy(np.broadcast_to(syn.kitti360_K(), (nv_c, 3, 3)).copy()).to(dev)[None]
            c2w = torch.from_numpy(np.stack([syn.view_pose_c2w(v) for v in range(nv_c)])).to(dev)[None]
            rnet.encoder.dim_reduction.precision = "fp16"
            rnet.encode(imgs * 2 - 1, Kc, c2w, ids_encoder=[0], ids_render=list(range(nv_c)), images_alt=imgs)
            rnet.set_scale(0)
            vposes_host = torch.from_numpy(np.stack([syn.view_pose_c2w(v) for v in views])).pin_memory()
            vK = torch.from_numpy(np.broadcast_to(syn.kitti360_K(), (len(views), 3, 3)).copy()).to(dev)[None]
            R_full = len(views) * syn.IMG_H * syn.IMG_W
            sel = None
            if ray_subset is not None:
                sel = torch.randperm(R_full, device=dev, generator=g)[:ray_subset]
            R = R_full if sel is None else ray_subset
            res = {}
            host_out = torch.empty(R * (1 + 3 * nv_c), dtype=torch.float32).pin_memory()

            def rstep(e2e=False):
                vp = vposes_host.to(dev, non_blocking=True)[None] if e2e else rstep.vp
                rays, _ = ray_sampler.sample(None, vp, vK)             # rays of whole views generated on the device
                if sel is not None:
                    rays = rays[:, sel].contiguous()
                with torch.no_grad():
                    out = wrapped(rays)
                lvl = out["fine"] if n_fine > 0 else out["coarse"]
                if expand:
                    res["full"] = rnet.encoder.expand_dim(lvl["dino_features"])
                if e2e:
                    host_out[:R].copy_(lvl["depth"].reshape(-1), non_blocking=True)
                    host_out[R:].copy_(lvl["rgb"].reshape(-1), non_blocking=True)
                return lvl

            rstep.vp = vposes_host.to(dev)[None]
            n_r = max(3, args.steps // 4)
            nl = _abi.launch_count()
            ms_k = timed_ms(rstep, n=n_r, warm=2)
            nl = (_abi.launch_count() - nl) / (n_r + 2)
            ms_e = timed_ms(lambda: rstep(True), n=n_r, warm=1)
            samples = R * (n_coarse + ((n_coarse + n_fine) if n_fine > 0 else 0))
            flop = FLOP_PER_POINT_768 if d_out == 769 else FLOP_PER_POINT
            ach = samples * flop / (ms_k * 1e-3) / 1e12
            renders[tag] = {"workload": what, "rays": R, "samples_per_step": samples, "ms": ms_k,
                            "msamples_per_s": samples / (ms_k * 1e-3) / 1e6, "launches_per_step": nl,
                            "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["tc_burst"], "unit": "TFLOP/s",
                                         "frac": ach / pk["tc_burst"], "traffic": None,
                                         "note": f"reference-algorithm FLOPs ({flop} per sample, SURVEY.md 8d) over the step time "
                                                 "(sampling, field kernel, head2 / expand kernels); the field kernel executes fewer: "
                                                 "the map is pre-projected and, for D > 64, W_out is applied to per-ray sums"},
                            "e2e": {"ms": ms_e, "msamples_per_s": samples / (ms_e * 1e-3) / 1e6,
                                    "h2d_bytes_per_step": len(views) * 64, "d2h_bytes_per_step": R * (1 + 3 * nv_c) * 4,
                                    "note": "poses from pinned host memory, rays generated on the device, depth + rgb images back to the "
                                            "host (the rendered features stay on the device for the loss / expansion)"}}
            del rnet, wrapped, hold
            torch.cuda.empty_cache()

        render_cfg("cfg1", "BASELINE configs[0]: 4096 random rays of the input view x 64 coarse samples, ViT-B/8 map, D=64 "
                           "(~20 us of B200 work: launch-bound, reported for completeness)", HF, WF, [0], 64, 0, D_OUT, 1, ray_subset=4096)
        render_cfg("cfg4", "BASELINE configs[3]: DINOv2 map 256x192x640, full 192x640 image of a stereo-offset view x 32 coarse samples, "
                           "D=64, then expand_dim 64->128->768 + L2 norm of the 122 880 rendered vectors", HF2, WF2, [1], 32, 0, D_OUT, 1,
                   expand=True)
        render_cfg("cfg3", "BASELINE configs[2]: 4 views x 192x640 rays against the view-0 ViT-B/8 map, 64 coarse + 32 fine samples "
                           "(coarse pass 64, fine pass 96), 768-d feature composite, 4 colour views", HF, WF, [0, 1, 2, 3], 64, 32, 769, 4)
        # ---- one training step's render: forward + backward of the unfused differentiable path on
        #      BASELINE configs[0]'s ray batch (4096 rays x 64 coarse samples): gradients to the encoder map and the head ----
        train = None
        try:
            tmap = torch.randn((1, C_FEAT, HF2, WF2), device=dev, generator=g).requires_grad_()
            tnet = build_net(sd, torch, {"map": tmap}, dev, "fp32", with_head=False, seed=0)
            tnet.train()
            tren = sd.NeRFRenderer.from_conf({"n_coarse": 64, "n_fine": 0, "lindisp": True, "hard_alpha_cap": False})
            tren.nan_check = False
            tren.train()
            timg = torch.from_numpy(syn.make_images(2, 1)).to(dev)[None]
            tK = torch.from_numpy(syn.kitti360_K()[None]).to(dev)[None]
            tpose = torch.from_numpy(syn.view_pose_c2w(0)[None]).to(dev)[None]
            tnet.encode(timg * 2 - 1, tK, tpose, ids_encoder=[0], ids_render=[0], images_alt=timg)
            tnet.set_scale(0)
            trays, _ = ray_sampler.sample(None, torch.from_numpy(syn.view_pose_c2w(1)[None]).to(dev)[None], tK)
            trays = trays[:, torch.randperm(trays.shape[1], device=dev, generator=g)[:4096]].contiguous()
            tparams = [tmap] + [p_ for p_ in tnet.heads.parameters()]

            def tstep():
                for p_ in tparams:
                    p_.grad = None
                o = tren(tnet, trays)["coarse"]
                (o["depth"].sum() * 0.01 + o["dino_features"].sum() + o["rgb"].sum()).backward()

            t_ms = timed_ms(tstep, n=max(3, args.steps // 4), warm=2)
            train = {"what": "training-mode render of 4096 rays x 64 coarse samples (DINOv2-sized map), forward + backward through "
                             "BTSNet / NeRFRenderer with autograd on: sd_sample_features, the head as torch modules (library "
                             "GEMMs), sd_composite, sd_composite_bwd, sd_sample_features_bwd; gradients to the encoder map and "
                             "the head weights", "ms": t_ms, "msamples_per_s": 4096 * 64 / (t_ms * 1e-3) / 1e6,
                     "grad_map_abs_sum": float(tmap.grad.abs().sum())}
            del tnet, tren, tmap
        except Exception as exc:      # noqa: BLE001 -- an extra: never takes the line down
            train = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
        if line is not None:
            line["train_step"] = train
            line["renders"] = renders
            line["render"] = {"msamples_per_s": renders["cfg4"]["msamples_per_s"], "ms": renders["cfg4"]["ms"],
                              "tensor_frac": renders["cfg4"]["roofline"]["frac"], "workload": "see renders.cfg4"}
        note("renders done")

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_voxels(feat_nchw.cpu().numpy(), mlp_w, pts_np)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
