"""Times the two built "next" rows of SURVEY 8f on the device (CUDA events, warm, median of 10):
sd_gen_rays on four 376 x 1408 views and sd_expand_dim (fp32 CUDA-core path and tensor-core kernel) on the 64-d
features of an SSC chunk / the whole SSC grid.
    python profiles/time_next_rows.py > gpurun_out/next_rows.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedino_b200 import ops  # noqa: E402
from scenedino_b200 import synthetic as syn  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    dev = "cuda:0"
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
    out = {"peaks": peaks}
    V, H, W = 4, 376, 1408
    c2w = torch.from_numpy(np.stack([syn.view_pose_c2w(v) for v in range(V)]).astype(np.float32)).to(dev)
    K = torch.from_numpy(np.broadcast_to(syn.kitti360_K(), (V, 3, 3)).astype(np.float32).copy()).to(dev)
    rays = torch.empty((V * H * W, 11), device=dev)
    ms = timed(lambda: ops.gen_rays(c2w, K, H, W, 3.0, 80.0, out=rays))
    out["gen_rays"] = {"views": V, "H": H, "W": W, "ms": ms, "bytes": rays.numel() * 4, "GBps": rays.numel() * 4 / ms / 1e6}
    rs = np.random.RandomState(0)
    w = [rs.randn(128, 64) * 0.1, rs.randn(128) * 0.1, rs.randn(768, 128) * 0.1, rs.randn(768) * 0.1]
    mlp = ops.Mlp(*[x.astype(np.float32) for x in w], device=dev)
    for N in (1 << 19, 1 << 21):
        f = torch.randn((N, 64), device=dev)
        full = torch.empty((N, 768), device=dev)
        for tag, prec in (("fp32_simt", ops.FP32), ("f16_tc", ops.F16)):
            if prec == ops.FP32 and N > (1 << 19):
                continue                    # 126 ms per call: measured once, the small size is enough to track it
            ms = timed(lambda: ops.expand_dim(mlp, f, precision=prec), n=5)
            out[f"expand_dim_{tag}_{N}"] = {"N": N, "ms": ms, "tflops": N * 212992 / ms / 1e9,
                                            "hbm_GBps": N * (768 + 64) * 4 / ms / 1e6}
        del f, full
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
