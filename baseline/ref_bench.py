"""BASELINE INFRASTRUCTURE: times the UNMODIFIED reference's PyTorch path on the host CPU (bench.py's reference arm and
``cpu_baseline`` with ``kind: "reference"``).  Needs the reference tree (baseline/install_ref.py -> baseline/_ref, or
/root/reference); never imported by the product package."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import ref_env


class _FakeEncoder(torch.nn.Module):
    """cf. EncoderDummy (training/trainer_overfit.py:21-30): the seeded map in place of the ViT (out of scope, no weights)."""

    def __init__(self, feat):
        super().__init__()
        self.latent_size, self.extra_outs = feat.shape[1], 0
        self._feat = feat

    def forward(self, x, ground_truth=False):
        if ground_truth:
            return [torch.zeros(x.shape[0], 8, 2, 2)]
        return [self._feat]


def available() -> bool:
    return ref_env.reference_root() is not None


def build_net(feat_nchw: np.ndarray, mlp_w, K: np.ndarray):
    """The reference's own BTSNet / ResnetFC / PositionalEncoding (scenedino/models/bts.py, ...) on the CPU."""
    bts = ref_env.import_reference_module("scenedino.models.bts")
    heads = ref_env.import_reference_module("scenedino.models.prediction_heads")
    pe = ref_env.import_reference_module("scenedino.common.positional_encoding")
    feat = torch.from_numpy(np.ascontiguousarray(feat_nchw))
    conf = {"predict_dino": True, "dino_dims": mlp_w[2].shape[0] - 1, "inv_z": True, "learn_empty": False, "code_mode": "z"}
    code = pe.PositionalEncoding.from_conf({"num_freqs": 6, "freq_factor": 1.5, "include_input": True}, d_in=3)
    enc = _FakeEncoder(feat)
    head = heads.make_head({"type": "resnet", "name": "normal_head", "args": {"n_blocks": 0, "d_hidden": 128}},
                           enc.latent_size + code.d_out, mlp_w[2].shape[0])
    with torch.no_grad():
        head.lin_in.weight.copy_(torch.from_numpy(mlp_w[0])); head.lin_in.bias.copy_(torch.from_numpy(mlp_w[1]))
        head.lin_out.weight.copy_(torch.from_numpy(mlp_w[2])); head.lin_out.bias.copy_(torch.from_numpy(mlp_w[3]))
    net = bts.BTSNet(conf, enc, code, {"normal_head": head}, None).eval()
    Kt = torch.from_numpy(K.astype(np.float32))[None, None]
    with torch.no_grad():
        net.encode(torch.zeros(1, 1, 3, 8, 8), Kt, torch.eye(4)[None, None], ids_encoder=[0], ids_render=[0])
    net.set_scale(0)
    return net


def time_voxel_query(net, pts: np.ndarray, budget_s: float = 12.0):
    """voxels/s of ``net.forward(xyz)`` (models/bts.py:476-595) on a strided sample of ``pts`` sized to ~budget_s (at most one
    524 288-voxel chunk, the size the reference itself evaluates per call, sscbench/evaluate_model_sscbench.py:675-717)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)

    def sample(n):
        return torch.from_numpy(np.ascontiguousarray(pts[:: max(1, len(pts) // n)][:n]))[None]

    with torch.no_grad():
        net(sample(2048))
        n, dt = 8192, 0.0
        for n in (8192, 65536):                       # the per-voxel cost grows with the sample (cache misses on the 503 MB map)
            sub = sample(n)
            t0 = time.perf_counter(); net(sub); dt = time.perf_counter() - t0
            if dt > budget_s / 3:
                break
        n2 = int(min(524288, len(pts), max(n, n * budget_s / max(dt, 1e-6))))
        n2 = max(n, (n2 // 4096) * 4096)
        sub = sample(n2)
        t0 = time.perf_counter(); net(sub); dt = time.perf_counter() - t0
    return dict(value=sub.shape[1] / dt, unit="voxels/s", cores=torch.get_num_threads(), kind="reference",
                sample=f"{sub.shape[1]} voxels strided over the 256x256x32 grid through the unmodified reference "
                       f"BTSNet.forward (PyTorch {torch.__version__}, CPU, fp32), {dt:.2f} s")
