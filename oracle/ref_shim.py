"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference (tum-vision/scenedino) from ``/root/reference`` so that
``oracle/make_golden.py`` can run the reference's own hot path on CPU and dump golden vectors.
``/root/reference`` only exists in the build container: nothing on the GPU box may import this.

The reference needs a few third-party packages that are not installed here (no network); they are
not on the hot path, so empty stand-in modules are registered for them.  The one import-time CUDA
call (``scenedino/models/bts.py:19``) is neutralised for the duration of the import.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("SCENEDINO_REFERENCE", "/root/reference")


class _DotMap(dict):
    """Stand-in for dotmap.DotMap as used at renderer/nerf.py:9,499-509,571."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def toDict(self):
        return {k: (v.toDict() if isinstance(v, _DotMap) else v) for k, v in self.items()}


_STUBS = [
    "timm", "timm.models", "timm.models.vision_transformer", "kornia", "pykeops", "pykeops.torch",
    "matplotlib", "matplotlib.pyplot", "pydensecrf", "pydensecrf.densecrf", "pydensecrf.utils",
    "lpips",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scenedino"))


def import_reference():
    """Returns a namespace with the reference classes on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "dotmap" not in sys.modules:
        m = types.ModuleType("dotmap"); m.DotMap = _DotMap; sys.modules["dotmap"] = m
    if "omegaconf" not in sys.modules:
        m = types.ModuleType("omegaconf")
        m.ListConfig = list; m.DictConfig = dict; m.OmegaConf = object
        sys.modules["omegaconf"] = m
    for n in _STUBS:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["timm.models.vision_transformer"].VisionTransformer = object
    sys.modules["pykeops.torch"].LazyTensor = object

    _ones = torch.ones
    if not torch.cuda.is_available():
        torch.ones = lambda *a, **k: _ones(*a, **{kk: v for kk, v in k.items() if kk != "device"})
    try:
        from scenedino.renderer.nerf import NeRFRenderer
        from scenedino.models.bts import BTSNet
        from scenedino.models.prediction_heads import make_head
        from scenedino.common.positional_encoding import PositionalEncoding
        from scenedino.common.ray_sampler import ImageRaySampler
        from scenedino.common.cameras import pinhole
        from scenedino.models.backbones.dino.dim_reduction import MlpDimReduction
    finally:
        torch.ones = _ones
    return types.SimpleNamespace(
        NeRFRenderer=NeRFRenderer, BTSNet=BTSNet, make_head=make_head,
        PositionalEncoding=PositionalEncoding, ImageRaySampler=ImageRaySampler, pinhole=pinhole,
        MlpDimReduction=MlpDimReduction,
    )


class FakeEncoder(torch.nn.Module):
    """Seeded random feature map standing in for the DINO ViT+DPT encoder
    (same trick as the reference's EncoderDummy, scenedino/training/trainer_overfit.py:21-30)."""

    def __init__(self, ref, feat: torch.Tensor, dino_dims: int = 64):
        super().__init__()
        self.latent_size = feat.shape[1]
        self.extra_outs = 0
        self.dino_pca_dim = dino_dims
        self.dim_reduction = ref.MlpDimReduction(768, dino_dims, 128)
        self._feat = feat  # [n, C, Hf, Wf]

    def forward(self, x, ground_truth=False):
        if ground_truth:
            return [torch.zeros(x.shape[0], 8, 2, 2)]
        return [self._feat.clone()]

    def expand_dim(self, f):
        return self.dim_reduction.transform_expand(f)


def build_reference_net(ref, feat, w_in, b_in, w_out, b_out, *, dino_dims=64, learn_empty=False,
                        empty_feature=None, expand_weights=None, conf_extra=None):
    """Builds reference BTSNet + ResnetFC head with the given weights (all torch fp32 CPU)."""
    conf = {"predict_dino": True, "dino_dims": dino_dims, "inv_z": True, "learn_empty": learn_empty,
            "code_mode": "z"}
    if conf_extra:
        conf.update(conf_extra)
    code = ref.PositionalEncoding.from_conf(
        {"num_freqs": 6, "freq_factor": 1.5, "include_input": True}, d_in=3)
    enc = FakeEncoder(ref, feat, dino_dims)
    head = ref.make_head({"type": "resnet", "name": "normal_head",
                          "args": {"n_blocks": 0, "d_hidden": w_in.shape[0]}},
                         enc.latent_size + code.d_out, 1 + dino_dims)
    with torch.no_grad():
        head.lin_in.weight.copy_(w_in); head.lin_in.bias.copy_(b_in)
        head.lin_out.weight.copy_(w_out); head.lin_out.bias.copy_(b_out)
        if expand_weights is not None:
            e = enc.dim_reduction
            e.linear_in.weight.copy_(expand_weights[0]); e.linear_in.bias.copy_(expand_weights[1])
            e.linear_out.weight.copy_(expand_weights[2]); e.linear_out.bias.copy_(expand_weights[3])
    net = ref.BTSNet(conf, enc, code, {"normal_head": head}, None).eval()
    if learn_empty and empty_feature is not None:
        with torch.no_grad():
            net.empty_feature.copy_(empty_feature)
    return net
