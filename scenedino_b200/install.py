"""``install()`` -- rebinds the reference's hot-path classes to the B200-native ones, in place.

The reference's entry points (``demo_script.py``, ``eval.py``, ``sscbench/evaluate_model_sscbench.py``) build their model
through ``scenedino.models.make_model`` and their renderer / ray sampler through ``scenedino.renderer.nerf.NeRFRenderer`` and
``scenedino.common.ray_sampler.ImageRaySampler``.  After ``import scenedino_b200; scenedino_b200.install()`` those names -- in
their defining modules AND in every already-imported module that pulled them in with ``from ... import`` -- refer to the
classes of this package, so the scripts run unchanged on the fused CUDA path:

    BTSNet, NeRFRenderer, _RenderWrapper, ImageRaySampler, ResnetFC, make_head, PositionalEncoding, MlpDimReduction,
    SemanticHead, make_downstream_head

The DINO encoder (``scenedino.models.backbones``) is not touched: it stays the reference's PyTorch module.
"""
from __future__ import annotations

import importlib
import sys

REBIND = (
    ("scenedino.models.bts", "BTSNet"),
    ("scenedino.renderer.nerf", "NeRFRenderer"),
    ("scenedino.renderer.nerf", "_RenderWrapper"),
    ("scenedino.common.ray_sampler", "ImageRaySampler"),
    ("scenedino.models.prediction_heads.resnetfc", "ResnetFC"),
    ("scenedino.models.prediction_heads", "make_head"),
    ("scenedino.common.positional_encoding", "PositionalEncoding"),
    ("scenedino.models.backbones.dino.dim_reduction", "MlpDimReduction"),
    ("scenedino.downstream_head.semantic_head", "SemanticHead"),
    ("scenedino.downstream_head", "make_downstream_head"),
)


def install(verbose: bool = False) -> dict:
    """Returns {"module.name": number of bindings replaced}.  Raises ImportError if the reference is not importable."""
    import scenedino_b200 as sd
    report = {}
    for mod_name, attr in REBIND:
        mod = importlib.import_module(mod_name)
        old, new = getattr(mod, attr), getattr(sd, attr)
        if old is new:
            report[f"{mod_name}.{attr}"] = 0
            continue
        n = 0
        for m in list(sys.modules.values()):
            d = getattr(m, "__dict__", None)
            if not isinstance(d, dict):
                continue
            for k, v in list(d.items()):
                if v is old:
                    d[k] = new
                    n += 1
        report[f"{mod_name}.{attr}"] = n
        if verbose:
            print(f"scenedino_b200.install: {mod_name}.{attr} -> scenedino_b200.{attr} ({n} bindings)")
    return report
