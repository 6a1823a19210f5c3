"""scenedino_b200 -- B200-native feature-field query-and-render path of SceneDINO.

Reference-shaped surface:  ``BTSNet`` (models/bts.py), ``NeRFRenderer`` (renderer/nerf.py),
``ImageRaySampler`` (common/ray_sampler.py), ``ResnetFC`` / ``make_head`` / ``PositionalEncoding`` /
``MlpDimReduction`` parameter containers, ``SemanticHead`` (downstream_head/semantic_head.py, inference half).
Functional layer over the C ABI (include/scenedino_b200.h):  ``scenedino_b200.ops``.
"""
from ._abi import SdError, launch_count, lib  # noqa: F401
from .bts import BTSNet  # noqa: F401
from .heads import MlpDimReduction, PositionalEncoding, ResnetFC, make_head  # noqa: F401
from .nerf import DotMap, NeRFRenderer, _RenderWrapper  # noqa: F401
from .ray_sampler import ImageRaySampler, RaySampler  # noqa: F401
from .semantic_head import SemanticHead, make_downstream_head  # noqa: F401
from .install import install  # noqa: F401

__all__ = ["BTSNet", "NeRFRenderer", "ResnetFC", "make_head", "PositionalEncoding", "MlpDimReduction", "ImageRaySampler", "RaySampler", "SemanticHead", "make_downstream_head", "install",
           "DotMap", "SdError", "launch_count", "lib"]
