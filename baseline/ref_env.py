"""TEST / BASELINE INFRASTRUCTURE -- never imported by the product package.

Makes the UNMODIFIED reference (tum-vision/scenedino) importable from a directory: ``/root/reference`` in the build
container, or ``baseline/_ref`` -- the git-ignored install location ``baseline/install_ref.py`` copies the reference's
Python tree to, so that it travels to the GPU box with the snapshot.  Used by the drop-in tests (the reference's own caller
functions on top of the B200-native classes) and by ``bench.py --impl reference`` (the reference's PyTorch CPU path).

The reference's launch scripts import packages that are not installed here and that the hot path never touches (hydra,
ignite, matplotlib, lpips, kornia, timm, ...; there is no network to install them): a meta-path finder answers those
imports with permissive stand-in modules.  ``dotmap.DotMap`` and ``omegaconf.ListConfig`` are used on the path and get
small real implementations.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.environ.get("SCENEDINO_REFERENCE", ""), "/root/reference", os.path.join(HERE, "_ref")]

STUB_PACKAGES = {"hydra", "ignite", "matplotlib", "lpips", "kornia", "timm", "pykeops", "pydensecrf", "plyfile", "skimage",
                 "open3d", "gradio", "optuna", "dotdict", "pulp", "tensorboardX", "wandb", "seaborn", "PIL_stub", "moviepy", "imageio"}


def reference_root():
    for c in CANDIDATES:
        if c and os.path.isdir(os.path.join(c, "scenedino")):
            return c
    return None


class DotMap(dict):
    """dotmap.DotMap as the renderer uses it (renderer/nerf.py:9,499-509,571)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def toDict(self):
        return {k: (v.toDict() if isinstance(v, DotMap) else v) for k, v in self.items()}


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_stub(f"{cls.__name__}.{name}")


def _make_stub(name):
    class Stub(metaclass=_StubMeta):
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            if len(a) == 1 and not k and callable(a[0]):
                return a[0]                     # used as a decorator factory
            return self

        def __getattr__(self, n):
            if n.startswith("__"):
                raise AttributeError(n)
            return Stub()

        def __iter__(self):
            return iter(())

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    Stub.__name__ = Stub.__qualname__ = name
    return Stub


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = _make_stub(f"{self.__name__}.{name}")
        setattr(self, name, v)
        return v


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUB_PACKAGES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_prepared = None


def prepare():
    """Puts the reference on sys.path and registers the stand-ins; returns the reference root (None if absent)."""
    global _prepared
    root = reference_root()
    if root is None:
        return None
    if _prepared == root:
        return root
    for p in (root, os.path.join(root, "sscbench")):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())       # behind the real finders: installed packages win
    if "dotmap" not in sys.modules:
        try:
            import dotmap  # noqa: F401
        except ImportError:
            m = types.ModuleType("dotmap"); m.DotMap = DotMap; sys.modules["dotmap"] = m
    if "omegaconf" not in sys.modules:
        try:
            import omegaconf  # noqa: F401
        except ImportError:
            m = types.ModuleType("omegaconf")
            m.ListConfig, m.DictConfig, m.OmegaConf = list, dict, _make_stub("OmegaConf")
            import contextlib
            m.open_dict = contextlib.nullcontext
            sys.modules["omegaconf"] = m
    _prepared = root
    return root


def import_reference_module(name):
    """Imports a reference module with its one import-time CUDA call (scenedino/models/bts.py:19) neutralised on CPU."""
    import importlib

    import torch
    if prepare() is None:
        raise RuntimeError("the reference tree is not available (neither /root/reference nor baseline/_ref)")
    _ones, _system, cwd = torch.ones, os.system, os.getcwd()
    if not torch.cuda.is_available():
        torch.ones = lambda *a, **k: _ones(*a, **{kk: v for kk, v in k.items() if kk != "device"})
    os.system = lambda *a, **k: 0          # sscbench/evaluate_model_sscbench.py:78 shells out to nvidia-smi at import
    os.chdir(_prepared)                    # sscbench/gen_voxelgrid_npy.py:102 opens "sscbench/label_maps.yaml" relative to the root
    try:
        return importlib.import_module(name)
    finally:
        torch.ones, os.system = _ones, _system
        os.chdir(cwd)
