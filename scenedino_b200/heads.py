"""Parameter containers for the decoder head and the positional code.

They mirror the constructor signatures, attribute names and state-dict keys of the reference's
``ResnetFC`` (models/prediction_heads/resnetfc.py:66-135, ``n_blocks = 0`` as in every shipped
config, configs/model/dino_downsampler.yaml:35-41), ``PositionalEncoding``
(common/positional_encoding.py:44-66) and ``MlpDimReduction``
(models/backbones/dino/dim_reduction.py:15-25) so that reference checkpoints load with
``load_state_dict`` -- but their arithmetic runs in libscenedino_b200 (ResnetFC.forward ->
sd_mlp_forward, transform_expand -> sd_expand_dim; the positional code only exists fused inside the
field kernels).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
import threading

import torch
from torch import nn

from . import _abi


def _ptr(t: torch.Tensor | None):
    """Device pointer of a tensor for the C ABI; host tensors are rejected (there is no CPU path)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise _abi.SdError(f"expected a CUDA tensor, got one on {t.device} (scenedino_b200 has no CPU path)")
    return C.c_void_p(t.data_ptr())


def _stream():
    """Current stream of the CURRENT device: every entry point first makes the device of its tensors current
    (``on_device``), because the library launches on the current device."""
    if _RAW_STREAM is not None and torch.cuda.is_initialized():
        # the handle alone, without building a torch.cuda.Stream object per launch (10 us of a 0.3 ms frame)
        return C.c_void_p(_RAW_STREAM(torch.cuda.current_device()))
    if not torch.cuda.is_available():
        raise _abi.SdError("no CUDA device available (scenedino_b200 has no CPU path)")
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _devices(objs, found):
    for o in objs:
        if torch.is_tensor(o):
            if o.is_cuda:
                found.add(o.device)
        elif isinstance(o, (list, tuple)):
            _devices(o, found)
        elif hasattr(o, "sd_tensors"):          # ops.Scene / ops.Mlp
            _devices(o.sd_tensors(), found)


class on_device:
    """Context manager: makes the device the given CUDA tensors live on the current one for the C-ABI calls inside.
    The library launches its kernels on the current device and that device's current stream, and keeps per-device
    kernel attributes; a model on cuda:1 while cuda:0 is current would otherwise run device-1 pointers on device 0.
    All CUDA tensors passed must share one device."""

    def __init__(self, *objs):
        found = set()
        _devices(objs, found)
        if len(found) > 1:
            raise _abi.SdError("all tensors of one call must live on ONE CUDA device, got " + ", ".join(sorted(map(str, found))))
        self._ctx = torch.cuda.device(next(iter(found))) if found else contextlib.nullcontext()

    def __enter__(self):
        return self._ctx.__enter__()

    def __exit__(self, *exc):
        return self._ctx.__exit__(*exc)


def device_guard(fn):
    """Decorator form of :class:`on_device` over every argument of ``fn``."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        with on_device(*args, *kwargs.values()):
            return fn(*args, **kwargs)
    return wrapper


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _abi.SdError(f"{what} must live on a CUDA device (scenedino_b200 has no CPU path); got {t.device}")


class PackedMlp:
    """Device blob written by sd_mlp_pack for one two-layer head, re-packed when weights change."""

    def __init__(self):
        self.blob = None
        self.key = None
        self.dims = None
        self.generation = 0          # bumped on every (re)pack: caches derived from the blob key on it

    def get(self, lin_in: nn.Linear, lin_out: nn.Linear, precision: int) -> _abi.SdMlp:
        params = (lin_in.weight, lin_in.bias, lin_out.weight, lin_out.bias)
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in params)
        d_hidden, d_in = lin_in.weight.shape
        d_out = lin_out.weight.shape[0]
        if key != self.key:
            require_cuda(lin_in.weight, "MLP weights")
            with on_device(*params):
                self._pack(params, d_in, d_hidden, d_out, key)
        m = _abi.SdMlp()
        m.packed = self.blob.data_ptr()
        m.d_in, m.d_hidden, m.d_out = self.dims
        m.precision = precision
        return m

    def _pack(self, params, d_in, d_hidden, d_out, key):
        lib = _abi.lib()
        nbytes = lib.sd_mlp_pack_bytes(d_in, d_hidden, d_out)
        blob = torch.empty(nbytes + 1024, dtype=torch.uint8, device=params[0].device)
        off = (-blob.data_ptr()) % 1024
        blob = blob[off:off + nbytes]
        ws = [_f32c(p) for p in params]
        _abi.check(lib.sd_mlp_pack(_ptr(ws[0]), _ptr(ws[1]), _ptr(ws[2]), _ptr(ws[3]), d_in, d_hidden,
                                   d_out, _ptr(blob), _stream()), "sd_mlp_pack")
        self.blob, self.key, self.dims = blob, key, (d_in, d_hidden, d_out)
        self.generation += 1


class ResnetFC(nn.Module):
    """resnetfc.py:66-203 with n_blocks = 0: lin_in -> ReLU -> lin_out."""

    def __init__(self, d_in, view_number=None, d_out=4, n_blocks=0, d_latent=0, d_hidden=128, beta=0.0,
                 combine_layer=1000, combine_type="average", use_spade=False):
        super().__init__()
        if n_blocks != 0 or d_latent != 0 or beta > 0 or use_spade:
            raise NotImplementedError(
                "scenedino_b200.ResnetFC implements the shipped configuration only "
                "(n_blocks=0, d_latent=0, ReLU; configs/model/dino_downsampler.yaml:35-41)")
        self.lin_in = nn.Linear(d_in, d_hidden)
        nn.init.constant_(self.lin_in.bias, 0.0)
        nn.init.kaiming_normal_(self.lin_in.weight, a=0, mode="fan_in")
        self.lin_out = nn.Linear(d_hidden, d_out)
        nn.init.constant_(self.lin_out.bias, 0.0)
        nn.init.kaiming_normal_(self.lin_out.weight, a=0, mode="fan_in")
        self.n_blocks, self.d_latent, self.d_in, self.view_number = n_blocks, d_latent, d_in, view_number
        self.d_out, self.d_hidden = d_out, d_hidden
        self.combine_layer, self.combine_type, self.use_spade = combine_layer, combine_type, use_spade
        self.blocks = nn.ModuleList([])
        self.activation = nn.ReLU()
        self._packed = PackedMlp()

    def packed(self, precision: int) -> _abi.SdMlp:
        return self._packed.get(self.lin_in, self.lin_out, precision)

    def forward(self, zx, combine_inner_dims=(1,), combine_index=None, dim_size=None, precision=_abi.SD_MLP_FP32,
                **kwargs):
        """zx [..., d_in] -> [..., d_out] (resnetfc.py:135-203; the view-combine arguments only
        matter for n_blocks > 0 and are accepted for signature compatibility)."""
        assert zx.size(-1) == self.d_latent + self.d_in  # resnetfc.py:155
        require_cuda(zx, "ResnetFC input")
        x = _f32c(zx).reshape(-1, self.d_in)
        out = torch.empty(x.shape[0], self.d_out, dtype=torch.float32, device=x.device)
        with on_device(x, self.lin_in.weight):
            mlp = self.packed(precision)
            _abi.check(_abi.lib().sd_mlp_forward(C.byref(mlp), _ptr(x), x.shape[0], _ptr(out), _stream()),
                       "sd_mlp_forward")
        return out.reshape(*zx.shape[:-1], self.d_out)

    @classmethod
    def from_conf(cls, conf, d_in, d_out, d_latent=0):
        return cls(d_in=d_in, d_out=d_out, **conf)


def make_head(conf, d_in: int, d_out: int):
    """models/prediction_heads/__init__.py:14-47 for the head type the shipped configs select."""
    head_type = conf.get("type", "resnet")
    if head_type != "resnet":
        raise NotImplementedError("Unsupported Head type")
    head = ResnetFC.from_conf(dict(conf["args"]), d_in, d_out)
    if conf.get("freeze", False):
        for p in head.parameters():
            p.requires_grad = False
    return head


class PositionalEncoding(nn.Module):
    """common/positional_encoding.py:44-66: carries (num_freqs, freq_factor, include_input); the code
    is evaluated inside the field kernels (sd_sample_features / sd_query_points); ``forward`` is the stand-alone form."""

    def __init__(self, num_freqs=6, d_in=3, freq_factor=math.pi, include_input=True):
        super().__init__()
        self.num_freqs, self.d_in, self.freq_factor, self.include_input = num_freqs, d_in, float(freq_factor), include_input
        self.freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
        self.d_out = self.num_freqs * 2 * d_in + (d_in if include_input else 0)
        self.register_buffer("_freqs", torch.repeat_interleave(self.freqs, 2).view(1, -1, 1))
        ph = torch.zeros(2 * self.num_freqs)
        ph[1::2] = math.pi * 0.5
        self.register_buffer("_phases", ph.view(1, -1, 1))

    def forward(self, x):
        """positional_encoding.py:68-80: x (batch, d_in) -> (batch, d_out) (sd_positional_encoding; inside the field
        kernels the same code is fused and never materialised)."""
        with torch.profiler.record_function("positional_enc"):
            require_cuda(x, "PositionalEncoding input")
            if x.dim() != 2 or x.shape[1] != self.d_in:
                raise ValueError(f"PositionalEncoding.forward expects (batch, {self.d_in}), got {tuple(x.shape)}")
            xs = _f32c(x)
            out = torch.empty((xs.shape[0], self.d_out), dtype=torch.float32, device=xs.device)
            with on_device(xs):
                _abi.check(_abi.lib().sd_positional_encoding(_ptr(xs), xs.shape[0], self.d_in, self.num_freqs, self.freq_factor,
                                                             int(bool(self.include_input)), _ptr(out), _stream()),
                           "sd_positional_encoding")
            return out.to(x.dtype)

    @classmethod
    def from_conf(cls, conf, d_in=3):
        return cls(conf.get("num_freqs", 6), d_in, conf.get("freq_factor", math.pi), conf.get("include_input", True))


_expand_precision = threading.local()


def note_field_precision(precision: int) -> None:
    """Remembers the precision of this thread's latest field query / render: an ``MlpDimReduction`` in "auto" mode
    follows it when it is called outside BTSNet.forward -- the demo expands the rendered features right after the
    render (demo_utils/utils.py:223-229) and should not drop from the tensor cores to the fp32 CUDA-core path there."""
    _expand_precision.last = precision


@contextlib.contextmanager
def expand_precision(precision: int):
    """Precision of MlpDimReduction.transform_expand calls made inside the block (BTSNet.forward wraps
    ``encoder.expand_dim`` in it so that the expansion follows the precision of the query that feeds it)."""
    prev = getattr(_expand_precision, "value", None)
    _expand_precision.value = precision
    try:
        yield
    finally:
        _expand_precision.value = prev


class MlpDimReduction(nn.Module):
    """dim_reduction.py:15-25; transform_expand = 64 -> 128 -> ReLU -> 768 -> L2 normalise.

    ``precision``: "fp32" (CUDA cores, rel 1e-4), "fp16" (tensor cores, rel 2e-2; 64 -> 128 -> multiple-of-128 shapes)
    or "auto" (default): what the enclosing BTSNet.forward runs in, else fp16 under torch autocast, else the precision of
    this thread's latest field query / render (the features being expanded came from it), else fp32."""

    def __init__(self, full_channels, reduced_channels, latent_channels):
        super().__init__()
        self.linear_in = nn.Linear(reduced_channels, latent_channels)
        self.linear_out = nn.Linear(latent_channels, full_channels)
        self.relu = nn.ReLU()
        self.precision = "auto"
        self._packed = PackedMlp()

    def _precision(self) -> int:
        if self.precision == "auto":
            ctx = getattr(_expand_precision, "value", None)
            if ctx is not None:
                return ctx
            if torch.is_autocast_enabled():
                return _abi.SD_MLP_F16_TC
            last = getattr(_expand_precision, "last", None)
            return last if last is not None else _abi.SD_MLP_FP32
        if self.precision not in ("fp32", "fp16"):
            raise ValueError(f"MlpDimReduction.precision must be 'auto', 'fp32' or 'fp16', got {self.precision!r}")
        return _abi.SD_MLP_F16_TC if self.precision == "fp16" else _abi.SD_MLP_FP32

    def transform_expand(self, features):
        require_cuda(features, "transform_expand input")
        d_red = self.linear_in.weight.shape[1]
        d_full = self.linear_out.weight.shape[0]
        x = _f32c(features).reshape(-1, d_red)
        out = torch.empty(x.shape[0], d_full, dtype=torch.float32, device=x.device)
        with on_device(x, self.linear_in.weight):
            mlp = self._packed.get(self.linear_in, self.linear_out, self._precision())
            _abi.check(_abi.lib().sd_expand_dim(C.byref(mlp), _ptr(x), x.shape[0], _ptr(out), _stream()), "sd_expand_dim")
        out = out.reshape(*features.shape[:-1], d_full)
        # the SSC head that follows (SemanticHead.forward, semantic_head.py) is evaluated fused with this expansion, starting
        # from the 64-d features: the result carries them (a plain attribute on the tensor object the caller passes on)
        out._sd_reduced = (features, self)
        return out
