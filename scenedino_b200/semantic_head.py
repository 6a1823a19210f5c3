"""``SemanticHead`` -- host-side mirror of the inference half of the reference's unsupervised SSC head
(scenedino/downstream_head/semantic_head.py:41-120, 285-373).

Same constructor / ``from_conf`` keys, sub-module and parameter names (``stego_head.linear_path.0.weight`` ...,
``stego_cluster_head.cluster_centers`` / ``pseudo_assignment``: the state-dict keys of reference checkpoints) and the same
``forward(features, mode="stego_kmeans") -> int64 labels`` convention.  The arithmetic runs in libscenedino_b200
(``sd_ssc_head``, csrc/ssc_head.cu): the head is fused with the 64 -> 128 -> 768 expansion that feeds it, and everything
linear between the two ReLUs is folded once per model, so the 768-d rows (6.4 GB per SSC grid) are never formed.

The fusion starts from the 64-d features.  ``MlpDimReduction.transform_expand`` (heads.py) tags the 768-d tensor it
returns with the 64-d tensor and the module it came from; when ``forward`` is handed such a tensor -- which is what the
reference's callers do (``models/bts.py:585-588``, ``demo_utils/utils.py:229-232``) -- it runs the fused kernel on the tag.
``forward_reduced`` is the direct form.  Only the modes the SSC evaluation uses are implemented (``stego_kmeans``); the
training half (``forward_training``: STEGO correlation losses, kNN buffers, CRF) stays with the reference.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _abi
from .heads import MlpDimReduction, _f32c, _ptr, _stream, on_device, require_cuda


class StegoClusterHead(nn.Module):
    """semantic_head.py:285-305: parameter container (1x1 convolutions = per-feature linear maps)."""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid_channels = in_channels if mid_channels is None else mid_channels
        self.linear_path = nn.Sequential(nn.Conv2d(in_channels, out_channels, (1, 1)), nn.Dropout2d(p=.1))
        self.nonlinear_path = nn.Sequential(nn.Conv2d(in_channels, mid_channels, (1, 1)), nn.ReLU(),
                                            nn.Conv2d(mid_channels, out_channels, (1, 1)), nn.Dropout2d(p=.1))

    def forward(self, x):
        raise NotImplementedError("StegoClusterHead is evaluated fused inside SemanticHead.forward (sd_ssc_head)")


class KMeansParamHead(nn.Module):
    """semantic_head.py:308-373: cluster centres + pseudo-label LUT (parameter container)."""

    def __init__(self, n_classes: int, gt_classes: int, dim: int):
        super().__init__()
        self.n_classes, self.dim = n_classes, dim
        self.init_type = "random"
        self.cluster_centers = nn.Parameter(torch.randn(n_classes, dim))
        self.centroids_initialized = False
        self.register_buffer("pseudo_assignment", torch.arange(0, n_classes).remainder(gt_classes))

    def forward(self, features, weight=None):
        raise NotImplementedError("KMeansParamHead is evaluated fused inside SemanticHead.forward (sd_ssc_head)")


class _Linear(nn.Module):
    def __init__(self, dim, gt_classes):
        super().__init__()
        self.linear = nn.Linear(dim, gt_classes)


class _Mlp(nn.Module):
    def __init__(self, dim, gt_classes):
        super().__init__()
        self.linear1 = nn.Linear(dim, 2 * dim)
        self.linear2 = nn.Linear(2 * dim, gt_classes)
        self.activation = nn.ReLU()


class SemanticHead(nn.Module):
    def __init__(self, n_classes, gt_classes, input_dim, code_dim, buffer_size=0, patch_sample_size=0, knn_neighbors=0,
                 mode="2d", mlp_head=False, apply_crf=False):
        super().__init__()
        self.n_classes, self.gt_classes = n_classes, gt_classes
        self.input_dim, self.code_dim = input_dim, code_dim
        self.knn_neighbors, self.mode, self.apply_crf = knn_neighbors, mode, apply_crf
        self.buffer_size = buffer_size
        # (the reference also allocates kNN buffers for the STEGO training losses on "cuda" here; training is not mirrored)
        self.direct_cluster_head = KMeansParamHead(n_classes, gt_classes, input_dim)
        self.stego_head = StegoClusterHead(input_dim, code_dim)
        self.stego_cluster_head = KMeansParamHead(n_classes, gt_classes, code_dim)
        head = _Mlp if mlp_head else _Linear
        self.direct_linear_head = head(input_dim, gt_classes)
        self.stego_linear_head = head(code_dim, gt_classes)
        self._blob = None
        self._key = None

    @classmethod
    def from_conf(cls, config):
        g = config.get
        return cls(n_classes=config["n_classes"], gt_classes=config["gt_classes"], input_dim=config["input_dim"],
                   code_dim=config["code_dim"], buffer_size=g("buffer_size", 0), patch_sample_size=g("patch_sample_size", 0),
                   knn_neighbors=g("knn_neighbors", 0), mode=g("mode", "2d"), mlp_head=g("mlp_head", False),
                   apply_crf=g("apply_crf", False))

    # ---- packing ---------------------------------------------------------------------------------------------------
    def _params(self, expand: MlpDimReduction):
        s, k = self.stego_head, self.stego_cluster_head
        return (expand.linear_in.weight, expand.linear_in.bias, expand.linear_out.weight, expand.linear_out.bias,
                s.linear_path[0].weight, s.linear_path[0].bias, s.nonlinear_path[0].weight, s.nonlinear_path[0].bias,
                s.nonlinear_path[2].weight, s.nonlinear_path[2].bias, k.cluster_centers, k.pseudo_assignment)

    def packed(self, expand: MlpDimReduction):
        """Blob of sd_ssc_head_pack for this head behind ``expand``; re-packed when any parameter changes."""
        params = self._params(expand)
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in params) + (id(expand),)
        if key != self._key:
            require_cuda(params[0], "SemanticHead / expansion weights")
            d_lat, d_red = params[0].shape
            d_full = params[2].shape[0]
            d_code, d_mid = params[8].shape[0], params[8].shape[1]
            n_cls = params[10].shape[0]
            lib = _abi.lib()
            nbytes = lib.sd_ssc_head_pack_bytes(d_red, d_lat, d_full, d_mid, d_code, n_cls)
            if nbytes == 0:
                raise NotImplementedError(
                    f"sd_ssc_head supports 64 -> 128 -> d_full expansions, d_mid a multiple of 128 (<= 1024), code_dim 64, "
                    f"<= 32 clusters; got {d_red} -> {d_lat} -> {d_full}, mid {d_mid}, code {d_code}, {n_cls} clusters")
            with on_device(*params):
                raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=params[0].device)
                off = (-raw.data_ptr()) % 1024
                blob = raw[off:off + nbytes]
                ws = [_f32c(p).reshape(p.shape[0], -1) if p.dim() > 1 else _f32c(p) for p in params[:11]]
                lut = params[11].detach().to(torch.int64).contiguous()
                _abi.check(lib.sd_ssc_head_pack(*[_ptr(w) for w in ws], _ptr(lut), d_red, d_lat, d_full, d_mid, d_code, n_cls,
                                                _ptr(blob), _stream()), "sd_ssc_head_pack")
            self._blob, self._key, self._dims = blob, key, (d_mid, n_cls)
        return self._blob, self._dims

    # ---- inference ---------------------------------------------------------------------------------------------------
    def forward_reduced(self, reduced, expand: MlpDimReduction, mode="stego_kmeans", want_scores=False, perm=None,
                        out=None):
        """reduced [..., 64] (the field's / renderer's 64-d features) -> int64 labels [...] of
        ``SemanticHead.forward(expand.transform_expand(reduced), mode)`` (semantic_head.py:107-112), fused.
        ``want_scores``: also the pseudo labels and cosine scores, as a dict.  ``out``: optional uint8 label buffer."""
        if mode != "stego_kmeans":
            raise NotImplementedError(f"mode {mode!r}: only 'stego_kmeans' is implemented on the device")
        require_cuda(reduced, "SemanticHead input")
        x = _f32c(reduced).reshape(-1, reduced.shape[-1])
        N = x.shape[0]
        with on_device(x, expand.linear_in.weight, self.stego_cluster_head.cluster_centers):
            blob, (d_mid, n_cls) = self.packed(expand)
            seg = out if out is not None else torch.empty((N,), dtype=torch.uint8, device=x.device)
            pseudo = torch.empty((N,), dtype=torch.uint8, device=x.device) if want_scores else None
            scores = torch.empty((N, n_cls), dtype=torch.float32, device=x.device) if want_scores else None
            _abi.check(_abi.lib().sd_ssc_head(_ptr(blob), d_mid, n_cls, _ptr(x), _ptr(perm), N, _ptr(seg), _ptr(pseudo),
                                              _ptr(scores), _stream()), "sd_ssc_head")
        if out is not None:
            return out
        labels = seg.to(torch.int64).reshape(reduced.shape[:-1])
        if want_scores:
            return {"segs_pred": labels, "pseudo_segs_pred": pseudo.to(torch.int64).reshape(reduced.shape[:-1]),
                    "inner_products": scores.reshape(*reduced.shape[:-1], n_cls)}
        return labels

    def forward(self, features, mode="stego_kmeans"):
        """semantic_head.py:107-120.  ``features`` must be the tensor ``MlpDimReduction.transform_expand`` returned (it
        carries the 64-d features it was expanded from): the head is evaluated fused with that expansion."""
        src = getattr(features, "_sd_reduced", None)
        if src is None:
            raise NotImplementedError(
                "SemanticHead.forward needs the 768-d tensor returned by MlpDimReduction.transform_expand (it carries the "
                "64-d features the fused kernel starts from); for other inputs call forward_reduced(reduced, expand)")
        reduced, expand = src
        return self.forward_reduced(reduced, expand, mode)

    def update_model_eval(self, metrics):
        self.direct_cluster_head.pseudo_assignment[:] = metrics["direct_cluster_assignment"]
        self.stego_cluster_head.pseudo_assignment[:] = metrics["stego_cluster_assignment"]


def make_downstream_head(conf):
    """downstream_head/__init__.py:4-9."""
    head_type = conf.get("type", None)
    if head_type == "segmentation":
        return SemanticHead.from_conf(conf)
    raise NotImplementedError(f"Downstream head type '{head_type}' is not implemented.")
