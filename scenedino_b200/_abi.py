"""ctypes binding of libscenedino_b200.so (include/scenedino_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, the caller gets an
exception -- never a silent PyTorch/CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SD_B200_LIB") or os.path.join(HERE, "libscenedino_b200.so")   # (override: kernel experiments)

SD_F32, SD_F16 = 0, 1
SD_MLP_FP32, SD_MLP_F16_TC, SD_MLP_F32_TC = 0, 1, 2
ABI_VERSION = 5


class SdError(RuntimeError):
    """A libscenedino_b200 entry point returned a negative sd_status."""


class SdScene(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p), ("feat_dtype", C.c_int),
        ("nv_f", C.c_int), ("C", C.c_int), ("Hf", C.c_int), ("Wf", C.c_int),
        ("K_f", C.c_void_p), ("w2c_f", C.c_void_p),
        ("rgb", C.c_void_p), ("nv_c", C.c_int), ("Hc", C.c_int), ("Wc", C.c_int),
        ("K_c", C.c_void_p), ("w2c_c", C.c_void_p),
        ("d_min", C.c_float), ("d_max", C.c_float), ("inv_z", C.c_int),
        ("num_freqs", C.c_int), ("freq_factor", C.c_float), ("include_input", C.c_int),
        ("learn_empty", C.c_int), ("empty_feature", C.c_void_p),
        ("feat_proj", C.c_void_p),
        ("feat_proj_x3", C.c_void_p),
    ]


class SdMlp(C.Structure):
    _fields_ = [("packed", C.c_void_p), ("d_in", C.c_int), ("d_hidden", C.c_int), ("d_out", C.c_int),
                ("precision", C.c_int)]


class SdRenderCfg(C.Structure):
    _fields_ = [("lindisp", C.c_int), ("hard_alpha_cap", C.c_int), ("white_bkgd", C.c_int)]


class SdRenderOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("depth", "dino", "rgb", "weights", "alphas", "z_samps", "invalid", "invalid_feat",
                                          "rgb_samps")]


class SdSampling(C.Structure):
    _fields_ = [("n_coarse", C.c_int), ("n_fine", C.c_int), ("n_fine_depth", C.c_int), ("depth_std", C.c_float)]


_P, _LL, _I, _F, _D, _SZ = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_double, C.c_size_t
_SC, _ML, _RC = C.POINTER(SdScene), C.POINTER(SdMlp), C.POINTER(SdRenderCfg)

# name -> (restype, argtypes); must list EVERY symbol declared in include/scenedino_b200.h
PROTOTYPES = {
    "sd_abi_version": (_I, []),
    "sd_last_error": (C.c_char_p, []),
    "sd_device_sm_count": (_I, []),
    "sd_launch_count": (_LL, []),
    "sd_profile_next_kernel": (_I, [_P, _P]),
    "sd_featmap_pack": (_I, [_P, _I, _I, _I, _I, _P, _I, _P]),
    "sd_mlp_pack_bytes": (_SZ, [_I, _I, _I]),
    "sd_mlp_pack": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "sd_field_project_bytes": (_SZ, [_SC]),
    "sd_field_project": (_I, [_SC, _ML, _P, _SZ, _P]),
    "sd_field_project_x3_bytes": (_SZ, [_SC]),
    "sd_field_project_x3": (_I, [_SC, _ML, _P, _SZ, _P]),
    "sd_project_points": (_I, [_P, _P, _P, _LL, _P, _P, _P, _P]),
    "sd_sample_features": (_I, [_SC, _P, _LL, _P, _P, _P]),
    "sd_sample_colors": (_I, [_SC, _P, _LL, _P, _P, _P]),
    "sd_mlp_forward": (_I, [_ML, _P, _LL, _P, _P]),
    "sd_query_workspace_bytes": (_SZ, [_SC, _ML, _LL]),
    "sd_query_points": (_I, [_SC, _ML, _P, _LL, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "sd_query_points_sorted": (_I, [_SC, _ML, _P, _LL, _P, _P, _P, _P, _P, _SZ, _P]),
    "sd_query_points_binned": (_I, [_SC, _ML, _P, _LL, _P, _P, _P, _P, _P, _SZ, _I, _P]),
    "sd_sample_coarse": (_I, [_P, _LL, _I, _P, _P, _I, _I, _P, _P]),
    "sd_sample_fine": (_I, [_P, _LL, _I, _P, _I, _P, _P, _I, _I, _P, _P, _P]),
    "sd_sample_fine_depth": (_I, [_P, _LL, _I, _P, _P, _I, _F, _P, _P]),
    "sd_sample_coarse_from_dist": (_I, [_LL, _P, _P, _I, _P, _P, _I, _I, _P, _P, _P]),
    "sd_sort_rows": (_I, [_P, _LL, _I, _P]),
    "sd_composite": (_I, [_P, _P, _P, _P, _LL, _I, _I, _I, _RC, _P, _P, _P, _P, _P, _P]),
    "sd_composite_bwd": (_I, [_P, _P, _P, _P, _LL, _I, _I, _I, _RC, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "sd_sample_features_bwd": (_I, [_SC, _P, _LL, _P, _P, _P, _P]),
    "sd_render_workspace_bytes": (_SZ, [_SC, _ML, _LL, _I]),
    "sd_render_pass": (_I, [_SC, _ML, _RC, _P, _LL, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                            _P, _SZ, _P]),
    "sd_render_rays_workspace_bytes": (_SZ, [_SC, _ML, C.POINTER(SdSampling), _LL]),
    "sd_render_rays": (_I, [_SC, _ML, _RC, C.POINTER(SdSampling), _P, _LL, _I, _P, _P, _P, _P, _P, C.POINTER(SdRenderOut),
                            C.POINTER(SdRenderOut), _P, _SZ, _P]),
    "sd_expand_dim": (_I, [_ML, _P, _LL, _P, _P]),
    "sd_gen_rays": (_I, [_P, _P, _P, _I, _I, _I, _F, _F, _I, _F, _F, _P, _P]),
    "sd_gen_voxel_grid": (_I, [_P, _D, _I, _I, _I, _I, _I, _P, _P, _P]),
    "sd_ssc_head_pack_bytes": (_SZ, [_I, _I, _I, _I, _I, _I]),
    "sd_ssc_head_pack": (_I, [_P] * 12 + [_I] * 6 + [_P, _P]),
    "sd_ssc_head": (_I, [_P, _I, _I, _P, _P, _LL, _P, _P, _P, _P]),
    "sd_positional_encoding": (_I, [_P, _LL, _I, _I, _F, _I, _P, _P]),
    "sd_debug_read_trace": (_I, [_P]),
    "sd_debug_read_trace_bin": (_I, [_P]),
    "sd_debug_read_cta_ns": (_I, [_P]),
    "sd_debug_read_tiles_bin": (_I, [_P]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdError(
                f"{LIB_PATH} is missing: build it with `python -m scenedino_b200.build` "
                "(scenedino_b200 has no PyTorch/CPU fallback)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        v = h.sd_abi_version()
        if v != ABI_VERSION:
            raise SdError(f"libscenedino_b200 ABI version {v}, binding expects {ABI_VERSION}")
        _lib = h
    return _lib


def last_error() -> str:
    return lib().sd_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        kind = {-1: "invalid argument", -2: "CUDA error", -3: "workspace too small"}.get(rc, f"status {rc}")
        raise SdError(f"{what or 'libscenedino_b200'}: {kind}: {last_error()}")


def launch_count() -> int:
    return int(lib().sd_launch_count())
