"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, compute entry points fail loudly without a GPU (no fallback), and the reference-shaped
Python surface keeps the reference's names, keys and shapes."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import scenedino_b200 as sd
from scenedino_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "scenedino_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(_abi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} is declared in include/scenedino_b200.h but not exported"
        assert n in _abi.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_abi.PROTOTYPES) == names
    assert _abi.lib().sd_abi_version() == _abi.ABI_VERSION == 5


def test_library_exports_nothing_the_header_does_not_declare():
    """the other direction: every dynamic `sd_*` symbol of the built library (diagnostics included) is in the header"""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("sd_")})
    assert exported == header_symbols()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_error_not_fallback():
    lib = _abi.lib()
    assert lib.sd_device_sm_count() == -2          # SD_ERR_CUDA
    assert "CUDA" in _abi.last_error()
    z = np.zeros((4, 8), np.float32)
    rc = lib.sd_sort_rows(z.ctypes.data_as(ctypes.c_void_p), 4, 8, None)
    assert rc == -2
    with pytest.raises(sd.SdError):
        from scenedino_b200 import ops
        ops.sort_rows(torch.zeros(4, 8))


def test_argument_validation_without_gpu():
    lib = _abi.lib()
    assert lib.sd_sort_rows(None, 4, 8, None) == -1 and "null" in _abi.last_error()
    assert lib.sd_sort_rows(None, 0, 8, None) == 0          # empty inputs are fine
    assert lib.sd_mlp_pack_bytes(295, 128, 65) > 4 * (295 * 128 + 128 * 65)
    assert lib.sd_mlp_pack_bytes(0, 128, 65) == 0
    assert lib.sd_render_workspace_bytes(None, None, 10, 10) == 0
    # projected scene: sizes are host arithmetic (header of operand images + 256 B per texel), errors are loud
    sc = _abi.SdScene()
    sc.Hf, sc.Wf, sc.C, sc.nv_f = 384, 1280, 256, 1
    assert lib.sd_field_project_bytes(ctypes.byref(sc)) == 50176 + 384 * 1280 * 256
    assert lib.sd_field_project_bytes(None) == 0
    ml = _abi.SdMlp()
    assert lib.sd_field_project(ctypes.byref(sc), ctypes.byref(ml), None, 0, None) == -1 and "null" in _abi.last_error()
    buf = ctypes.create_string_buffer(2048)
    assert lib.sd_field_project(ctypes.byref(sc), ctypes.byref(ml), buf, 2048, None) == -1   # no fp16 map in the scene
    assert lib.sd_profile_next_kernel(None, None) == 0
    # the sort's scratch grows with the points: 4 + 2 + 2 B of indices, a 32-byte record, a 16-byte entry per 128-point tile
    sc.feat_dtype = _abi.SD_F16
    ml.d_in, ml.d_hidden, ml.d_out, ml.precision = 295, 128, 65, _abi.SD_MLP_F16_TC
    n = 1 << 21
    need = lib.sd_query_workspace_bytes(ctypes.byref(sc), ctypes.byref(ml), n)
    # ... plus what does not: per range of the two passes (at most 320) a 4-byte offset and a 2-byte list entry per texel bin
    nbins = ((384 - 1) // 7 + 1) * ((1280 - 1) // 7 + 1)
    assert n * (4 + 2 + 2 + 32) + (n // 128) * 16 + 320 * nbins * 6 <= need <= n * 41 + (1 << 20) + 320 * nbins * 6
    assert lib.sd_query_workspace_bytes(ctypes.byref(sc), ctypes.byref(ml), 1000) == 0        # small queries are not sorted
    ml.precision = _abi.SD_MLP_FP32
    assert lib.sd_query_workspace_bytes(ctypes.byref(sc), ctypes.byref(ml), n) == 0           # nor is the fp32 path


def _net(conf_extra=None):
    class Enc(torch.nn.Module):
        latent_size, extra_outs = 256, 0

        def forward(self, x, ground_truth=False):
            return [torch.zeros(x.shape[0], 256, 4, 4)]

    conf = {"predict_dino": True, "dino_dims": 64, "learn_empty": False, "code_mode": "z"}
    conf.update(conf_extra or {})
    code = sd.PositionalEncoding.from_conf({"num_freqs": 6, "freq_factor": 1.5, "include_input": True})
    head = sd.make_head({"type": "resnet", "name": "normal_head", "args": {"n_blocks": 0, "d_hidden": 128}}, 256 + code.d_out, 65)
    return sd.BTSNet(conf, Enc(), code, {"normal_head": head}, None)


def test_state_dict_keys_match_reference_names():
    """Checkpoint keys: renderer.net.heads.<name>.lin_in/lin_out.{weight,bias}, renderer.renderer
    buffers iter_idx / last_sched."""
    net = _net({"learn_empty": True})
    ren = sd.NeRFRenderer.from_conf({"n_coarse": 32})
    wrapped = ren.bind_parallel(net)

    class Wrapper(torch.nn.Module):     # like BTSWrapper: attribute `renderer`
        def __init__(self):
            super().__init__()
            self.renderer = wrapped

    keys = set(Wrapper().state_dict().keys())
    for k in ("renderer.net.heads.normal_head.lin_in.weight", "renderer.net.heads.normal_head.lin_in.bias",
              "renderer.net.heads.normal_head.lin_out.weight", "renderer.net.heads.normal_head.lin_out.bias",
              "renderer.net.empty_feature", "renderer.renderer.iter_idx", "renderer.renderer.last_sched"):
        assert k in keys, k
    assert net.heads["normal_head"].lin_in.weight.shape == (128, 295)
    assert net._d_in == 295 and net._d_out == 65
    assert float(net.heads["normal_head"].lin_in.bias.detach().abs().sum()) == 0.0     # resnetfc.py:91


def test_renderer_conf_and_schedule():
    ren = sd.NeRFRenderer.from_conf({})
    assert (ren.n_coarse, ren.n_fine, ren.lindisp, ren.hard_alpha_cap, ren.render_mode) == (128, 0, True, False, "volumetric")
    assert sd.NeRFRenderer().lindisp is False               # __init__ default differs from from_conf (nerf.py:82,631)
    ren = sd.NeRFRenderer(n_coarse=8, sched=[[2, 4], [16, 32], [0, 8]])
    ren.sched_step(); assert ren.n_coarse == 8
    ren.sched_step(); assert (ren.n_coarse, ren.n_fine, int(ren.last_sched)) == (16, 0, 1)
    ren.sched_step(2); assert (ren.n_coarse, ren.n_fine, int(ren.last_sched)) == (32, 8, 2)
    with pytest.raises(NotImplementedError):
        ren.bind_parallel(None, gpus=[0, 1])


def test_format_outputs_shapes_and_quirk():
    """_format_outputs (nerf.py:541-598): keys, super-batch reshape, and invalid_features reshaped with
    invalid's last dimension (nv_c) as the reference does."""
    ren = sd.NeRFRenderer(n_coarse=4)
    B, K, nv = 8, 4, 2
    tup = (torch.rand(B, K), torch.rand(B, 3 * nv), torch.rand(B), torch.rand(B, K), torch.zeros(B, K, nv), torch.rand(B, K),
           torch.rand(B, K, 3 * nv), torch.rand(B, 1, 3), None,
           {"dino_features": torch.rand(B, 64), "invalid_features": torch.zeros(B, K, 1, dtype=torch.bool)})
    o = ren._format_outputs(tup, 2, want_weights=True, want_alphas=True, want_z_samps=True, want_rgb_samps=True)
    assert sorted(o.keys()) == sorted(["rgb", "depth", "invalid", "ray_info", "weights", "alphas", "z_samps", "rgb_samps",
                                       "dino_features", "invalid_features"])
    assert o.rgb.shape == (2, 4, 6) and o.depth.shape == (2, 4) and o.invalid.shape == (2, 4, K, nv)
    assert o.ray_info.shape == (2, 4, 3) and o.rgb_samps.shape == (2, 4, K, 6) and o.dino_features.shape == (2, 4, 64)
    assert o.invalid_features.shape == (2, 4 // nv, K, nv)      # the reference's reshape with out_d_i = nv_c
    o = ren._format_outputs(tup, 2)
    assert sorted(o.keys()) == sorted(["rgb", "depth", "invalid", "ray_info", "dino_features", "invalid_features"])
    d = sd.DotMap(a=sd.DotMap(b=1), c=2)
    assert d.a.b == 1 and d.toDict() == {"a": {"b": 1}, "c": 2}


def test_gen_rays_argument_validation_without_gpu():
    lib = _abi.lib()
    buf = ctypes.create_string_buffer(64)
    assert lib.sd_gen_rays(None, None, None, 0, 4, 4, 3.0, 80.0, 1, 0.0, 0.0, None, None) == 0       # no views: nothing to do
    assert lib.sd_gen_rays(None, None, None, 1, 4, 4, 3.0, 80.0, 1, 0.0, 0.0, None, None) == -1 and "null" in _abi.last_error()
    assert lib.sd_gen_rays(buf, buf, None, 1, 1, 4, 3.0, 80.0, 1, 0.0, 0.0, buf, None) == -1 and "2 x 2" in _abi.last_error()
    assert lib.sd_gen_rays(buf, buf, None, -1, 4, 4, 3.0, 80.0, 1, 0.0, 0.0, buf, None) == -1
    with pytest.raises(sd.SdError):          # host tensors are rejected: there is no CPU path
        sd.ImageRaySampler(3.0, 80.0, 4, 4).sample(None, torch.eye(4).view(1, 1, 4, 4), torch.eye(3).view(1, 1, 3, 3))


def test_image_ray_sampler_reconstruct_shapes():
    """ImageRaySampler.reconstruct (ray_sampler.py:515-607): every per-ray entry becomes [n, v_in, H, W, ...] as a view
    (checked against the reference's own reconstruct in the build container), ground truth included; the patch size
    of the DINO ground truth is inferred from the element count as the reference does."""
    n, v, H, W, K, vr, c, dd = 2, 2, 8, 12, 5, 3, 3, 16
    R = v * H * W

    def part():
        return dict(rgb=torch.rand(n, R, vr * c), weights=torch.rand(n, R, K), depth=torch.rand(n, R), invalid=torch.rand(n, R, K, vr),
                    invalid_features=torch.rand(n, R, K, vr) > 0.5, alphas=torch.rand(n, R, K), z_samps=torch.rand(n, R, K),
                    rgb_samps=torch.rand(n, R, K, vr * c), ray_info=torch.rand(n, R, 3), extras=torch.rand(n, R, 4),
                    dino_features=torch.rand(n, R, dd))

    d = dict(coarse=part(), fine=part(), rgb_gt=torch.rand(n, R, c), dino_gt=torch.rand(n, v * (H // 4) * (W // 4), dd),
             dino_artifacts=torch.rand(n, v * (H // 4) * (W // 4), dd), other=3)
    flat = {k: t.clone() for k, t in d["coarse"].items()}
    o = sd.ImageRaySampler(3.0, 80.0, H, W).reconstruct(d)
    want = {"rgb": (vr, c), "weights": (K,), "depth": (), "invalid": (K, vr), "invalid_features": (K, vr), "alphas": (K,),
            "z_samps": (K,), "rgb_samps": (K, vr, c), "ray_info": (3,), "extras": (4,), "dino_features": (1, dd)}
    for lvl in ("coarse", "fine"):
        for k, tail in want.items():
            assert o[lvl][k].shape == (n, v, H, W) + tail, (lvl, k)
    for k, t in flat.items():                       # same memory order: a view, not a permutation
        assert torch.equal(o["coarse"][k].reshape(t.shape), t)
    assert o["rgb_gt"].shape == (n, v, H, W, c) and o["other"] == 3
    assert o["dino_gt"].shape == o["dino_artifacts"].shape == (n, v, H // 4, W // 4, dd)
    o = sd.ImageRaySampler(3.0, 80.0, H, W, dino_upscaled=True).reconstruct(dict(coarse=part(), dino_gt=torch.rand(n, R, dd)))
    assert o["dino_gt"].shape == (n, v, H, W, dd)
    with pytest.raises(KeyError):                   # the reference needs weights / depth / invalid next to rgb as well
        sd.ImageRaySampler(3.0, 80.0, H, W).reconstruct(dict(coarse=dict(rgb=torch.rand(n, R, 3))))


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        _net({"code_mode": "distance"})
    with pytest.raises(NotImplementedError):
        _net({"use_viewdirs": True})
    with pytest.raises(NotImplementedError):
        sd.ResnetFC(295, d_out=65, n_blocks=5)
    net = _net()
    with pytest.raises(RuntimeError, match="encode"):
        net._state(0)
    with pytest.raises(sd.SdError):
        net.eval()(torch.zeros(1, 4, 3))      # CPU points


def test_header_constants_match_the_binding():
    """ABI version, precision / dtype enums and the sd_scene layout the ctypes binding assumes, read from the header."""
    src = open(os.path.join(ROOT, "include", "scenedino_b200.h")).read()
    assert int(re.search(r"#define SD_ABI_VERSION (\d+)", src).group(1)) == _abi.ABI_VERSION
    m = re.search(r"typedef enum sd_precision \{([^}]*)\}", src).group(1)
    enum = {k.strip(): int(v) for k, v in (kv.split("=") for kv in m.split(","))}
    assert enum == {"SD_MLP_FP32": _abi.SD_MLP_FP32, "SD_MLP_F16_TC": _abi.SD_MLP_F16_TC, "SD_MLP_F32_TC": _abi.SD_MLP_F32_TC}
    body = re.search(r"typedef struct sd_scene \{(.*?)\} sd_scene;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            fields += [f.strip().lstrip("*").strip() for f in decl.split(",")]
    fields = [re.split(r"[\s\*]+", f)[-1] for f in fields]
    assert fields == [n for n, _ in _abi.SdScene._fields_], fields


def test_x3_projection_sizes_and_errors_without_gpu():
    """sd_field_project_x3: host arithmetic of the blob size (two operand-image pairs + two fp16 maps) and loud argument errors."""
    lib = _abi.lib()
    sc = _abi.SdScene()
    sc.Hf, sc.Wf, sc.C, sc.nv_f = 384, 1280, 256, 1
    assert lib.sd_field_project_x3_bytes(ctypes.byref(sc)) == 32768 + 2 * 2 * 80 * 128 + 2 * 384 * 1280 * 256
    assert lib.sd_field_project_x3_bytes(None) == 0
    ml = _abi.SdMlp()
    assert lib.sd_field_project_x3(ctypes.byref(sc), ctypes.byref(ml), None, 0, None) == -1
    assert lib.sd_query_points_binned(ctypes.byref(sc), ctypes.byref(ml), None, 10, None, None, None, None, None, 0, 0, None) == -1
    org = (ctypes.c_float * 3)(0, 0, 0)
    T = (ctypes.c_double * 12)(*([0.0] * 12))
    assert lib.sd_gen_voxel_grid(org, 0.2, 4, 4, 4, 3, 2, T, None, None) == -1 and "slab" in _abi.last_error()
    assert lib.sd_gen_voxel_grid(org, 0.2, 4, 4, 4, 2, 2, T, None, None) == 0      # an empty slab is fine
    assert lib.sd_composite_bwd(None, None, None, None, 0, 8, 4, 3, ctypes.byref(_abi.SdRenderCfg()), *([None] * 8), None) == 0
    assert lib.sd_composite_bwd(None, None, None, None, 5, 300, 4, 3, ctypes.byref(_abi.SdRenderCfg()), *([None] * 8), None) == -1


def test_encode_stash_on_the_host_matches_the_reference_indexing():
    """BTSNet.encode (bts.py:112-259) with host tensors: the stash (views picked by ids_encoder / ids_render, inverted
    poses, colour images rescaled to [0, 1]) is what the reference keeps; the device-side shortcuts of encode (index
    tensors kept on the GPU, the pose inverse replayed from a CUDA graph) leave host tensors to the plain torch calls."""
    net = _net()
    rs = np.random.RandomState(0)
    n, nv = 2, 3
    images = torch.from_numpy(rs.uniform(-1, 1, (n, nv, 3, 8, 8)).astype(np.float32))
    Ks = torch.from_numpy(rs.uniform(0.5, 1.5, (n, nv, 3, 3)).astype(np.float32))
    poses = torch.eye(4).repeat(n, nv, 1, 1)
    poses[..., :3, 3] = torch.from_numpy(rs.uniform(-2, 2, (n, nv, 3)).astype(np.float32))
    net.encode(images, Ks, poses, ids_encoder=[0], ids_render=[2, 1])
    w2c = torch.inverse(poses)
    assert torch.equal(net.grid_f_poses_w2c, w2c[:, [0]]) and torch.equal(net.grid_c_poses_w2c, w2c[:, [2, 1]])
    assert torch.equal(net.grid_f_Ks, Ks[:, [0]]) and torch.equal(net.grid_c_Ks, Ks[:, [2, 1]])
    assert torch.equal(net.grid_c_imgs, (images * 0.5 + 0.5)[:, [2, 1]])
    assert net.grid_f_features[0].shape == (n, 1, 256, 4, 4)
    assert net._pose_inverse._by_shape == {} and net._ids_cache == {}      # nothing device-side was set up
    net.encode(images, Ks, poses)                                          # no ids: everything, no copies
    assert net.grid_c_Ks is Ks and torch.equal(net.grid_c_poses_w2c, w2c)
    with pytest.raises(NotImplementedError):
        net.encode(images, Ks, poses, combine_ids=[[0, 1]])
