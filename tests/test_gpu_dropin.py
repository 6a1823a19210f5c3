"""Drop-in proof (north_star: "drops in behind demo_script.py, eval.py and sscbench/evaluate_model_sscbench.py unchanged").

The reference's OWN caller functions -- ``demo_utils/utils.py: inference_3d, inference_rendered_2d`` and
``sscbench/evaluate_model_sscbench.py: downsample_and_predict / predict_grid`` -- run unmodified on top of the B200-native
classes after ``scenedino_b200.install()`` has rebound them inside the reference's modules; the model is built by the
reference's own factory ``scenedino.models.make_model`` from a config shaped like configs/model/dino_downsampler.yaml (only
the encoder factory is replaced by a seeded feature map: the ViT stays out of scope and there are no weights to load).
Results are checked against the CPU oracle.

The reference tree is test infrastructure here: ``baseline/_ref`` (copied by baseline/install_ref.py in the build
container, git-ignored, shipped to the GPU box) or /root/reference.  Needs a B200: ``-m gpu``.
"""
import numpy as np
import pytest
import torch

import scenedino_b200 as sd
from baseline import ref_env
from helpers import TOL_F16, assert_close
from oracle import oracle as O
from scenedino_b200 import _abi
from scenedino_b200 import synthetic as syn

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(ref_env.reference_root() is None, reason="reference tree not installed (baseline/install_ref.py)")]
DEV = "cuda"
C_, HF, WF = 256, 48, 160
IMG_H, IMG_W = 24, 80


class FakeEncoder(torch.nn.Module):
    """Seeded map standing in for DINOv2Module (cf. EncoderDummy, training/trainer_overfit.py:21-30); its dim_reduction is
    built by the reference's own build_dim_reduction (rebound to the native MlpDimReduction by install())."""

    def __init__(self, feat, dinov2_module):
        super().__init__()
        self.latent_size, self.extra_outs = feat.shape[1], 0
        self.dim_reduction = dinov2_module.build_dim_reduction("mlp", 768, 64)
        self.register_buffer("feat", feat)

    def forward(self, x, ground_truth=False):
        if ground_truth:
            return [torch.zeros(x.shape[0], 8, 2, 2, device=x.device)]
        return [self.feat.clone()]

    def expand_dim(self, f):
        return self.dim_reduction.transform_expand(f)


@pytest.fixture(scope="module")
def ref():
    demo = ref_env.import_reference_module("demo_utils.utils")
    ssc = ref_env.import_reference_module("evaluate_model_sscbench")
    models = ref_env.import_reference_module("scenedino.models")
    dmod = ref_env.import_reference_module("scenedino.models.backbones.dino.dinov2_module")
    report = sd.install()
    assert all(v >= 1 for v in report.values()), report
    assert models.BTSNet is sd.BTSNet and demo.NeRFRenderer is sd.NeRFRenderer and demo.ImageRaySampler is sd.ImageRaySampler
    return demo, ssc, models, dmod


def build_net(ref, precision="fp16"):
    demo, ssc, models, dmod = ref
    feat = syn.make_feature_map(31, C_, HF, WF)
    models.make_backbone = lambda conf: FakeEncoder(torch.from_numpy(feat), dmod)
    conf = {"arch": "BTSNet", "predict_dino": True, "dino_dims": 64, "encoder": {"type": "dinov2"},
            "code": {"num_freqs": 6, "freq_factor": 1.5, "include_input": True},
            "decoder_heads": [{"type": "resnet", "name": "normal_head", "freeze": False, "args": {"n_blocks": 0, "d_hidden": 128}}],
            "final_prediction_head": "normal_head", "inv_z": True, "learn_empty": False, "code_mode": "z",
            "sd_precision": precision}
    down = {"type": "segmentation", "n_classes": 27, "gt_classes": 19, "input_dim": 768, "code_dim": 64, "knn_neighbors": 4,
            "buffer_size": 256, "patch_sample_size": 576, "mode": "3d", "apply_crf": False}
    net = models.make_model(conf, down)                     # the reference's factory, native classes
    assert type(net) is sd.BTSNet and type(net.downstream_head) is sd.SemanticHead
    mlp_w, expand, hw = syn.make_mlp(0, bias_scale=0.05), syn.make_expand(3), syn.make_ssc_head(21)
    state = {"heads.normal_head.lin_in.weight": mlp_w[0], "heads.normal_head.lin_in.bias": mlp_w[1],
             "heads.normal_head.lin_out.weight": mlp_w[2], "heads.normal_head.lin_out.bias": mlp_w[3],
             "encoder.dim_reduction.linear_in.weight": expand[0], "encoder.dim_reduction.linear_in.bias": expand[1],
             "encoder.dim_reduction.linear_out.weight": expand[2], "encoder.dim_reduction.linear_out.bias": expand[3],
             "downstream_head.stego_head.linear_path.0.weight": hw["wl"].reshape(64, 768, 1, 1),
             "downstream_head.stego_head.linear_path.0.bias": hw["bl"],
             "downstream_head.stego_head.nonlinear_path.0.weight": hw["wn1"].reshape(768, 768, 1, 1),
             "downstream_head.stego_head.nonlinear_path.0.bias": hw["bn1"],
             "downstream_head.stego_head.nonlinear_path.2.weight": hw["wn2"].reshape(64, 768, 1, 1),
             "downstream_head.stego_head.nonlinear_path.2.bias": hw["bn2"],
             "downstream_head.stego_cluster_head.cluster_centers": hw["centres"],
             "downstream_head.stego_cluster_head.pseudo_assignment": hw["lut"]}
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=False)
    assert not unexpected, unexpected
    net = net.to(DEV).eval()
    osc = O.Scene(feat=feat, K_f=syn.kitti360_K()[None], w2c_f=np.eye(4, dtype=np.float32)[None],
                  rgb=syn.make_images(32, 1, IMG_H, IMG_W), K_c=syn.kitti360_K()[None], w2c_c=np.eye(4, dtype=np.float32)[None])
    return net, osc, O.Mlp(*mlp_w), expand, hw


def encode(net, osc):
    imgs = torch.from_numpy(osc.rgb)[None].to(DEV)          # [1, 1, 3, H, W] in [0, 1]
    K = torch.from_numpy(syn.kitti360_K())[None, None].to(DEV)
    poses = torch.eye(4)[None, None].to(DEV)
    net.encode(imgs * 2 - 1, K, poses, ids_encoder=[0], ids_render=[0], images_alt=imgs)
    net.set_scale(0)
    return K, poses


def labels_match(seg, f_oracle, expand, hw, min_gap=2e-2):
    x = O.expand_dim(f_oracle, *expand)
    oseg, _, oip = O.ssc_head(x, hw["wl"], hw["bl"], hw["wn1"], hw["bn1"], hw["wn2"], hw["bn2"], hw["centres"], hw["lut"])
    top2 = np.sort(oip, 1)[:, -2:]
    clear = top2[:, 1] - top2[:, 0] > min_gap
    assert clear.mean() > 0.5
    assert np.array_equal(np.asarray(seg).ravel()[clear], oseg[clear])


def test_inference_3d_unchanged(ref):
    """demo_utils/utils.py:144-186 (the demo's grid query) on the native net."""
    demo = ref[0]
    net, osc, omlp, expand, hw = build_net(ref)
    encode(net, osc)
    n0 = _abi.launch_count()
    with torch.no_grad():
        xyz, dino_full, sigma, seg = demo.inference_3d(net, (-9, 9), (-3, 1), (3, 33), 0.5)
    assert _abi.launch_count() > n0, "the native kernels ran"
    nx, ny, nz = 37, 9, 61
    assert dino_full.shape == (nx, ny, nz, 768) and sigma.shape == (nx, ny, nz) and seg.shape == (nx, ny, nz)
    pts = xyz[0].cpu().numpy()
    oq = O.query_points(osc, omlp, pts, want_rgb=False)
    assert_close(sigma.reshape(-1).cpu().numpy(), oq["sigma"], TOL_F16, "sigma")
    ox = O.expand_dim(oq["dino"], *expand)
    assert_close(dino_full.reshape(-1, 768).cpu().numpy(), ox, TOL_F16, "dino_full")
    labels_match(seg.cpu().numpy(), oq["dino"], expand, hw)


def test_inference_rendered_2d_unchanged(ref):
    """demo_utils/utils.py:199-236 (the demo's full-image render + expansion + segmentation) on the native renderer."""
    demo = ref[0]
    net, osc, omlp, expand, hw = build_net(ref)
    K, poses = encode(net, osc)
    renderer = demo.NeRFRenderer.from_conf({"n_coarse": 32, "n_fine": 0, "lindisp": True, "eval_batch_size": 65536})
    renderer.hard_alpha_cap = False                          # demo_utils/utils.py:45
    renderer = renderer.bind_parallel(net, gpus=None).eval()
    sampler = demo.ImageRaySampler(z_near=3, z_far=80, width=IMG_W, height=IMG_H)
    view = torch.from_numpy(syn.view_pose_c2w(1))[None, None].to(DEV)      # a stereo-offset novel view
    torch.manual_seed(0)
    u = torch.rand((IMG_H * IMG_W, 32), device=DEV)
    orig = torch.rand_like
    torch.rand_like = lambda *a, **k: u                       # the one draw of the coarse pass (nerf.py:134)
    try:
        with torch.no_grad():
            dino_full, depth, seg = demo.inference_rendered_2d(net, view, K, sampler, renderer)
    finally:
        torch.rand_like = orig
    assert dino_full.shape == (IMG_H, IMG_W, 768) and depth.shape == (IMG_H, IMG_W) and seg.shape == (IMG_H, IMG_W)
    rays = O.gen_rays(syn.view_pose_c2w(1)[None], syn.kitti360_K()[None], IMG_H, IMG_W, 3.0, 80.0)
    lin = np.linspace(0, 1 - 1 / 32, 32, dtype=np.float32)
    oz = O.sample_coarse(rays, u.cpu().numpy(), lin, True)
    orp = O.render_pass(osc, omlp, rays, oz)
    assert_close(depth.reshape(-1).cpu().numpy(), orp["depth"], TOL_F16, "depth")
    ox = O.expand_dim(orp["dino_features"], *expand)
    assert_close(dino_full.reshape(-1, 768).cpu().numpy(), ox, TOL_F16, "dino_full")
    labels_match(seg.cpu().numpy(), orp["dino_features"], expand, hw)


def test_sscbench_downsample_and_predict_unchanged(ref):
    """sscbench/evaluate_model_sscbench.py:660-758, 829-854: the whole 256 x 256 x 32 grid in the script's own four chunks,
    alpha-weighted argmax and max-pool "grow" included, on the native net."""
    ssc = ref[1]
    net, osc, omlp, expand, hw = build_net(ref)
    pts_np = syn.ssc_voxel_grid()
    pts = torch.from_numpy(pts_np).to(DEV)
    data = {"imgs": [torch.from_numpy(osc.rgb[0] * 2 - 1)], "poses": [np.eye(4, dtype=np.float32)], "projs": [syn.kitti360_K()]}
    n0 = _abi.launch_count()
    with torch.no_grad():
        sigmas, segs, dino = ssc.downsample_and_predict(data, net, pts, 1, "stego_kmeans")
    assert _abi.launch_count() - n0 >= 8 and dino is None
    assert sigmas.shape == segs.shape == (256, 256, 32)
    # the script's post-processing of OUR per-voxel results: max-pool grow of sigma, label where alpha > 0
    with torch.no_grad():
        _, _, sig_direct, seg_direct = net(pts[None], predict_segmentation=True, prediction_mode="stego_kmeans")
    grown = torch.nn.functional.max_pool3d(sig_direct.reshape(1, 256, 256, 32).cpu(), kernel_size=3, stride=1, padding=1)[0].numpy()
    assert np.array_equal(sigmas, grown)
    assert np.array_equal(segs.astype(np.int64), seg_direct[0].argmax(-1).reshape(256, 256, 32).cpu().numpy())
    # and those per-voxel results against the oracle on a strided subset
    sub = np.arange(0, len(pts_np), 509)
    oq = O.query_points(osc, omlp, pts_np[sub], want_rgb=False)
    assert_close(sig_direct.reshape(-1).cpu().numpy()[sub], oq["sigma"], TOL_F16, "sigma")
    labels_match(seg_direct[0].argmax(-1).cpu().numpy()[sub], oq["dino"], expand, hw)
