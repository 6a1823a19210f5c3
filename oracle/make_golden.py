"""TEST INFRASTRUCTURE: generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Inputs that are cheap to regenerate (feature maps, images) come from numpy seeds in
scenedino_b200/synthetic.py and are NOT stored; the fixtures hold the small inputs (cameras, MLP
weights, points, rays, the random draws torch made inside the reference) and the reference's outputs.
A checksum of every regenerated input is stored so that generator drift is detected.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from scenedino_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

C, HF, WF = 256, 48, 160
HC, WC = 96, 320
FEAT_SEED, IMG_SEED, MLP_SEED, EXP_SEED = 11, 12, 13, 14


class DrawRecorder:
    """Records every random tensor the reference draws (rand_like / rand / randn_like) and the
    indices torch.searchsorted returns, in call order."""

    def __init__(self):
        self.draws, self.searches = [], []

    def __enter__(self):
        self._o = (torch.rand_like, torch.rand, torch.randn_like, torch.searchsorted)
        rec = self

        def rand_like(*a, **k):
            t = rec._o[0](*a, **k); rec.draws.append(("rand_like", t.clone())); return t

        def rand(*a, **k):
            t = rec._o[1](*a, **k); rec.draws.append(("rand", t.clone())); return t

        def randn_like(*a, **k):
            t = rec._o[2](*a, **k); rec.draws.append(("randn_like", t.clone())); return t

        def searchsorted(*a, **k):
            t = rec._o[3](*a, **k); rec.searches.append(t.clone()); return t

        torch.rand_like, torch.rand, torch.randn_like, torch.searchsorted = rand_like, rand, randn_like, searchsorted
        return self

    def __exit__(self, *exc):
        torch.rand_like, torch.rand, torch.randn_like, torch.searchsorted = self._o


def checksum(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a, np.float64).ravel()
    return np.array([a.sum(), np.abs(a).sum(), a[:: max(1, a.size // 97)].sum()], np.float64)


def scene_inputs(nv_c: int, n: int = 1):
    feat = np.concatenate([syn.make_feature_map(FEAT_SEED + i, C, HF, WF) for i in range(n)], 0)
    imgs = np.stack([syn.make_images(IMG_SEED + i, nv_c, HC, WC) for i in range(n)], 0)
    K = np.broadcast_to(syn.kitti360_K(), (n, nv_c, 3, 3)).copy()
    c2w = np.stack([np.stack([syn.view_pose_c2w(v + 2 * i) if v else syn.view_pose_c2w(0)
                              for v in range(nv_c)], 0) for i in range(n)], 0)
    return feat, imgs, K, c2w


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def build_net(ref, feat, mlp, expand, learn_empty=False, empty=None):
    net = ref_shim.build_reference_net(ref, t(feat), *[t(w) for w in mlp], learn_empty=learn_empty,
                                       empty_feature=None if empty is None else t(empty),
                                       expand_weights=[t(w) for w in expand])
    return net


def encode(net, imgs, K, c2w, ids_render):
    n, nv = imgs.shape[:2]
    dummy = torch.zeros(n, nv, 3, 8, 8)
    net.encode(dummy, t(K), t(c2w), ids_encoder=[0], ids_render=ids_render, images_alt=t(imgs))
    net.set_scale(0)


def golden_query(ref):
    nv_c = 2
    feat, imgs, K, c2w = scene_inputs(nv_c)
    mlp = syn.make_mlp(MLP_SEED, bias_scale=0.1)
    expand = syn.make_expand(EXP_SEED)
    pts = np.concatenate([
        syn.random_points(21, 1536),
        syn.ssc_voxel_grid()[::4099][:512],
        np.array([[0, 0, 0], [0, 0, 1e-3], [0, 0, -5], [1.3, 0.1, 1.0], [0, 0, 3], [5, 1, 5e-4]], np.float32),
    ], 0).astype(np.float32)
    out = {}
    for tag, learn_empty in (("", False), ("_le", True)):
        empty = np.random.RandomState(31).standard_normal(C).astype(np.float32) if learn_empty else None
        net = build_net(ref, feat, mlp, expand, learn_empty, empty)
        encode(net, imgs, K, c2w, [0, 1])
        with torch.no_grad():
            xyz = t(pts)[None]
            cam = ref.pinhole.pts_into_camera(xyz, net.grid_f_poses_w2c)
            xy, z = ref.pinhole.project_to_image(cam, net.grid_f_Ks)
            inv = ref.pinhole.outside_frustum(xy, z)
            sf, sinv = net.sample_features(xyz)
            rgb, invalid, sigma, extras, sd = net(xyz)
            raw = net.heads["normal_head"](sf.flatten(0, 1)).reshape(1, -1, 65)
            dino_full, _, sigma2, seg = net(xyz[:, :256], predict_segmentation=True)
        if learn_empty:
            out["empty_feature"] = empty
        else:
            out.update(xy=xy[0, 0].numpy(), z=z[0, 0, :, 0].numpy(), frustum_invalid=inv[0, 0, :, 0].numpy())
        out.update({
            "sample_features" + tag: sf[0, :768, 0].numpy(),
            "sample_invalid" + tag: sinv[0, :, 0].numpy(),
            "mlp_raw" + tag: raw[0].numpy(),
            "sigma" + tag: sigma[0, :, 0].numpy(),
            "dino" + tag: sd["dino_features"][0].numpy(),
            "rgb" + tag: rgb[0].numpy(),
            "invalid" + tag: invalid[0].numpy(),
            "invalid_features" + tag: sd["invalid_features"][0, :, 0].numpy(),
            "dino_full" + tag: dino_full[0].numpy(),
            "sigma_seg" + tag: sigma2[0, :, 0].numpy(),
        })
        assert seg is None and extras is None
    out.update(points=pts, K=K[0], c2w=c2w[0], w_in=mlp[0], b_in=mlp[1], w_out=mlp[2], b_out=mlp[3],
               e_w1=expand[0], e_b1=expand[1], e_w2=expand[2], e_b2=expand[3],
               feat_checksum=checksum(feat), img_checksum=checksum(imgs),
               shape=np.array([C, HF, WF, HC, WC, nv_c]))
    np.savez_compressed(os.path.join(OUT, "query.npz"), **out)
    print("query.npz:", {k: v.shape for k, v in out.items()})


def golden_query_big(ref):
    """A query big enough for the texel sort + projected-map tile kernel (>= 65 536 points): the reference's outputs for a
    strided subset of 70 001 points (regenerated from their seed by the test), the frustum mask for all of them."""
    nv_c = 2
    feat, imgs, K, c2w = scene_inputs(nv_c)
    mlp = syn.make_mlp(MLP_SEED, bias_scale=0.1)
    expand = syn.make_expand(EXP_SEED)
    n_pts, pts_seed, stride = 70001, 41, 35
    pts = syn.random_points(pts_seed, n_pts)
    sub = np.arange(0, n_pts, stride)
    out = {}
    for tag, learn_empty in (("", False), ("_le", True)):
        empty = np.random.RandomState(31).standard_normal(C).astype(np.float32) if learn_empty else None
        net = build_net(ref, feat, mlp, expand, learn_empty, empty)
        encode(net, imgs, K, c2w, [0, 1])
        with torch.no_grad():
            rgb, invalid, sigma, extras, sd = net(t(pts)[None])
        if learn_empty:
            out["empty_feature"] = empty
        out.update({
            "sigma" + tag: sigma[0, sub, 0].numpy(),
            "dino" + tag: sd["dino_features"][0, sub].numpy(),
            "invalid_features" + tag: np.packbits(sd["invalid_features"][0, :, 0].numpy()),
        })
    out.update(n_pts=np.array(n_pts), pts_seed=np.array(pts_seed), stride=np.array(stride), pts_checksum=checksum(pts),
               K=K[0], c2w=c2w[0], w_in=mlp[0], b_in=mlp[1], w_out=mlp[2], b_out=mlp[3],
               feat_checksum=checksum(feat), img_checksum=checksum(imgs), shape=np.array([C, HF, WF, HC, WC, nv_c]))
    np.savez_compressed(os.path.join(OUT, "query_big.npz"), **out)
    print("query_big.npz:", {k: v.shape for k, v in out.items()})


def pick_rays(c2w_list, K, n_each, seed):
    rs = np.random.RandomState(seed)
    rays = []
    for i, c2w in enumerate(c2w_list):
        r = syn.image_rays(c2w, K, syn.IMG_H, syn.IMG_W, frame_id=float(i))
        rays.append(r[rs.choice(len(r), n_each, replace=False)])
    return np.concatenate(rays, 0)


def run_renderer(ref, net, rays, conf, *, hard_alpha_cap, seed, sample_from_dist=None, training=False):
    ren = ref.NeRFRenderer.from_conf(conf)
    ren.hard_alpha_cap = hard_alpha_cap
    wrapped = ren.bind_parallel(net, gpus=None).eval()
    if training:
        wrapped.train()
    torch.manual_seed(seed)
    with DrawRecorder() as rec, torch.no_grad():
        out = wrapped(t(rays), want_weights=True, want_alphas=True, want_z_samps=True,
                      want_rgb_samps=True, sample_from_dist=sample_from_dist)
    return out, rec


def flat(prefix, d, out):
    for k, v in d.items():
        if isinstance(v, dict):
            flat(prefix + k + ".", v, out)
        elif torch.is_tensor(v):
            out[prefix + k] = v.numpy()


def golden_render(ref):
    nv_c = 2
    feat, imgs, K, c2w = scene_inputs(nv_c)
    mlp = syn.make_mlp(MLP_SEED, bias_scale=0.1)
    expand = syn.make_expand(EXP_SEED)
    net = build_net(ref, feat, mlp, expand)
    encode(net, imgs, K, c2w, [0, 1])
    novel = syn.view_pose_c2w(3)
    rays = pick_rays([c2w[0, 0], novel], K[0, 0], 96, seed=5)[None]  # [1,192,11]
    common = dict(K=K[0], c2w=c2w[0], w_in=mlp[0], b_in=mlp[1], w_out=mlp[2], b_out=mlp[3],
                  feat_checksum=checksum(feat), img_checksum=checksum(imgs),
                  shape=np.array([C, HF, WF, HC, WC, nv_c]))

    # -- coarse only, cfg-1 style (Kc=64, lindisp, hard_alpha_cap off as in demo_utils/utils.py:45)
    out, rec = run_renderer(ref, net, rays, {"n_coarse": 64, "n_fine": 0, "lindisp": True,
                                             "eval_batch_size": 4096}, hard_alpha_cap=False, seed=100)
    g = dict(common, rays=rays, u_coarse=rec.draws[0][1].numpy(),
             lin=torch.linspace(0, 1 - 1.0 / 64, 64).numpy())
    assert len(rec.draws) == 1
    flat("", out, g)
    np.savez_compressed(os.path.join(OUT, "render_coarse.npz"), **g)
    print("render_coarse.npz:", sorted(g.keys()))

    # -- coarse + fine + depth samples, training-shaped (hard_alpha_cap on)
    for name, lindisp, white in (("render_fine", True, False), ("render_fine_lin", False, True)):
        conf = {"n_coarse": 32, "n_fine": 16, "n_fine_depth": 4, "lindisp": lindisp, "depth_std": 0.5,
                "white_bkgd": white, "eval_batch_size": 100000, "hard_alpha_cap": True}
        out, rec = run_renderer(ref, net, rays, conf, hard_alpha_cap=True, seed=101)
        kinds = [k for k, _ in rec.draws]
        assert kinds == ["rand_like", "rand", "rand_like", "randn_like"], kinds
        g = dict(common, rays=rays, u_coarse=rec.draws[0][1].numpy(), u_fine0=rec.draws[1][1].numpy(),
                 u_fine1=rec.draws[2][1].numpy(), n_depth=rec.draws[3][1].numpy(),
                 fine_inds=(rec.searches[0] - 1).clamp_min(0).numpy().astype(np.int32),
                 lin=torch.linspace(0, 1 - 1.0 / 32, 32).numpy(),
                 conf=np.array([32, 16, 4, int(lindisp), int(white)]), depth_std=np.float32(0.5))
        flat("", out, g)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **g)
        print(name + ".npz:", sorted(g.keys()))

    # -- proposal resampling (sample_from_dist), nerf.py:143-179,485-490
    rs = np.random.RandomState(7)
    Kp = 48
    pw = rs.uniform(0, 1, (1, 192, Kp)).astype(np.float32) ** 4
    pz = np.sort(rs.uniform(3, 80, (1, 192, Kp)).astype(np.float32), -1)
    out, rec = run_renderer(ref, net, rays, {"n_coarse": 40, "n_fine": 0, "lindisp": True,
                                             "eval_batch_size": 100000}, hard_alpha_cap=False,
                            seed=102, sample_from_dist=(t(pw), t(pz)))
    kinds = [k for k, _ in rec.draws]
    assert kinds == ["rand", "rand_like"], kinds
    g = dict(common, rays=rays, prop_weights=pw, prop_z=pz, u0=rec.draws[0][1].numpy(),
             u1=rec.draws[1][1].numpy(),
             inds=(rec.searches[0] - 1).clamp(0, 39).numpy().astype(np.int32))
    flat("", out, g)
    np.savez_compressed(os.path.join(OUT, "render_from_dist.npz"), **g)
    print("render_from_dist.npz:", sorted(g.keys()))


def golden_render_d768(ref):
    """SURVEY 8d cfg 3 in small: the D = 768 head (d_out = 769), four colour views, Kc = 64 coarse + Kf = 32 importance
    samples (n_fine_depth = 0) -> a 96-sample fine pass, training-shaped (hard_alpha_cap on)."""
    nv_c, D = 4, 768
    feat, imgs, K, c2w = scene_inputs(nv_c)
    mlp = syn.make_mlp(MLP_SEED + 7, d_out=1 + D, bias_scale=0.1)
    net = ref_shim.build_reference_net(ref, t(feat), *[t(w) for w in mlp], dino_dims=D)
    encode(net, imgs, K, c2w, list(range(nv_c)))
    rays = pick_rays([c2w[0, 0], syn.view_pose_c2w(3)], K[0, 0], 24, seed=9)[None]      # [1,48,11]
    conf = {"n_coarse": 64, "n_fine": 32, "n_fine_depth": 0, "lindisp": True, "eval_batch_size": 100000, "hard_alpha_cap": True}
    ren = ref.NeRFRenderer.from_conf(conf)
    ren.hard_alpha_cap = True
    wrapped = ren.bind_parallel(net, gpus=None).eval()
    torch.manual_seed(103)
    with DrawRecorder() as rec, torch.no_grad():
        out = wrapped(t(rays), want_weights=True, want_alphas=True, want_z_samps=True)
    kinds = [k for k, _ in rec.draws]
    assert kinds == ["rand_like", "rand", "rand_like"], kinds
    g = dict(K=K[0], c2w=c2w[0], w_in=mlp[0], b_in=mlp[1], w_out=mlp[2], b_out=mlp[3], feat_checksum=checksum(feat),
             img_checksum=checksum(imgs), shape=np.array([C, HF, WF, HC, WC, nv_c]), rays=rays,
             u_coarse=rec.draws[0][1].numpy(), u_fine0=rec.draws[1][1].numpy(), u_fine1=rec.draws[2][1].numpy(),
             fine_inds=(rec.searches[0] - 1).clamp_min(0).numpy().astype(np.int32),
             lin=torch.linspace(0, 1 - 1.0 / 64, 64).numpy(), conf=np.array([64, 32, 0, 1, 0]))
    flat("", out, g)
    assert g["fine.dino_features"].shape == (1, 48, D) and g["fine.z_samps"].shape == (1, 48, 96)
    np.savez_compressed(os.path.join(OUT, "render_d768.npz"), **g)
    print("render_d768.npz:", sorted(g.keys()))


def golden_superbatch(ref):
    """sb = 2 scenes with different feature maps / cameras, 4 colour views (training-shaped)."""
    nv_c, n = 4, 2
    feat, imgs, K, c2w = scene_inputs(nv_c, n)
    mlp = syn.make_mlp(MLP_SEED, bias_scale=0.1)
    expand = syn.make_expand(EXP_SEED)
    net = build_net(ref, feat, mlp, expand)
    encode(net, imgs, K, c2w, [0, 1, 2, 3])
    rays = np.stack([pick_rays([c2w[i, 1], c2w[i, 2]], K[i, 0], 40, seed=50 + i) for i in range(n)], 0)
    out, rec = run_renderer(ref, net, rays, {"n_coarse": 32, "n_fine": 0, "lindisp": True,
                                             "eval_batch_size": 1 << 20, "hard_alpha_cap": True},
                            hard_alpha_cap=True, seed=103)
    g = dict(K=K, c2w=c2w, w_in=mlp[0], b_in=mlp[1], w_out=mlp[2], b_out=mlp[3], rays=rays,
             u_coarse=rec.draws[0][1].numpy(), lin=torch.linspace(0, 1 - 1.0 / 32, 32).numpy(),
             feat_checksum=checksum(feat), img_checksum=checksum(imgs),
             shape=np.array([C, HF, WF, HC, WC, nv_c]))
    flat("", out, g)
    np.savez_compressed(os.path.join(OUT, "render_superbatch.npz"), **g)
    print("render_superbatch.npz:", {k: v.shape for k, v in g.items()})


def golden_rays(ref):
    """ImageRaySampler.sample (ray_sampler.py:439-513) on two batch elements of two views, and on
    odd image sizes with explicit frame ids and unnormalised directions."""
    out = {}
    K = syn.kitti360_K()
    for tag, (n, v, H, W, ids, norm) in {"a": (2, 2, 24, 80, None, True), "b": (1, 3, 17, 37, [4, 7, 9], False)}.items():
        c2w = np.stack([np.stack([syn.view_pose_c2w(j + 3 * i) for j in range(v)], 0) for i in range(n)], 0).astype(np.float32)
        proj = np.broadcast_to(K, (n, v, 3, 3)).astype(np.float32).copy()
        proj[:, 1:, 0, 2] += 0.03          # views with different principal points / focal lengths
        proj[:, 1:, 1, 1] *= 1.1
        imgs = np.stack([syn.make_images(IMG_SEED + i, v, H, W) for i in range(n)], 0)
        s = ref.ImageRaySampler(3.0, 80.0, H, W, norm_dir=norm)
        rays, rgb_gt = s.sample(t(imgs), t(c2w), t(proj), image_ids=ids)
        assert rays.shape == (n, v * H * W, 11) and rgb_gt.shape == (n, v * H * W, 3)
        out.update({f"{tag}_c2w": c2w, f"{tag}_proj": proj, f"{tag}_hw": np.array([H, W]), f"{tag}_norm_dir": np.array(norm),
                    f"{tag}_ids": np.array(ids if ids is not None else [], np.float32), f"{tag}_rays": rays.numpy(),
                    f"{tag}_img_sum": checksum(imgs), f"{tag}_rgb_gt_sum": checksum(rgb_gt.numpy())})
    out["z"] = np.array([3.0, 80.0], np.float32)
    np.savez_compressed(os.path.join(OUT, "rays.npz"), **out)
    print("rays.npz", {k: v.shape for k, v in out.items() if k.endswith("rays")})


def golden_ssc_head(ref):
    """SemanticHead.forward(mode="stego_kmeans") (downstream_head/semantic_head.py:107-112) on the 768-d
    expansion of queried features.  SemanticHead's constructor allocates its training buffers on "cuda"
    (semantic_head.py:69-70), so its own submodules (StegoClusterHead, KMeansParamHead) are composed here in the order
    of its forward; eval mode (no dropout)."""
    from scenedino.downstream_head.semantic_head import KMeansParamHead, StegoClusterHead, _norm
    q = np.load(os.path.join(OUT, "query.npz"))
    x = syn.ssc_head_inputs(q["dino_full"], q["dino_full_le"])             # [577, 768]; regenerated by the tests, not stored
    w = syn.make_ssc_head(21)                                              # likewise (checksums stored)
    n_cls, d_code = w["centres"].shape
    d_in = w["wl"].shape[1]
    stego = StegoClusterHead(d_in, d_code).eval()
    km = KMeansParamHead(n_cls, 19, d_code).eval()
    with torch.no_grad():
        stego.linear_path[0].weight.copy_(t(w["wl"]).view(d_code, d_in, 1, 1)); stego.linear_path[0].bias.copy_(t(w["bl"]))
        stego.nonlinear_path[0].weight.copy_(t(w["wn1"]).view(d_in, d_in, 1, 1)); stego.nonlinear_path[0].bias.copy_(t(w["bn1"]))
        stego.nonlinear_path[2].weight.copy_(t(w["wn2"]).view(d_code, d_in, 1, 1)); stego.nonlinear_path[2].bias.copy_(t(w["bn2"]))
        km.cluster_centers.copy_(t(w["centres"]))
        km.pseudo_assignment.copy_(t(w["lut"]))
        feats = _norm(t(x)[None])                                   # [1, N, 768] as BTSNet.forward hands it over (bts.py:586-589)
        code = stego(feats)
        res = km(code)
        ip = torch.nn.functional.normalize(code.flatten(0, -2), dim=1) @ torch.nn.functional.normalize(km.cluster_centers, dim=1).t()
    out = dict(seed=np.array(21), x_sum=checksum(x), w_sum=np.stack([checksum(w[k]) for k in sorted(w)]),
               code=code[0].numpy(), seg=res["segs_pred"][0].numpy(), pseudo=res["pseudo_segs_pred"][0].numpy(), ip=ip.numpy())
    assert out["seg"].dtype == np.int64 and out["seg"].shape == (x.shape[0],)
    np.savez_compressed(os.path.join(OUT, "ssc_head.npz"), **out)
    print("ssc_head.npz", {k: v.shape for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = ref_shim.import_reference()
    only = sys.argv[1:]          # e.g. `python oracle/make_golden.py query_big` regenerates one fixture
    jobs = {"query": golden_query, "query_big": golden_query_big, "render": golden_render, "superbatch": golden_superbatch,
            "rays": golden_rays, "ssc_head": golden_ssc_head, "render_d768": golden_render_d768}
    for name, fn in jobs.items():
        if not only or name in only:
            fn(ref)


if __name__ == "__main__":
    main()
