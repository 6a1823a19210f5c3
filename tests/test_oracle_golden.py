"""Pins the CPU oracle (oracle/sd_oracle.c) against golden vectors produced by the unmodified
reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import TOL_FP32, assert_close, big_query_points, golden_scene_arrays, rel_err
from oracle import oracle as O


def _w2c(c2w):
    # the reference inverts poses with torch.inverse in fp32 (models/bts.py:125-126)
    return torch.inverse(torch.from_numpy(np.ascontiguousarray(c2w))).numpy()


def _scene(g, feat, imgs, b=None, **kw):
    K = g["K"] if b is None else g["K"][b]
    c2w = g["c2w"] if b is None else g["c2w"][b]
    w2c = _w2c(c2w)
    i = 0 if b is None else b
    return O.Scene(feat=feat[i:i + 1], K_f=K[:1], w2c_f=w2c[:1], rgb=imgs[i], K_c=K, w2c_c=w2c, **kw)


def _mlp(g):
    return O.Mlp(g["w_in"], g["b_in"], g["w_out"], g["b_out"])


def test_projection_and_mask_bit_exact(golden):
    g = golden("query")
    w2c = _w2c(g["c2w"])
    xy, z, inv = O.project(g["K"][0], w2c[0], g["points"])
    assert np.array_equal(inv, g["frustum_invalid"])          # bit-exact mask (a-9)
    assert np.array_equal(xy, g["xy"]) and np.array_equal(z, g["z"])
    assert 0.15 < inv.mean() < 0.6 and (g["z"] <= 1e-3).sum() >= 3   # the case is non-trivial


@pytest.mark.parametrize("tag,learn_empty", [("", False), ("_le", True)])
def test_query_points(golden, tag, learn_empty):
    g = golden("query")
    feat, imgs = golden_scene_arrays(g)
    sc = _scene(g, feat, imgs, learn_empty=learn_empty,
                empty_feature=g["empty_feature"] if learn_empty else None)
    sf, sinv = O.sample_features(sc, g["points"])
    assert np.array_equal(sinv[:, 0], g["sample_invalid" + tag])
    n = g["sample_features" + tag].shape[0]
    assert np.array_equal(sf[:n, 0, :256], g["sample_features" + tag][:, :256])   # gather (a-12/13)
    assert_close(sf[:n, 0, 256:], g["sample_features" + tag][:, 256:], 1e-6, "xyz code")  # a-10/11
    q = O.query_points(sc, _mlp(g), g["points"], want_raw=True)
    assert_close(q["raw"], g["mlp_raw" + tag], TOL_FP32, "mlp output")            # a-15
    assert_close(q["sigma"], g["sigma" + tag], TOL_FP32, "sigma")                 # a-16
    assert_close(q["dino"], g["dino" + tag], TOL_FP32, "dino")
    assert np.array_equal(q["rgb"], g["rgb" + tag])                               # a-17
    assert np.array_equal(q["invalid"], g["invalid" + tag])                       # a-18
    assert np.array_equal(q["invalid_features"], g["invalid_features" + tag])
    ex = O.expand_dim(g["dino" + tag][:256], g["e_w1"], g["e_b1"], g["e_w2"], g["e_b2"])
    assert_close(ex, g["dino_full" + tag], TOL_FP32, "expand_dim")                # 8f-1


def _check_pass(o, g, prefix, tol=TOL_FP32):
    R = g["rays"].shape[0] * g["rays"].shape[1]
    K = g[prefix + "z_samps"].shape[-1]
    assert_close(o["weights"], g[prefix + "weights"].reshape(R, K), tol, prefix + "weights")
    assert_close(o["alphas"], g[prefix + "alphas"].reshape(R, K), tol, prefix + "alphas")
    assert_close(o["depth"], g[prefix + "depth"].reshape(R), tol, prefix + "depth")
    assert_close(o["rgb"], g[prefix + "rgb"].reshape(R, -1), tol, prefix + "rgb")
    assert_close(o["dino_features"], g[prefix + "dino_features"].reshape(R, -1), tol, prefix + "dino")
    assert np.array_equal(o["invalid"], g[prefix + "invalid"].reshape(R, K, -1))


@pytest.mark.parametrize("tag,learn_empty", [("", False), ("_le", True)])
def test_query_points_big(golden, tag, learn_empty):
    """70 001 points (the size class that takes the texel sort + tile kernel on the GPU): mask of every point bit-exact,
    densities / features of the stored subset within the fp32 bar."""
    g = golden("query_big")
    feat, imgs = golden_scene_arrays(g)
    pts, sub = big_query_points(g)
    kw = dict(learn_empty=learn_empty, empty_feature=g["empty_feature"] if learn_empty else None)
    o = O.query_points(_scene(g, feat, imgs, **kw), _mlp(g), pts, want_rgb=False)
    inv = np.unpackbits(g["invalid_features" + tag])[:len(pts)].astype(bool)
    assert np.array_equal(o["invalid_features"], inv)
    assert_close(o["sigma"][sub], g["sigma" + tag], TOL_FP32, "sigma")
    assert_close(o["dino"][sub], g["dino" + tag], TOL_FP32, "dino")


def test_render_coarse(golden):
    g = golden("render_coarse")
    feat, imgs = golden_scene_arrays(g)
    sc = _scene(g, feat, imgs)
    rays = g["rays"][0]
    z = O.sample_coarse(rays, g["u_coarse"], g["lin"], lindisp=True)
    assert np.array_equal(z, g["coarse.z_samps"][0])                               # a-1 bit-exact
    o = O.render_pass(sc, _mlp(g), rays, z, hard_alpha_cap=False, want_rgb_samps=True)
    _check_pass(o, g, "coarse.")
    # _format_outputs reshapes invalid_features with invalid's last dim (nv_c), nerf.py:564,596:
    # for nv_c = 2 the reference returns [1, R/2, K, 2]; the flat order is still [R, K].
    assert g["coarse.invalid_features"].shape == (1, rays.shape[0] // 2, 64, 2)
    assert np.array_equal(o["invalid_features"].ravel(), g["coarse.invalid_features"].ravel())
    assert np.array_equal(o["rgb_samps"], g["coarse.rgb_samps"][0])
    assert np.array_equal(g["coarse.ray_info"][0], rays[:, 8:])


@pytest.mark.parametrize("name", ["render_fine", "render_fine_lin"])
def test_render_fine(golden, name):
    g = golden(name)
    feat, imgs = golden_scene_arrays(g)
    sc = _scene(g, feat, imgs)
    rays = g["rays"][0]
    Kc, Kf, Kfd, lindisp, white = [int(v) for v in g["conf"]]
    out = O.render_rays(sc, _mlp(g), rays, lin=g["lin"], u_coarse=g["u_coarse"], u_fine0=g["u_fine0"],
                        u_fine1=g["u_fine1"], n_depth=g["n_depth"], depth_std=float(g["depth_std"]),
                        lindisp=bool(lindisp), hard_alpha_cap=True, white_bkgd=bool(white))
    assert np.array_equal(out["coarse"]["z_samps"], g["coarse.z_samps"][0])
    _check_pass(out["coarse"], g, "coarse.")
    # a-2: importance-sample indices.  The oracle's normaliser is a sequential double sum, torch's a
    # vectorised fp32 sum, so an index may flip only where u sits within an ulp of a CDF entry.
    flips = (out["fine_inds"] != g["fine_inds"]).sum()
    assert flips <= max(1, int(1e-4 * g["fine_inds"].size)), f"{flips} index flips"
    # sample_fine on the GOLDEN coarse weights must reproduce the golden fine z-samples bit for bit
    # wherever the index agrees
    zf, inds = O.sample_fine(rays, g["coarse.weights"][0], g["u_fine0"], g["u_fine1"], bool(lindisp))
    zd = O.sample_fine_depth(rays, g["coarse.depth"][0], g["n_depth"], float(g["depth_std"]))
    zall = O.sort_rows(np.concatenate([g["coarse.z_samps"][0], zf, zd], 1))
    same_rows = (inds == g["fine_inds"]).all(1)
    assert same_rows.mean() > 0.99
    assert np.array_equal(zall[same_rows], g["fine.z_samps"][0][same_rows])        # a-2,a-3,a-5
    o = O.render_pass(sc, _mlp(g), rays, g["fine.z_samps"][0], hard_alpha_cap=True, white_bkgd=bool(white))
    _check_pass(o, g, "fine.")


def test_render_d768(golden):
    """SURVEY 8d cfg 3 in small: the 768-d head (d_out = 769), four colour views, 64 coarse + 32 importance samples
    merged into a 96-sample fine pass, hard_alpha_cap on."""
    g = golden("render_d768")
    feat, imgs = golden_scene_arrays(g)
    sc = _scene(g, feat, imgs)
    rays = g["rays"][0]
    assert g["w_out"].shape == (769, 128) and g["K"].shape[0] == 4
    out = O.render_rays(sc, _mlp(g), rays, lin=g["lin"], u_coarse=g["u_coarse"], u_fine0=g["u_fine0"], u_fine1=g["u_fine1"],
                        lindisp=True, hard_alpha_cap=True)
    assert np.array_equal(out["coarse"]["z_samps"], g["coarse.z_samps"][0])
    _check_pass(out["coarse"], g, "coarse.")
    assert out["coarse"]["dino_features"].shape == (48, 768)
    flips = (out["fine_inds"] != g["fine_inds"]).sum()
    assert flips <= 1, f"{flips} index flips"
    o = O.render_pass(sc, _mlp(g), rays, g["fine.z_samps"][0], hard_alpha_cap=True)
    _check_pass(o, g, "fine.")
    assert np.array_equal(o["invalid_features"].ravel(), g["fine.invalid_features"].ravel())


def test_render_from_dist(golden):
    g = golden("render_from_dist")
    feat, imgs = golden_scene_arrays(g)
    sc = _scene(g, feat, imgs)
    rays = g["rays"][0]
    z, inds = O.sample_coarse_from_dist(g["prop_weights"][0], g["prop_z"][0], g["u0"], g["u1"], True)
    flips = (inds != g["inds"]).sum()
    assert flips <= 1
    zs = O.sort_rows(z)
    same = (inds == g["inds"]).all(1)
    assert np.array_equal(zs[same], g["coarse.z_samps"][0][same])                  # a-4 bit-exact
    o = O.render_pass(sc, _mlp(g), rays, g["coarse.z_samps"][0])
    _check_pass(o, g, "coarse.")


def test_render_superbatch(golden):
    g = golden("render_superbatch")
    feat, imgs = golden_scene_arrays(g, n=2)
    u = g["u_coarse"].reshape(2, -1, g["u_coarse"].shape[-1])
    for b in range(2):
        sc = _scene(g, feat, imgs, b=b)
        rays = g["rays"][b]
        z = O.sample_coarse(rays, u[b], g["lin"], True)
        assert np.array_equal(z, g["coarse.z_samps"][b])
        o = O.render_pass(sc, _mlp(g), rays, z, hard_alpha_cap=True)
        gb = {k: v[b:b + 1] for k, v in g.items() if k.startswith("coarse.")}
        gb["rays"] = g["rays"][b:b + 1]
        _check_pass(o, gb, "coarse.")


def test_composite_properties():
    """Size-independent properties of a-19: weights are a sub-partition of unity; hard_alpha_cap
    makes them sum to one; a delta-function density returns the sample's own depth/feature."""
    rs = np.random.RandomState(0)
    R, K, D = 64, 48, 8
    z = np.sort(rs.uniform(3, 80, (R, K)).astype(np.float32), 1)
    sigma = rs.uniform(0, 0.3, (R, K)).astype(np.float32)
    feat = rs.standard_normal((R, K, D)).astype(np.float32)
    o = O.composite(z, sigma, feat, None, hard_alpha_cap=False)
    assert (o["weights"] >= 0).all() and (o["weights"].sum(1) <= 1 + 1e-5).all()
    o = O.composite(z, sigma * 0, feat, None, hard_alpha_cap=True)
    np.testing.assert_allclose(o["weights"].sum(1), 1.0, rtol=1e-6)
    np.testing.assert_allclose(o["depth"], z[:, -1], rtol=1e-6)
    spike = np.zeros_like(sigma); spike[:, 7] = 1e4
    o = O.composite(z, spike, feat, None)
    np.testing.assert_allclose(o["depth"], z[:, 7], rtol=1e-5)
    np.testing.assert_allclose(o["dino"], feat[:, 7], rtol=1e-5, atol=1e-6)


def _ray_cases(g):
    for tag in "ab":
        H, W = (int(x) for x in g[f"{tag}_hw"])
        ids = g[f"{tag}_ids"]
        yield tag, H, W, (ids if ids.size else None), bool(g[f"{tag}_norm_dir"])


def test_gen_rays_bit_exact(golden):
    """ImageRaySampler.sample's rays (ray_sampler.py:439-486) -- every column bit-for-bit, including the
    linspace pixel centres (odd sizes, explicit frame ids and unnormalised directions in case b)."""
    g = golden("rays")
    for tag, H, W, ids, norm in _ray_cases(g):
        for i in range(g[f"{tag}_c2w"].shape[0]):
            r = O.gen_rays(g[f"{tag}_c2w"][i], g[f"{tag}_proj"][i], H, W, float(g["z"][0]), float(g["z"][1]), frame_ids=ids,
                           norm_dir=norm)
            assert np.array_equal(r, g[f"{tag}_rays"][i]), (tag, i)


def test_gen_rays_properties():
    """Size-independent properties at the full 192 x 640 image: unit directions, pixel columns symmetric around 0,
    and the projection of origin + t * direction lands back on the pixel."""
    from scenedino_b200 import synthetic as syn
    K = syn.kitti360_K().astype(np.float32)
    c2w = np.stack([syn.view_pose_c2w(v) for v in range(2)], 0).astype(np.float32)
    H, W = 192, 640
    r = O.gen_rays(c2w, np.stack([K, K]), H, W, 3.0, 80.0).reshape(2, H, W, 11)
    assert np.allclose(np.linalg.norm(r[..., 3:6], axis=-1), 1.0, atol=1e-6)
    assert np.array_equal(r[0, 0, :, 9], -r[0, 0, ::-1, 9]) and np.array_equal(r[0, :, 0, 10], -r[0, ::-1, 0, 10])
    assert np.array_equal(r[1, ..., 8], np.ones((H, W), np.float32)) and np.all(r[..., 6] == 3.0) and np.all(r[..., 7] == 80.0)
    pts = (r[1, ..., :3] + 7.5 * r[1, ..., 3:6]).reshape(-1, 3)
    xy, z, _ = O.project(K, _w2c(c2w[1:2])[0], pts)
    assert np.allclose(xy, r[1, ..., 9:11].reshape(-1, 2), atol=2e-5) and np.all(z > 0)


def test_ssc_head(golden):
    """SURVEY 8f-2 (oracle only so far): SemanticHead.forward(mode="stego_kmeans") -- STEGO code, cosine scores against
    the centres, argmax, pseudo-label LUT -- against the reference's own modules on the reference's 768-d expansions."""
    from scenedino_b200 import synthetic as syn
    g, q = golden("ssc_head"), golden("query")
    x = syn.ssc_head_inputs(q["dino_full"], q["dino_full_le"])
    w = syn.make_ssc_head(int(g["seed"]))

    def checksum(a):
        a = np.asarray(a, np.float64).ravel()
        return np.array([a.sum(), np.abs(a).sum(), a[:: max(1, a.size // 97)].sum()], np.float64)

    assert np.allclose(checksum(x), g["x_sum"], rtol=1e-12) and np.allclose([checksum(w[k]) for k in sorted(w)], g["w_sum"], rtol=1e-12)
    seg, pseudo, ip = O.ssc_head(x, w["wl"], w["bl"], w["wn1"], w["bn1"], w["wn2"], w["bn2"], w["centres"], w["lut"])
    assert seg.dtype == np.int64 and seg.shape == (len(x),)
    assert np.abs(ip - g["ip"]).max() < 2e-5                      # cosine scores, |.| <= 1
    top2 = np.sort(g["ip"], 1)[:, -2:]
    clear = top2[:, 1] - top2[:, 0] > 1e-4                        # labels are compared where the reference's top-2 gap is clear
    assert clear.mean() > 0.98
    assert np.array_equal(pseudo[clear], g["pseudo"][clear]) and np.array_equal(seg[clear], g["seg"][clear])
    assert np.array_equal(seg, w["lut"][pseudo])
    # scaled copies of a row get the row's label (the head normalises its input); the all-zero row is finite
    assert np.array_equal(pseudo[512:576], pseudo[:64]) and np.isfinite(ip[-1]).all()


def test_voxel_grid_is_the_references_grid():
    """synthetic.ssc_voxel_grid (the host construction sd_gen_voxel_grid is tested against, tests/test_gpu_zzz_voxel_grid.py) equals
    the grid the reference builds (sscbench/evaluate_model_sscbench.py:270-278) bit for bit: SHA-256 of all 2 097 152 fp32
    centres, a strided sample, a slab, and a second grid with odd dimensions, size and origin.  The fixture was made by
    running the reference's own functions (oracle/make_golden_grid.py)."""
    import hashlib
    import os
    from scenedino_b200 import synthetic as syn
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxel_grid.npz"))
    assert np.array_equal(syn.velo_to_cam(), g["T"])
    full = syn.ssc_voxel_grid()
    assert full.dtype == np.float32 and full.shape == (256 * 256 * 32, 3)
    assert hashlib.sha256(np.ascontiguousarray(full).tobytes()).digest() == g["sha256_f32"].tobytes()
    assert np.array_equal(full[g["sample_idx"]], g["sample"])
    assert np.array_equal(syn.ssc_voxel_grid(x_range=(37, 101))[:64], g["slab_37_101"])
    kw = dict(dims=tuple(int(d) for d in g["odd_dims"]), voxel_size=float(g["odd_voxel_size"]),
              origin=tuple(float(o) for o in g["odd_origin"]))
    assert np.array_equal(syn.ssc_voxel_grid(**kw), g["odd"])
    # the C oracle's restatement of the same construction (the expression the CUDA kernel evaluates)
    ofull = O.voxel_grid(g["T"])
    assert hashlib.sha256(np.ascontiguousarray(ofull).tobytes()).digest() == g["sha256_f32"].tobytes()
    assert np.array_equal(O.voxel_grid(g["T"], x_range=(37, 101))[:64], g["slab_37_101"])
    assert np.array_equal(O.voxel_grid(g["T"], x_range=(255, 256)), full[255 * 256 * 32:])
    assert np.array_equal(O.voxel_grid(g["T"], **kw), g["odd"])
