"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's golden
vectors.  Needs a B200: run with ``-m gpu``.

Bars (north_star): bit-exact for masks, sample depths, sample indices and sorted merges; rel 1e-4
for fp32-mode densities / features / depth / weights; rel 2e-2 for the reduced-precision (fp16 operand) tensor-core mode.
"""
import numpy as np
import pytest
import torch

from helpers import TOL_F16, TOL_FP32, assert_close, big_query_points, golden_scene_arrays
from oracle import oracle as O
from scenedino_b200 import ops
from scenedino_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

DEV = "cuda"


def sd_launches():
    import scenedino_b200
    return scenedino_b200.launch_count()


def g2n(t):
    return t.detach().cpu().numpy()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def torch_w2c(c2w):
    return torch.inverse(torch.from_numpy(np.ascontiguousarray(c2w))).numpy()


def scenes_from_golden(g, b=None, n=1, feat_dtype=torch.float32, **kw):
    """(oracle scene, device scene, oracle mlp, device mlp) of a golden fixture."""
    feat, imgs = golden_scene_arrays(g, n=n)
    K = g["K"] if b is None else g["K"][b]
    c2w = g["c2w"] if b is None else g["c2w"][b]
    w2c = torch_w2c(c2w)
    i = 0 if b is None else b
    osc = O.Scene(feat=feat[i:i + 1], K_f=K[:1], w2c_f=w2c[:1], rgb=imgs[i], K_c=K, w2c_c=w2c, **kw)
    dkw = dict(kw)
    dsc = ops.Scene.from_arrays(feat[i:i + 1], K[:1], w2c[:1], imgs[i], K, w2c, device=DEV, feat_dtype=feat_dtype, **dkw)
    omlp = O.Mlp(g["w_in"], g["b_in"], g["w_out"], g["b_out"])
    dmlp = ops.Mlp(g["w_in"], g["b_in"], g["w_out"], g["b_out"], device=DEV)
    return osc, dsc, omlp, dmlp


# ------------------------------------------------------------------------------------------------
def test_library_reports_launches():
    from scenedino_b200 import _abi
    assert _abi.lib().sd_device_sm_count() >= 100
    n0 = _abi.launch_count()
    ops.sort_rows(torch.rand(8, 8, device=DEV))
    assert _abi.launch_count() == n0 + 1


def test_project_points_bit_exact(golden):
    g = golden("query")
    w2c = torch_w2c(g["c2w"])
    pts = np.concatenate([g["points"], syn.random_points(3, 50000)], 0)
    xy, z, inv = ops.project_points(dev(g["K"][0]), dev(w2c[0]), dev(pts))
    oxy, oz, oinv = O.project(g["K"][0], w2c[0], pts)
    assert np.array_equal(g2n(inv), oinv)
    assert np.array_equal(g2n(xy), oxy) and np.array_equal(g2n(z), oz)
    n = len(g["points"])   # and against the reference itself
    assert np.array_equal(g2n(inv)[:n], g["frustum_invalid"])
    assert np.array_equal(g2n(xy)[:n], g["xy"]) and np.array_equal(g2n(z)[:n], g["z"])


@pytest.mark.parametrize("tag,learn_empty", [("", False), ("_le", True)])
def test_sample_features(golden, tag, learn_empty):
    g = golden("query")
    kw = dict(learn_empty=learn_empty, empty_feature=g["empty_feature"] if learn_empty else None)
    osc, dsc, _, _ = scenes_from_golden(g, **kw)
    pts = np.concatenate([g["points"], syn.random_points(4, 3001)], 0)
    f, inv = ops.sample_features(dsc, dev(pts))
    of, oinv = O.sample_features(osc, pts)
    f, inv = g2n(f), g2n(inv)
    assert np.array_equal(inv, oinv[:, 0])
    assert np.array_equal(f[:, :256], of[:, 0, :256])                       # gather: bit-exact
    assert np.abs(f[:, 256:] - of[:, 0, 256:]).max() <= 2e-6               # positional code (sinf ulps)
    n = g["sample_features" + tag].shape[0]
    assert np.array_equal(f[:n, :256], g["sample_features" + tag][:, :256])  # vs the reference
    assert np.array_equal(inv[:len(g["points"])], g["sample_invalid" + tag])


def test_sample_colors(golden):
    g = golden("query")
    osc, dsc, omlp, _ = scenes_from_golden(g)
    rgb, inv = ops.sample_colors(dsc, dev(g["points"]))
    q = O.query_points(osc, omlp, g["points"])
    assert np.array_equal(g2n(rgb), q["rgb"]) and np.array_equal(g2n(rgb), g["rgb"])
    # invalid of forward = invalid_colors | invalid_features
    merged = g2n(inv) | g["invalid_features"][:, None]
    assert np.array_equal(merged.astype(np.float32), g["invalid"])


@pytest.mark.parametrize("precision,tol,feat_dtype", [
    (ops.FP32, TOL_FP32, torch.float32),
    (ops.F16, TOL_F16, torch.float16),
])
@pytest.mark.parametrize("tag,learn_empty", [("", False), ("_le", True)])
def test_query_points(golden, precision, tol, feat_dtype, tag, learn_empty):
    g = golden("query")
    kw = dict(learn_empty=learn_empty, empty_feature=g["empty_feature"] if learn_empty else None)
    osc, dsc, omlp, dmlp = scenes_from_golden(g, feat_dtype=feat_dtype, **kw)
    q = ops.query_points(dsc, dmlp, dev(g["points"]), precision=precision)
    o = O.query_points(osc, omlp, g["points"])
    for ref, what in ((o, "oracle"), ({k: g[k + tag] for k in ("sigma", "dino", "rgb", "invalid", "invalid_features")}, "reference")):
        assert_close(g2n(q["sigma"]), ref["sigma"], tol, f"sigma vs {what}")
        assert_close(g2n(q["dino"]), ref["dino"], tol, f"dino vs {what}")
        assert np.array_equal(g2n(q["rgb"]), ref["rgb"])
        assert np.array_equal(g2n(q["invalid"]), ref["invalid"])
        assert np.array_equal(g2n(q["invalid_features"]), ref["invalid_features"])


# ---- projected-map tile kernel (sd_field_project + texel sort + field_bin_kernel) ---------------------------------
def test_field_project_matches_matmul(golden):
    """P = W_in[:, :C] . F per texel, fp16 operands, fp32 accumulation (field_proj.cu) against a torch matmul of the same
    operands; the code-block image carries W_feat . empty_feature in column 47."""
    g = golden("query")
    _, dsc, _, dmlp = scenes_from_golden(g, feat_dtype=torch.float16, learn_empty=True, empty_feature=g["empty_feature"])
    p = dsc.project(dmlp)
    Hf, Wf, C_ = dsc.feat.shape
    P = p.proj[50176:].view(torch.float16).view(Hf * Wf, 128).float()
    W = dev(g["w_in"][:, :C_]).half().float()
    ref = dsc.feat.view(Hf * Wf, C_).float() @ W.T
    assert (P - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    img = g2n(p.proj[32768:49152].view(torch.float16)).astype(np.float32)
    pe = (W @ dev(g["empty_feature"]).half().float()).cpu().numpy()
    for n in (0, 5, 127):   # UMMA K-major SWIZZLE_128B: element (row n, k = 47) sits in chunk (47 >> 3) ^ (n & 7)
        got = img[(n * 128 + ((5 ^ (n & 7)) << 4) + 7 * 2) // 2]
        assert abs(got - pe[n]) <= 2e-3 * max(1.0, abs(pe[n]))


@pytest.mark.parametrize("tag,learn_empty", [("", False), ("_le", True)])
def test_query_points_projected(golden, tag, learn_empty):
    """The tile kernel against the reference's own outputs (golden points, repeated so that the bins fill up) and against
    the oracle on random points: masks and colours bit-exact, densities / features within the reduced-precision bar."""
    from scenedino_b200 import _abi
    g = golden("query")
    kw = dict(learn_empty=learn_empty, empty_feature=g["empty_feature"] if learn_empty else None)
    osc, dsc, omlp, dmlp = scenes_from_golden(g, feat_dtype=torch.float16, **kw)
    dscp = dsc.project(dmlp)
    n_g = len(g["points"])
    pts = np.concatenate([g["points"], g["points"], syn.random_points(7, 70000 - 2 * n_g + 77)]).astype(np.float32)
    n0 = _abi.launch_count()
    q = ops.query_points(dscp, dmlp, dev(pts), precision=ops.F16)
    assert _abi.launch_count() - n0 == 4, "expected sort (3 launches) + field_bin_kernel"
    o = O.query_points(osc, omlp, pts)
    for lo in (0, n_g):   # both copies of the golden points against the reference
        sl = slice(lo, lo + n_g)
        assert_close(g2n(q["sigma"])[sl], g["sigma" + tag], TOL_F16, "sigma vs reference")
        assert_close(g2n(q["dino"])[sl], g["dino" + tag], TOL_F16, "dino vs reference")
        assert np.array_equal(g2n(q["rgb"])[sl], g["rgb" + tag])
        assert np.array_equal(g2n(q["invalid"])[sl], g["invalid" + tag])
        assert np.array_equal(g2n(q["invalid_features"])[sl], g["invalid_features" + tag])
    assert_close(g2n(q["sigma"]), o["sigma"], TOL_F16, "sigma vs oracle")
    assert_close(g2n(q["dino"]), o["dino"], TOL_F16, "dino vs oracle")
    assert np.array_equal(g2n(q["rgb"]), o["rgb"])
    assert np.array_equal(g2n(q["invalid"]), o["invalid"])
    assert np.array_equal(g2n(q["invalid_features"]), o["invalid_features"])
    # a small query on the projected scene takes the gather kernel with the projected map (no sort below 65 536 points)
    qs = ops.query_points(dscp, dmlp, dev(g["points"]), precision=ops.F16)
    assert_close(g2n(qs["sigma"]), g["sigma" + tag], TOL_F16, "sigma, gather kernel on the projected map")
    assert_close(g2n(qs["dino"]), g["dino" + tag], TOL_F16, "dino, gather kernel on the projected map")
    assert np.array_equal(g2n(qs["invalid_features"]), g["invalid_features" + tag])
    # same bits as the gather kernel's masks, and close to its values (two roundings of the same contraction)
    q2 = ops.query_points(dsc, dmlp, dev(pts), precision=ops.F16)
    assert torch.equal(q2["invalid_features"], q["invalid_features"]) and torch.equal(q2["rgb"], q["rgb"])
    assert_close(g2n(q["dino"]), g2n(q2["dino"]), TOL_F16, "tile kernel vs gather kernel")


@pytest.mark.parametrize("tag,learn_empty", [("", False), ("_le", True)])
def test_tile_kernel_vs_reference_big(golden, tag, learn_empty):
    """The texel sort + tile kernel against the REFERENCE's own outputs on a 70 001-point query (fixture query_big,
    oracle/make_golden.py): frustum mask of every point bit-exact, densities / features of the stored subset within 2e-2."""
    from scenedino_b200 import _abi
    g = golden("query_big")
    kw = dict(learn_empty=learn_empty, empty_feature=g["empty_feature"] if learn_empty else None)
    _, dsc, _, dmlp = scenes_from_golden(g, feat_dtype=torch.float16, **kw)
    dscp = dsc.project(dmlp)
    pts, sub = big_query_points(g)
    n0 = _abi.launch_count()
    q = ops.query_points(dscp, dmlp, dev(pts), want_rgb=False, precision=ops.F16)
    assert _abi.launch_count() - n0 == 4
    inv = np.unpackbits(g["invalid_features" + tag])[:len(pts)].astype(bool)
    assert np.array_equal(g2n(q["invalid_features"]), inv)
    assert_close(g2n(q["sigma"])[sub], g["sigma" + tag], TOL_F16, "sigma vs reference")
    assert_close(g2n(q["dino"])[sub], g["dino" + tag], TOL_F16, "dino vs reference")


def test_query_points_sorted_reuses_the_sort(golden):
    """sd_query_points_sorted: same points and cameras, NEW feature map -- one launch (the tile kernel), the bits of a
    full query on the new scene."""
    from scenedino_b200 import _abi
    g = golden("query")
    _, dsc, _, dmlp = scenes_from_golden(g, feat_dtype=torch.float16)
    pts = dev(syn.random_points(13, 70003))
    out = None
    scp = dsc.project(dmlp)
    q1 = ops.query_points(scp, dmlp, pts, want_rgb=False, precision=ops.F16)
    out = {k: v for k, v in q1.items()}
    out["invalid_features"] = out["invalid_features"].view(torch.uint8)
    ops.query_points(scp, dmlp, pts, want_rgb=False, precision=ops.F16, out=out)       # leaves the sort in out["_workspace"]
    feat2 = syn.make_feature_map(77, *[int(v) for v in g["shape"][:3]])
    scp2 = dsc.with_feat_dtype(feat2, torch.float16).project(dmlp)
    ref = ops.query_points(scp2, dmlp, pts, want_rgb=False, precision=ops.F16)
    assert not torch.equal(ref["dino"], q1["dino"])
    n0 = _abi.launch_count()
    q2 = ops.query_points_sorted(scp2, dmlp, pts, out)
    assert _abi.launch_count() - n0 == 1
    torch.cuda.synchronize()
    assert torch.equal(q2["sigma"], ref["sigma"]) and torch.equal(q2["dino"], ref["dino"])
    assert torch.equal(q2["invalid_features"], ref["invalid_features"])


def test_query_graph_replay_matches_direct_calls(golden):
    """ops.QueryGraph: the captured query gives the direct call's bits, and follows the CONTENTS of its input buffer."""
    g = golden("query")
    _, dsc, _, dmlp = scenes_from_golden(g, feat_dtype=torch.float16)
    dscp = dsc.project(dmlp)
    pts = dev(syn.random_points(11, 70001))
    ref = ops.query_points(dscp, dmlp, pts, want_rgb=False, precision=ops.F16)
    out = dict(sigma=torch.empty_like(ref["sigma"]), dino=torch.empty_like(ref["dino"]),
               invalid_features=torch.empty(len(pts), dtype=torch.uint8, device=DEV))
    qg = ops.QueryGraph(dscp, dmlp, pts, out, precision=ops.F16)
    assert qg.launches == 4
    for k in out:
        if not k.startswith("_"):
            out[k].zero_()
    qg.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["sigma"], ref["sigma"]) and torch.equal(out["dino"], ref["dino"])
    assert torch.equal(out["invalid_features"].view(torch.bool), ref["invalid_features"])
    pts2 = dev(syn.random_points(12, 70001))
    ref2 = ops.query_points(dscp, dmlp, pts2, want_rgb=False, precision=ops.F16)
    pts.copy_(pts2)
    qg.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["sigma"], ref2["sigma"]) and torch.equal(out["dino"], ref2["dino"])


@pytest.mark.parametrize("Hf,Wf,d_out,nv_c", [(9, 13, 65, 0), (16, 8, 33, 2), (50, 300, 65, 4), (7, 7, 2, 1)])
def test_tile_kernel_shapes(Hf, Wf, d_out, nv_c):
    """Tile kernel on awkward shapes: maps smaller than / not a multiple of the 7-texel bins (TMA boxes hang over the
    border), narrow heads (D < 64: the per-thread output path; D = 1), 0 to 4 colour views, a ragged last tile."""
    from scenedino_b200 import _abi
    C_ = 256
    feat = syn.make_feature_map(21, C_, Hf, Wf)
    nvc = max(nv_c, 1)
    imgs = syn.make_images(22, nvc, 24, 40)
    K = np.broadcast_to(syn.kitti360_K(), (nvc, 3, 3)).copy()
    c2w = np.stack([syn.view_pose_c2w(v) for v in range(nvc)])
    w2c = np.linalg.inv(c2w.astype(np.float64)).astype(np.float32)
    mlp_w = syn.make_mlp(4, 295, 128, d_out, bias_scale=0.1)
    pts = syn.random_points(9, 65536 + 321)
    if nv_c:
        osc = O.Scene(feat=feat, K_f=K[:1], w2c_f=w2c[:1], rgb=imgs, K_c=K, w2c_c=w2c)
        dsc = ops.Scene.from_arrays(feat, K[:1], w2c[:1], imgs, K, w2c, device=DEV, feat_dtype=torch.float16)
    else:
        osc = O.Scene(feat=feat, K_f=K[:1], w2c_f=w2c[:1])
        dsc = ops.Scene.from_arrays(feat, K[:1], w2c[:1], device=DEV, feat_dtype=torch.float16)
    dmlp = ops.Mlp(*mlp_w, device=DEV)
    dscp = dsc.project(dmlp)
    n0 = _abi.launch_count()
    q = ops.query_points(dscp, dmlp, dev(pts), precision=ops.F16, want_rgb=nv_c > 0)
    assert _abi.launch_count() - n0 == 4, "expected the texel sort + the tile kernel"
    o = O.query_points(osc, O.Mlp(*mlp_w), pts, want_rgb=nv_c > 0)
    assert q["dino"].shape == (len(pts), d_out - 1)
    assert np.array_equal(g2n(q["invalid_features"]), o["invalid_features"])
    assert_close(g2n(q["sigma"]), o["sigma"], TOL_F16, "sigma")
    if d_out > 1 + 0:
        assert_close(g2n(q["dino"]), o["dino"], TOL_F16, "dino")
    if nv_c:
        assert np.array_equal(g2n(q["rgb"]), o["rgb"]) and np.array_equal(g2n(q["invalid"]), o["invalid"])
    assert torch.isfinite(q["sigma"]).all() and torch.isfinite(q["dino"]).all()


@pytest.mark.parametrize("precision,tol", [(ops.FP32, TOL_FP32), (ops.F16, TOL_F16)])
@pytest.mark.parametrize("d_in,d_out,n", [(295, 65, 1000), (295, 769, 300), (64, 768, 257), (40, 3, 65), (312, 33, 129)])
def test_mlp_forward(precision, tol, d_in, d_out, n):
    w = syn.make_mlp(5, d_in, 128, d_out, bias_scale=0.2)
    x = np.random.RandomState(1).standard_normal((n, d_in)).astype(np.float32)
    mlp = ops.Mlp(*w, device=DEV)
    if precision == ops.F16 and d_out - 1 > 64:
        from scenedino_b200 import SdError
        with pytest.raises(SdError, match="d_out"):      # the tensor-core head covers d_out <= 65; wider heads run in fp32
            ops.mlp_forward(mlp, dev(x), precision=precision)
        return
    out = ops.mlp_forward(mlp, dev(x), precision=precision)
    assert_close(g2n(out), O.mlp_forward(O.Mlp(*w), x), tol, "mlp")


def test_expand_dim(golden):
    g = golden("query")
    m = ops.Mlp(g["e_w1"], g["e_b1"], g["e_w2"], g["e_b2"], device=DEV)
    out = g2n(ops.expand_dim(m, dev(g["dino"][:256])))
    assert_close(out, g["dino_full"], TOL_FP32, "expand_dim vs reference")
    np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, rtol=1e-5)


def test_expand_dim_tensor_cores_vs_reference(golden):
    """The reference's own 768-d expansion of 256 queried features (tests/golden/query.npz), fp16-operand kernel."""
    g = golden("query")
    m = ops.Mlp(g["e_w1"], g["e_b1"], g["e_w2"], g["e_b2"], device=DEV)
    n0 = sd_launches()
    out = g2n(ops.expand_dim(m, dev(g["dino"][:256]), precision=ops.F16))
    assert sd_launches() - n0 == 1
    assert_close(out, g["dino_full"], TOL_F16, "expand_dim (tensor cores) vs reference")
    np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, rtol=1e-5)


@pytest.mark.parametrize("N,d_full", [(1, 768), (127, 768), (129, 768), (1000, 128), (70001, 768), (19000, 1024)])
def test_expand_dim_tensor_cores_shapes(N, d_full):
    """Ragged tiles, several tiles per CTA (70 001 rows = 547 tiles on 148 CTAs), one to eight 128-column blocks;
    rows that are exactly zero (norm clamp) and large inputs; against the oracle and the fp32 CUDA-core path."""
    rs = np.random.RandomState(N + d_full)
    w = [(rs.randn(128, 64) * 0.15).astype(np.float32), (rs.randn(128) * 0.1).astype(np.float32),
         (rs.randn(d_full, 128) * 0.1).astype(np.float32), (rs.randn(d_full) * 0.05).astype(np.float32)]
    f = (rs.randn(N, 64) * rs.choice([0.1, 1.0, 8.0], (N, 1))).astype(np.float32)
    m = ops.Mlp(*w, device=DEV)
    out = g2n(ops.expand_dim(m, dev(f), precision=ops.F16))
    want = O.expand_dim(f, *w)
    assert out.shape == (N, d_full) and np.isfinite(out).all()
    assert_close(out, want, TOL_F16, "expand_dim (tensor cores) vs oracle")
    np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, rtol=1e-5)
    assert_close(out, g2n(ops.expand_dim(m, dev(f))), TOL_F16, "expand_dim (tensor cores) vs fp32 path")
    # a permuted batch returns the permuted result bit for bit (rows are independent of their tile position)
    perm = rs.permutation(N)
    assert np.array_equal(g2n(ops.expand_dim(m, dev(f[perm]), precision=ops.F16)), out[perm])
    assert ops.expand_dim(m, dev(f[:0]), precision=ops.F16).shape == (0, d_full)


def test_expand_dim_unsupported_shape_takes_the_fp32_path():
    rs = np.random.RandomState(3)
    w = [rs.randn(128, 32).astype(np.float32) * 0.1, np.zeros(128, np.float32), rs.randn(96, 128).astype(np.float32) * 0.1,
         np.zeros(96, np.float32)]
    f = rs.randn(300, 32).astype(np.float32)
    m = ops.Mlp(*w, device=DEV)
    assert np.array_equal(g2n(ops.expand_dim(m, dev(f), precision=ops.F16)), g2n(ops.expand_dim(m, dev(f))))


# ---- sampling: all bit-exact -------------------------------------------------------------------
@pytest.mark.parametrize("lindisp", [True, False])
def test_sampling_bit_exact(lindisp):
    rs = np.random.RandomState(7)
    R, Kc, Kf, Kfd = 777, 64, 32, 8
    rays = syn.image_rays(syn.view_pose_c2w(3), syn.kitti360_K())[rs.choice(192 * 640, R, replace=False)]
    rays[:, 6] = rs.uniform(0.5, 4, R); rays[:, 7] = rs.uniform(30, 100, R)
    u = rs.uniform(0, 1, (R, Kc)).astype(np.float32)
    lin = torch.linspace(0, 1 - 1.0 / Kc, Kc).numpy()
    z = ops.sample_coarse(dev(rays), dev(u), dev(lin), lindisp)
    oz = O.sample_coarse(rays, u, lin, lindisp)
    assert np.array_equal(g2n(z), oz)
    w = (rs.uniform(0, 1, (R, Kc)) ** 6).astype(np.float32)
    w[:5] = 0.0; w[5, 17] = 1.0
    u0 = rs.uniform(0, 1, (R, Kf)).astype(np.float32); u0[0, 0] = 0.0; u0[1, 1] = np.float32(1.0 - 2 ** -24)
    u1 = rs.uniform(0, 1, (R, Kf)).astype(np.float32)
    zf, inds = ops.sample_fine(dev(rays), dev(w), dev(u0), dev(u1), lindisp)
    ozf, oinds = O.sample_fine(rays, w, u0, u1, lindisp)
    assert np.array_equal(g2n(inds), oinds) and np.array_equal(g2n(zf), ozf)
    depth = rs.uniform(2, 90, R).astype(np.float32)
    noise = rs.standard_normal((R, Kfd)).astype(np.float32)
    zd = ops.sample_fine_depth(dev(rays), dev(depth), dev(noise), 0.5)
    assert np.array_equal(g2n(zd), O.sample_fine_depth(rays, depth, noise, 0.5))
    zp = np.sort(rs.uniform(3, 80, (R, 48)).astype(np.float32), 1)
    wp = (rs.uniform(0, 1, (R, 48)) ** 4).astype(np.float32)
    u0p = rs.uniform(0, 1, (R, 40)).astype(np.float32); u1p = rs.uniform(0, 1, (R, 40)).astype(np.float32)
    zz, ii = ops.sample_coarse_from_dist(dev(wp), dev(zp), dev(u0p), dev(u1p), lindisp)
    ozz, oii = O.sample_coarse_from_dist(wp, zp, u0p, u1p, lindisp)
    assert np.array_equal(g2n(ii), oii) and np.array_equal(g2n(zz), ozz)
    allz = np.concatenate([oz, ozf, g2n(zd)], 1)
    assert np.array_equal(g2n(ops.sort_rows(dev(allz))), np.sort(allz, 1))
    for K in (1, 2, 3, 33, 100):
        a = rs.standard_normal((19, K)).astype(np.float32)
        assert np.array_equal(g2n(ops.sort_rows(dev(a))), np.sort(a, 1))


def test_sample_fine_vs_reference(golden):
    for name in ("render_fine", "render_fine_lin"):
        g = golden(name)
        lindisp = bool(g["conf"][3])
        rays = g["rays"][0]
        zf, inds = ops.sample_fine(dev(rays), dev(g["coarse.weights"][0]), dev(g["u_fine0"]), dev(g["u_fine1"]), lindisp)
        flips = (g2n(inds) != g["fine_inds"])
        assert flips.sum() <= max(1, int(1e-4 * flips.size))
        zd = ops.sample_fine_depth(dev(rays), dev(g["coarse.depth"][0]), dev(g["n_depth"]), float(g["depth_std"]))
        zall = ops.sort_rows(torch.cat([dev(g["coarse.z_samps"][0]), zf, zd], 1))
        same = ~flips.any(1)
        assert np.array_equal(g2n(zall)[same], g["fine.z_samps"][0][same])


# ---- composite ---------------------------------------------------------------------------------
@pytest.mark.parametrize("R,K,D,Crgb", [(257, 64, 64, 3), (33, 96, 768, 12), (100, 1, 8, 0), (64, 37, 100, 6), (5, 130, 64, 3)])
@pytest.mark.parametrize("cap,white", [(False, False), (True, True)])
def test_composite_vs_oracle(R, K, D, Crgb, cap, white):
    rs = np.random.RandomState(R + K)
    z = np.sort(rs.uniform(3, 80, (R, K)).astype(np.float32), 1)
    sigma = (rs.uniform(0, 1, (R, K)) ** 3 * 0.5).astype(np.float32)
    sigma[0] = 0; sigma[1, K // 2] = 1e4; sigma[2] = -1.0
    feat = rs.standard_normal((R, K, D)).astype(np.float32)
    rgb = rs.uniform(0, 1, (R, K, Crgb)).astype(np.float32) if Crgb else None
    o = O.composite(z, sigma, feat, rgb, cap, white)
    c = ops.composite(dev(z), dev(sigma), dev(feat), None if rgb is None else dev(rgb), cap, white)
    for k in ("weights", "alphas", "depth", "dino"):
        assert_close(g2n(c[k]), o[k], TOL_FP32, k)
    if Crgb:
        assert_close(g2n(c["rgb"]), o["rgb"], TOL_FP32, "rgb")


def _check_pass(o, g, prefix, tol, b=None):
    sl = (lambda a: a) if b is None else (lambda a: a[b:b + 1])
    R = sl(g["rays"]).shape[0] * g["rays"].shape[1]
    K = g[prefix + "z_samps"].shape[-1]
    assert_close(g2n(o["weights"]), sl(g[prefix + "weights"]).reshape(R, K), tol, prefix + "weights")
    assert_close(g2n(o["alphas"]), sl(g[prefix + "alphas"]).reshape(R, K), tol, prefix + "alphas")
    assert_close(g2n(o["depth"]), sl(g[prefix + "depth"]).reshape(R), tol, prefix + "depth")
    assert_close(g2n(o["rgb"]), sl(g[prefix + "rgb"]).reshape(R, -1), tol, prefix + "rgb")
    assert_close(g2n(o["dino_features"]), sl(g[prefix + "dino_features"]).reshape(R, -1), tol, prefix + "dino")
    assert np.array_equal(g2n(o["invalid"]), sl(g[prefix + "invalid"]).reshape(R, K, -1))
    assert np.array_equal(g2n(o["invalid_features"]).ravel(), sl(g[prefix + "invalid_features"]).ravel())


@pytest.mark.parametrize("precision,tol,feat_dtype,projected", [
    (ops.FP32, TOL_FP32, torch.float32, False),
    (ops.F16, TOL_F16, torch.float16, False),
    (ops.F16, TOL_F16, torch.float16, True),     # gather kernel on the projected map (128 channels, identity layer 1)
])
def test_render_pass_vs_reference(golden, precision, tol, feat_dtype, projected):
    g = golden("render_coarse")
    osc, dsc, omlp, dmlp = scenes_from_golden(g, feat_dtype=feat_dtype)
    if projected:
        dsc = dsc.project(dmlp)
    rays = g["rays"][0]
    z = ops.sample_coarse(dev(rays), dev(g["u_coarse"]), dev(g["lin"]), True)
    assert np.array_equal(g2n(z), g["coarse.z_samps"][0])
    o = ops.render_pass(dsc, dmlp, dev(rays), z, hard_alpha_cap=False, want_rgb_samps=True, precision=precision)
    _check_pass(o, g, "coarse.", tol)
    assert np.array_equal(g2n(o["rgb_samps"]), g["coarse.rgb_samps"][0])
    oo = O.render_pass(osc, omlp, rays, g2n(z), want_rgb_samps=True)
    for k in ("weights", "alphas", "depth", "dino_features", "rgb", "sigma"):
        assert_close(g2n(o[k]), oo[k], tol, k + " vs oracle")
    # per-ray-only call (what the throughput path uses) gives the same per-ray results
    o2 = ops.render_pass(dsc, dmlp, dev(rays), z, per_sample=False, precision=precision)
    for k in ("depth", "dino_features", "rgb"):
        assert_close(g2n(o2[k]), g2n(o[k]), 1e-6, k + " per-ray-only")
    for name in ("render_fine", "render_fine_lin"):
        g = golden(name)
        _, dsc, _, dmlp = scenes_from_golden(g, feat_dtype=feat_dtype)
        if projected:
            dsc = dsc.project(dmlp)
        white = bool(g["conf"][4])
        for p in ("coarse.", "fine."):
            o = ops.render_pass(dsc, dmlp, dev(g["rays"][0]), dev(g[p + "z_samps"][0]), hard_alpha_cap=True,
                                white_bkgd=white, precision=precision)
            _check_pass(o, g, p, tol)


def test_render_superbatch_vs_reference(golden):
    g = golden("render_superbatch")
    for b in range(2):
        _, dsc, _, dmlp = scenes_from_golden(g, b=b, n=2)
        o = ops.render_pass(dsc, dmlp, dev(g["rays"][b]), dev(g["coarse.z_samps"][b]), hard_alpha_cap=True)
        _check_pass(o, g, "coarse.", TOL_FP32, b=b)


# ---- BASELINE-size properties --------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol,feat_dtype,projected,map_hw", [
    (ops.FP32, TOL_FP32, torch.float32, False, (192, 640)),
    (ops.F16, TOL_F16, torch.float16, False, (192, 640)),
    (ops.F16, TOL_F16, torch.float16, True, (192, 640)),
    # the map bench.py times (DINO ViT-B/8: 384 x 1280): TMA box coordinates, bin counts and the 16-bit compact bin ids of
    # the sort all scale with it
    (ops.F16, TOL_F16, torch.float16, True, (384, 1280)),
    (ops.FP32, TOL_FP32, torch.float32, False, (384, 1280)),
])
def test_ssc_grid_full_size(precision, tol, feat_dtype, projected, map_hw):
    """configs[1]: the 256x256x32 voxel grid against a DINOv2-sized and a ViT-B/8-sized map: masks bit-exact on all
    2 097 152 voxels, values against the oracle on a strided subset, and batch-position independence
    (a permuted query returns the permuted result bit for bit)."""
    C_, (Hf, Wf) = 256, map_hw
    feat = syn.make_feature_map(1, C_, Hf, Wf)
    K = syn.kitti360_K()[None]; w2c = np.eye(4, dtype=np.float32)[None]
    mlp_w = syn.make_mlp(0, bias_scale=0.05)
    pts = syn.ssc_voxel_grid()
    assert pts.shape == (2097152, 3)
    dsc = ops.Scene.from_arrays(feat, K, w2c, device=DEV, feat_dtype=feat_dtype)
    dmlp = ops.Mlp(*mlp_w, device=DEV)
    if projected:   # the path bench.py times: texel sort + projected-map tile kernel
        dsc = dsc.project(dmlp)
    dp = dev(pts)
    q = ops.query_points(dsc, dmlp, dp, precision=precision)
    _, _, oinv = O.project(K[0], w2c[0], pts)
    assert np.array_equal(g2n(q["invalid_features"]), oinv)
    assert 0.15 < oinv.mean() < 0.35
    sub = np.arange(0, len(pts), 257)
    o = O.query_points(O.Scene(feat=feat, K_f=K, w2c_f=w2c), O.Mlp(*mlp_w), pts[sub])
    assert_close(g2n(q["sigma"])[sub], o["sigma"], tol, "sigma")
    assert_close(g2n(q["dino"])[sub], o["dino"], tol, "dino")
    perm = torch.randperm(len(pts), device=DEV, generator=torch.Generator(DEV).manual_seed(0))
    q2 = ops.query_points(dsc, dmlp, dp[perm].contiguous(), precision=precision)
    assert torch.equal(q2["sigma"], q["sigma"][perm]) and torch.equal(q2["dino"], q["dino"][perm])
    assert torch.isfinite(q["sigma"]).all() and torch.isfinite(q["dino"]).all()
    # the tensor-core path walks the points in texel-binned order when given scratch space: same bits (without scratch
    # space a projected scene falls back to the gather kernel: same masks, values within the bar)
    q3 = ops.query_points(dsc, dmlp, dp, precision=precision, binned=False)
    assert torch.equal(q3["invalid_features"], q["invalid_features"])
    if projected:
        assert_close(g2n(q3["sigma"])[sub], g2n(q["sigma"])[sub], tol, "gather vs tile kernel")
    else:
        assert torch.equal(q3["sigma"], q["sigma"]) and torch.equal(q3["dino"], q["dino"])


# ---- edge cases and error behaviour ----------------------------------------------------------------
def test_edge_sizes(golden):
    g = golden("query")
    osc, dsc, omlp, dmlp = scenes_from_golden(g)
    for n in (0, 1, 63, 64, 65, 129):
        pts = g["points"][:n]
        q = ops.query_points(dsc, dmlp, dev(pts).reshape(n, 3))
        assert q["sigma"].shape == (n,) and q["dino"].shape == (n, 64)
        if n:
            o = O.query_points(osc, omlp, pts)
            assert_close(g2n(q["sigma"]), o["sigma"], TOL_FP32, f"sigma n={n}")
    rays = g["points"][:0].reshape(0, 3)
    o = ops.render_pass(dsc, dmlp, torch.zeros(0, 11, device=DEV), torch.zeros(0, 16, device=DEV))
    assert o["depth"].shape == (0,)


def test_errors_are_loud(golden):
    from scenedino_b200 import SdError
    g = golden("query")
    _, dsc, _, dmlp = scenes_from_golden(g)
    with pytest.raises(SdError):
        ops.query_points(dsc, dmlp, torch.zeros(4, 3))                       # CPU tensor: no fallback
    bad = ops.Mlp(*syn.make_mlp(0, d_in=100), device=DEV)
    with pytest.raises(SdError, match="d_in"):
        ops.query_points(dsc, bad, torch.zeros(4, 3, device=DEV))
    with pytest.raises(SdError):
        ops.sort_rows(torch.zeros(2, 5000, device=DEV))


# ---- rays of whole views on the device ------------------------------------------------------------------
def test_gen_rays_matches_reference_and_oracle(golden):
    """sd_gen_rays against the reference's ImageRaySampler.sample (tests/golden/rays.npz) and the oracle: bit-exact,
    all eleven columns, odd image sizes / explicit frame ids / unnormalised directions included."""
    g = golden("rays")
    for tag in "ab":
        H, W = (int(x) for x in g[f"{tag}_hw"])
        ids = g[f"{tag}_ids"] if g[f"{tag}_ids"].size else None
        norm = bool(g[f"{tag}_norm_dir"])
        for i in range(g[f"{tag}_c2w"].shape[0]):
            c2w, proj = g[f"{tag}_c2w"][i], g[f"{tag}_proj"][i]
            r = g2n(ops.gen_rays(dev(c2w), dev(proj), H, W, 3.0, 80.0, frame_ids=None if ids is None else dev(ids), norm_dir=norm))
            assert np.array_equal(r, g[f"{tag}_rays"][i]), (tag, i)
            assert np.array_equal(r, O.gen_rays(c2w, proj, H, W, 3.0, 80.0, frame_ids=ids, norm_dir=norm))


@pytest.mark.parametrize("V,H,W", [(1, 2, 2), (3, 7, 5), (2, 192, 640), (4, 376, 1408)])
def test_gen_rays_shapes_and_shift(V, H, W):
    """Ragged block tails (ray counts that are not multiples of 256 or 4), the full KITTI-360 image sizes, and the
    sub-pixel shift, bit-exact against the oracle; the launch counter sees exactly one kernel."""
    rng = np.random.default_rng(V * 1000 + H)
    c2w = np.stack([syn.view_pose_c2w(v) for v in range(V)], 0).astype(np.float32)
    proj = np.broadcast_to(syn.kitti360_K(), (V, 3, 3)).astype(np.float32).copy()
    proj[:, 0, 2] += rng.uniform(-0.05, 0.05, V).astype(np.float32)
    ids = rng.integers(0, 9, V).astype(np.float32)
    for shift in ((0.0, 0.0), (0.25 * 2 / W, -0.5 * 2 / H)):
        n0 = sd_launches()
        r = ops.gen_rays(dev(c2w), dev(proj), H, W, 0.5, 120.0, frame_ids=dev(ids), xy_shift=shift)
        assert sd_launches() - n0 == 1
        want = O.gen_rays(c2w, proj, H, W, 0.5, 120.0, frame_ids=ids, xy_shift=shift)
        assert np.array_equal(g2n(r), want)
    assert np.allclose(np.linalg.norm(want[:, 3:6], axis=1), 1.0, atol=1e-6)
    assert ops.gen_rays(dev(c2w[:0]), dev(proj[:0]), H, W, 0.5, 120.0).shape == (0, 11)
