"""GPU parity of the fused expansion + unsupervised SSC head (SURVEY 8f-1 + 8f-2, csrc/ssc_head.cu) and of
PositionalEncoding.forward.  Needs a B200: ``-m gpu``."""
import numpy as np
import pytest
import torch

import scenedino_b200 as sd
from helpers import TOL_F16
from oracle import oracle as O
from scenedino_b200 import _abi, ops
from scenedino_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def oracle_head(f, expand, w):
    x = O.expand_dim(f, *expand)
    return O.ssc_head(x, w["wl"], w["bl"], w["wn1"], w["bn1"], w["wn2"], w["bn2"], w["centres"], w["lut"])


def check_against(out, seg, pseudo, ip, margin=4 * TOL_F16 / 10):
    """scores within the reduced-precision bar (they are cosines, |.| <= 1: absolute); labels equal wherever the reference's
    top-2 gap is clear of twice the measured score error."""
    sc = out["scores"].cpu().numpy()
    err = float(np.abs(sc - ip).max())
    assert err <= TOL_F16, f"cosine scores off by {err:.3e}"
    top2 = np.sort(ip, 1)[:, -2:]
    clear = top2[:, 1] - top2[:, 0] > max(margin, 2.5 * err)
    assert clear.mean() > 0.7
    ps, sg = out["pseudo"].cpu().numpy().astype(np.int64), out["seg"].cpu().numpy().astype(np.int64)
    assert np.array_equal(ps[clear], pseudo[clear]) and np.array_equal(sg[clear], seg[clear])
    assert np.array_equal(ps, sc.argmax(1)), "labels are the first maximum of the scores the kernel reports"
    return err


def test_ssc_head_vs_reference(golden):
    """The reference's own StegoClusterHead / KMeansParamHead outputs (tests/golden/ssc_head.npz) on the expansions of the
    golden query's 64-d features: the fused kernel starts from those 64-d features."""
    g, q = golden("ssc_head"), golden("query")
    w = syn.make_ssc_head(int(g["seed"]))
    nq = len(q["dino_full"])                       # the fixture expands the first 256 points of each variant
    f = np.concatenate([q["dino"][:nq], q["dino_le"][:nq]], 0).astype(np.float32)
    n = len(f)
    expand = (q["e_w1"], q["e_b1"], q["e_w2"], q["e_b2"])
    # the fixture's first rows are the reference's expansions of exactly these features
    x = O.expand_dim(f, *expand)
    ref_x = syn.ssc_head_inputs(q["dino_full"], q["dino_full_le"])[:n]
    assert np.abs(x - ref_x).max() < 2e-5
    head = ops.SscHead(expand, w, device=DEV)
    out = ops.ssc_head(head, dev(f))
    err = check_against(out, g["seg"][:n], g["pseudo"][:n], g["ip"][:n])
    assert err < 5e-3, "measured ~3e-4 in emulation: anything near the bar is a bug"
    assert np.array_equal(out["seg"].cpu().numpy(), w["lut"][out["pseudo"].cpu().numpy().astype(np.int64)])


@pytest.mark.parametrize("n", [1, 127, 128, 129, 256, 385, 70001])
def test_ssc_head_vs_oracle(n):
    """Ragged sizes, odd / even tile counts (the kernel keeps two tiles in flight per CTA), inputs of mixed scale."""
    rs = np.random.RandomState(n)
    f = (rs.standard_normal((n, 64)) * rs.choice([0.05, 1.0, 6.0], size=(n, 1))).astype(np.float32)
    expand, w = syn.make_expand(3), syn.make_ssc_head(7)
    seg, pseudo, ip = oracle_head(f, expand, w)
    head = ops.SscHead(expand, w, device=DEV)
    n0 = _abi.launch_count()
    out = ops.ssc_head(head, dev(f))
    assert _abi.launch_count() - n0 == 1
    check_against(out, seg, pseudo, ip)
    # labels only + scattered through a permutation: out[perm[r]] = label of row r
    perm = torch.randperm(n, device=DEV, generator=torch.Generator(DEV).manual_seed(1)).to(torch.int32)
    o2 = ops.ssc_head(head, dev(f), want_scores=False, perm=perm)
    assert torch.equal(o2["seg"][perm.long()], out["seg"])


def test_ssc_head_27_and_32_clusters_and_small_heads():
    rs = np.random.RandomState(3)
    f = rs.standard_normal((1000, 64)).astype(np.float32)
    for n_cls, d_full in ((32, 768), (5, 768), (27, 256)):
        expand = syn.make_expand(4, 64, 128, d_full)
        w = syn.make_ssc_head(9, d_in=d_full, n_cls=n_cls)
        seg, pseudo, ip = oracle_head(f, expand, w)
        out = ops.ssc_head(ops.SscHead(expand, w, device=DEV), dev(f))
        check_against(out, seg, pseudo, ip)
    with pytest.raises(sd.SdError):
        ops.SscHead(syn.make_expand(4), syn.make_ssc_head(9, n_cls=40), device=DEV)


def test_semantic_head_module_and_btsnet_forward(golden):
    """BTSNet.forward(predict_segmentation=True) (models/bts.py:584-592) with scenedino_b200.SemanticHead as the downstream
    head: (dino_full, None, sigma, one-hot seg) like the reference; the head runs in sd_ssc_head, not in PyTorch."""
    from test_gpu_surface import build
    g = golden("query")
    net = build(g, precision="fp16")
    w = syn.make_ssc_head(21)
    head = sd.SemanticHead.from_conf({"n_classes": 27, "gt_classes": 19, "input_dim": 768, "code_dim": 64})
    state = {"stego_head.linear_path.0.weight": w["wl"].reshape(64, 768, 1, 1), "stego_head.linear_path.0.bias": w["bl"],
             "stego_head.nonlinear_path.0.weight": w["wn1"].reshape(768, 768, 1, 1), "stego_head.nonlinear_path.0.bias": w["bn1"],
             "stego_head.nonlinear_path.2.weight": w["wn2"].reshape(64, 768, 1, 1), "stego_head.nonlinear_path.2.bias": w["bn2"],
             "stego_cluster_head.cluster_centers": w["centres"], "stego_cluster_head.pseudo_assignment": w["lut"]}
    missing, unexpected = head.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=False)
    assert not unexpected
    net.downstream_head, net.gt_classes = head.to(DEV), 19
    pts = dev(g["points"])[None]
    with torch.no_grad():
        dino_full, none, sigma, seg = net(pts, predict_segmentation=True)
    N = pts.shape[1]
    assert none is None and dino_full.shape == (1, N, 768) and sigma.shape == (1, N, 1)
    assert seg.shape == (1, N, 19) and seg.dtype == torch.int64 and bool((seg.sum(-1) == 1).all())
    # the same labels from the oracle on the oracle's features, wherever the top-2 gap is clear
    expand = (g["e_w1"], g["e_b1"], g["e_w2"], g["e_b2"])
    oseg, opseudo, oip = oracle_head(g["dino"], expand, w)
    top2 = np.sort(oip, 1)[:, -2:]
    clear = top2[:, 1] - top2[:, 0] > 2e-2
    assert clear.mean() > 0.5
    assert np.array_equal(seg[0].argmax(-1).cpu().numpy()[clear], oseg[clear])
    # the caller-facing module form: SemanticHead.forward on the tensor expand_dim returned
    with torch.no_grad():
        lab = net.downstream_head(dino_full, mode="stego_kmeans")
    assert lab.shape == (1, N) and lab.dtype == torch.int64 and torch.equal(lab, seg.argmax(-1))
    with pytest.raises(NotImplementedError):
        net.downstream_head(dino_full.clone())            # a bare 768-d tensor carries no 64-d source
    # without the 768-d rows (what the SSC evaluation needs: sigma + seg)
    net.materialize_dino_full = False
    n0 = _abi.launch_count()
    with torch.no_grad():
        none_full, _, sigma2, seg2 = net(pts, predict_segmentation=True)
    assert none_full is None and torch.equal(seg2, seg) and torch.equal(sigma2, sigma)
    assert _abi.launch_count() - n0 <= 3, "field query + fused head (no expansion kernel)"


def test_positional_encoding_forward():
    """positional_encoding.py:68-80 against the same torch ops on the CPU."""
    pe = sd.PositionalEncoding.from_conf({"num_freqs": 6, "freq_factor": 1.5, "include_input": True}, d_in=3).to(DEV)
    rs = np.random.RandomState(0)
    x = torch.from_numpy(np.concatenate([rs.uniform(-2, 2, (5000, 3)), rs.uniform(-6000, 6000, (64, 3))]).astype(np.float32))
    out = pe(x.to(DEV)).cpu()
    emb = x.unsqueeze(1).repeat(1, 12, 1)
    ref = torch.sin(torch.addcmul(pe._phases.cpu(), emb, pe._freqs.cpu())).view(x.shape[0], -1)
    ref = torch.cat((x, ref), -1)
    assert out.shape == ref.shape == (5064, 39)
    assert torch.equal(out[:, :3], x)
    assert float((out[:5000] - ref[:5000]).abs().max()) <= 2e-6          # sinf ulps
    assert float((out[5000:] - ref[5000:]).abs().max()) <= 2e-3          # |arg| up to 3e5 rad: ill-conditioned
    pe2 = sd.PositionalEncoding(num_freqs=4, d_in=2, freq_factor=np.pi, include_input=False).to(DEV)
    y = pe2(x[:100, :2].to(DEV))
    assert y.shape == (100, 16)
