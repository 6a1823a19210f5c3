"""Shared helpers for the parity tests: tolerances and scene reconstruction from golden fixtures."""
import numpy as np

from scenedino_b200 import synthetic as syn

# north_star tolerances: rel 1e-4 in fp32 mode, 2e-2 in the reduced-precision (fp16 tensor-core) mode.  "rel" is measured against the
# magnitude of the reference tensor: |a-b| <= tol * max(|b|, rms(b)); near-zero entries of a tensor
# are therefore compared against the tensor's own scale rather than against themselves.
TOL_FP32 = 1e-4
TOL_F16 = 2e-2


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.sqrt(np.mean(b * b)) if b.size else 1.0
    den = np.maximum(np.abs(b), max(scale, 1e-30))
    return np.abs(a - b) / den


#: element-wise bar on the entries that are not near zero (|b| > 0.1 rms): |a-b| <= ELEMENTWISE * tol * |b|.  The scale-relative
#: measure above alone would let an entry of 0.1 rms be off by 10 tol of itself.
ELEMENTWISE = 5.0


def assert_close(a, b, tol, what=""):
    assert np.shape(a) == np.shape(b), f"{what}: shape {np.shape(a)} vs {np.shape(b)}"
    if np.size(b) == 0:
        return
    e = rel_err(a, b)
    worst = float(e.max())
    assert worst <= tol, f"{what}: max rel err {worst:.3e} > {tol:.1e} at {np.unravel_index(e.argmax(), e.shape)}"
    a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.sqrt(np.mean(b64 * b64))
    big = np.abs(b64) > 0.1 * scale
    if big.any():
        ew = np.abs(a64 - b64)[big] / np.abs(b64)[big]
        assert float(ew.max()) <= ELEMENTWISE * tol, \
            f"{what}: element-wise rel err {float(ew.max()):.3e} > {ELEMENTWISE * tol:.1e} on an entry with |b| > 0.1 rms"


def checksum(a):
    a = np.asarray(a, np.float64).ravel()
    return np.array([a.sum(), np.abs(a).sum(), a[:: max(1, a.size // 97)].sum()], np.float64)


def golden_scene_arrays(g, n=1):
    """Regenerates the seeded feature maps / images of a fixture and checks their checksums."""
    C, HF, WF, HC, WC, nv_c = [int(v) for v in g["shape"]]
    feat = np.concatenate([syn.make_feature_map(11 + i, C, HF, WF) for i in range(n)], 0)
    imgs = np.stack([syn.make_images(12 + i, nv_c, HC, WC) for i in range(n)], 0)
    np.testing.assert_allclose(checksum(feat), g["feat_checksum"], rtol=1e-12)
    np.testing.assert_allclose(checksum(imgs), g["img_checksum"], rtol=1e-12)
    return feat, imgs


def w2c_of(c2w):
    return np.linalg.inv(c2w.astype(np.float64)).astype(np.float32)


def big_query_points(g):
    """Regenerates the points of the query_big fixture from their seed and checks their checksum."""
    pts = syn.random_points(int(g["pts_seed"]), int(g["n_pts"]))
    np.testing.assert_allclose(checksum(pts), g["pts_checksum"], rtol=1e-12)
    sub = np.arange(0, len(pts), int(g["stride"]))
    return pts, sub
